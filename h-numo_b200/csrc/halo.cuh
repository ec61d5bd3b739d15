// halo.cuh -- face-halo exchange between element partitions.
//
// Replaces the MPI pack / Isend / Irecv / Waitall / unpack of the reference
//   btp_create_{pre,post}communicator, create_rhs_lap_{pre,post}communicator_df, bcl_create_communicator
//       src/create_rhs_communicator.F90:193-327,384-438
//   pack_data_dg_df, unpack_data_dg_general_df, send_bound_dg_general_df, *_quad_layer
//       src/send_receive_bound.F90:61-108,272-327,455-509,806-969
//   create_nbhs_face_df / _quad_layer ("side 2 := neighbour's side 1")   src/create_rhs_dynamics_flux.F90:104-182,509-591
// One message per neighbour rank per exchange, in the face order of mod_parallel's nbh_send_recv.  State and LDG
// gradient traces travel in ONE message per stage (the reference sends two).
//
// Back ends: NCCL point-to-point (ncclSend/ncclRecv grouped, loaded with dlopen so that the library also loads on
// machines without NCCL) and an in-process "local" back end (several partitions on one GPU, host-side barrier) used
// by the parity tests on single-GPU boxes.
#pragma once
#include <dlfcn.h>
#include <pthread.h>

#include <condition_variable>
#include <mutex>

#include "hnumo_dev.cuh"

namespace hn {

inline int sops_doubles_host(int ngl, int nq) { return 2 * ngl * nq + ngl * ngl + nq + ngl; }

__global__ void k_scale(double* x, double s, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = s * x[i];
}

// send[(h*nv + v)*ngl + n] = trace plane v at the local slot of halo face h
__global__ void k_pack_traces(const double* tr, size_t trstride, const int* halo_slot, int nhalo, int nv, int ngl, double* send) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t tot = (size_t)nhalo * nv * ngl;
    if (t >= tot) return;
    int n = t % ngl; size_t r = t / ngl; int v = r % nv; int h = r / nv;
    send[t] = tr[(size_t)v * trstride + (size_t)halo_slot[h] * ngl + n];
}
__global__ void k_unpack_traces(double* tr, size_t trstride, int nslots, int nhalo, int nv, int ngl, const double* recv) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t tot = (size_t)nhalo * nv * ngl;
    if (t >= tot) return;
    int n = t % ngl; size_t r = t / ngl; int v = r % nv; int h = r / nv;
    tr[(size_t)v * trstride + ((size_t)nslots + h) * ngl + n] = recv[t];
}
// trace records [slot][tside] (stage_pair.cuh): gather the records of the processor-face slots in exchange order; the
// receive side needs no unpack -- the halo region of the trace buffer is the receive buffer
__global__ void k_pack_trace_records(const double* tr, const int* halo_slot, int nhalo, int tside, double* send) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nhalo * tside) return;
    size_t h = t / tside; int j = (int)(t - h * tside);
    send[t] = tr[(size_t)halo_slot[h] * tside + j];
}
__global__ void k_pack_nodal(const double* planes, size_t stride, const int* halo_slot, int nhalo, int np, int ngl, int npts, double* send) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t tot = (size_t)nhalo * np * ngl;
    if (t >= tot) return;
    int n = t % ngl; size_t r = t / ngl; int p = r % np; int h = r / np;
    int slot = halo_slot[h], e = slot >> 2, s = slot & 3;
    send[t] = planes[(size_t)p * stride + (size_t)e * npts + face_node(s, n, ngl)];
}
__global__ void k_unpack_nodal(double* hout, size_t hstride, int nhalo, int np, int ngl, const double* recv) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t tot = (size_t)nhalo * np * ngl;
    if (t >= tot) return;
    int n = t % ngl; size_t r = t / ngl; int p = r % np; int h = r / np;
    hout[(size_t)p * hstride + (size_t)h * ngl + n] = recv[t];
}

// ---- NCCL through dlopen ---------------------------------------------------------------------------------------
struct Id128 { char internal[128]; };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, /*ncclUniqueId by value*/ Id128, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
    if (g_nccl.lib) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
    if (!g_nccl.lib) { set_error("NCCL", "libnccl.so.2 not found"); return -1; }
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.GroupStart = (int (*)())dlsym(g_nccl.lib, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(g_nccl.lib, "ncclGroupEnd");
    g_nccl.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclSend");
    g_nccl.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclRecv");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart || !g_nccl.GroupEnd) {
        set_error("NCCL", "missing symbols in libnccl"); return -1;
    }
    return 0;
}
static const int NCCL_DOUBLE = 8;  // ncclFloat64

// ---- local (in-process) back end -------------------------------------------------------------------------------
// Host barrier with an abort flag: a rank that fails (anywhere in its step) aborts the group, which releases every
// waiter with an error instead of leaving the peers blocked for ever.
struct LocalGroup {
    int nranks = 0, members = 0;
    std::vector<Solver*> peers;
    std::mutex m;
    std::condition_variable cv;
    int waiting = 0;
    unsigned long generation = 0;
    bool aborted = false;
    int wait() {   // 0 ok, -1 group aborted
        std::unique_lock<std::mutex> lk(m);
        if (aborted) return -1;
        const unsigned long gen = generation;
        if (++waiting == nranks) { waiting = 0; ++generation; cv.notify_all(); return 0; }
        cv.wait(lk, [&] { return generation != gen || aborted; });
        return aborted ? -1 : 0;
    }
    void abort() { std::lock_guard<std::mutex> lk(m); aborted = true; cv.notify_all(); }
};
static std::mutex g_local_mutex;
static std::map<int, LocalGroup*> g_local_groups;
static std::map<LocalGroup*, int> g_local_ids;

inline int halo_get_unique_id(void* id128) {
    if (nccl_load()) return -1;
    int rc = g_nccl.GetUniqueId(id128);
    if (rc) { set_error("ncclGetUniqueId", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"); return -1; }
    return 0;
}

// id128 == "LOCAL" + int group id in bytes 8..11 selects the in-process back end
inline int halo_comm_init(Solver& S, const void* id128) {
    const char* c = (const char*)id128;
    if (!memcmp(c, "LOCAL", 5)) {
        int gid; memcpy(&gid, c + 8, sizeof(int));
        std::lock_guard<std::mutex> lk(g_local_mutex);
        LocalGroup*& G = g_local_groups[gid];
        if (G && G->nranks != S.desc.nranks) { set_error("halo", "local group id already in use with another number of ranks"); return -1; }
        if (!G) { G = new LocalGroup(); G->nranks = S.desc.nranks; G->peers.assign(S.desc.nranks, nullptr); g_local_ids[G] = gid; }
        if (S.desc.rank < 0 || S.desc.rank >= G->nranks || G->peers[S.desc.rank]) { set_error("halo", "rank out of range or already registered in the local group"); return -1; }
        G->peers[S.desc.rank] = &S; G->members++;
        S.local_group = G;
        return 0;
    }
    if (nccl_load()) return -1;
    Id128 id; memcpy(&id, id128, sizeof(id));
    cudaSetDevice(S.device);
    int rc = g_nccl.CommInitRank(&S.nccl_comm, S.desc.nranks, id, S.desc.rank);
    if (rc) { set_error("ncclCommInitRank", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"); return -1; }
    return 0;
}
inline void halo_comm_destroy(Solver& S) {
    if (S.nccl_comm && g_nccl.CommDestroy) { g_nccl.CommDestroy(S.nccl_comm); S.nccl_comm = nullptr; }
    if (S.local_group) {
        std::lock_guard<std::mutex> lk(g_local_mutex);
        LocalGroup* G = (LocalGroup*)S.local_group;
        G->abort();   // a peer still stepping must not wait for a rank that is gone
        if (S.desc.rank >= 0 && S.desc.rank < G->nranks && G->peers[S.desc.rank] == &S) { G->peers[S.desc.rank] = nullptr; G->members--; }
        if (G->members == 0) { g_local_groups.erase(g_local_ids[G]); g_local_ids.erase(G); delete G; }
        S.local_group = nullptr;
    }
}
// a failing rank releases its peers (no-op for NCCL: a failed NCCL call is fatal for the communicator anyway)
inline void halo_abort(Solver& S) { if (S.local_group) ((LocalGroup*)S.local_group)->abort(); }

// move S.d_send -> neighbours' S.d_recv ; `per_face` doubles per halo face
static int halo_sendrecv(Solver& S, size_t per_face, double* recv_base = nullptr, cudaStream_t st = nullptr) {
    if (!recv_base) recv_base = S.d_recv;
    if (!st) st = S.stream;
    if ((size_t)S.nhalo * per_face > S.halo_capacity) { set_error("halo", "staging buffer too small"); return -1; }
    if (S.nccl_comm) {
        int rc = g_nccl.GroupStart();
        for (size_t i = 0; i < S.nbh_rank.size() && !rc; ++i) {
            size_t off = (size_t)S.nbh_offset[i] * per_face, cnt = (size_t)S.nbh_count[i] * per_face;
            rc = g_nccl.Send(S.d_send + off, cnt, NCCL_DOUBLE, S.nbh_rank[i], S.nccl_comm, st);
            if (!rc) rc = g_nccl.Recv(recv_base + off, cnt, NCCL_DOUBLE, S.nbh_rank[i], S.nccl_comm, st);
        }
        int rc2 = g_nccl.GroupEnd();
        if (rc || rc2) { set_error("ncclSend/Recv", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc ? rc : rc2) : "error"); return -1; }
        return 0;
    }
    if (S.local_group) {
        LocalGroup* G = (LocalGroup*)S.local_group;
        int err = 0;
        if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("halo", "stream sync failed"); err = 1; }
        if (err) G->abort();
        if (G->wait()) { if (!err) set_error("halo", "a peer rank of the local group failed"); return -1; }   // every rank's send buffer is complete
        for (size_t i = 0; i < S.nbh_rank.size() && !err; ++i) {
            Solver* P = G->peers[S.nbh_rank[i]];
            // find my block in the peer's neighbour list: the peer lists me with the same number of faces in the same order
            size_t poff = 0; bool found = false;
            for (size_t j = 0; P && j < P->nbh_rank.size(); ++j)
                if (P->nbh_rank[j] == S.desc.rank) { poff = (size_t)P->nbh_offset[j] * per_face; found = true; break; }
            if (!found) { set_error("halo", "asymmetric neighbour lists"); err = 1; break; }
            size_t off = (size_t)S.nbh_offset[i] * per_face, cnt = (size_t)S.nbh_count[i] * per_face;
            if (cudaMemcpyAsync(recv_base + off, P->d_send + poff, cnt * sizeof(double), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
                set_error("halo", "peer copy failed"); err = 1;
            }
        }
        if (!err && cudaStreamSynchronize(st) != cudaSuccess) { set_error("halo", "copy failed"); err = 1; }
        if (err) G->abort();
        if (G->wait()) { if (!err) set_error("halo", "a peer rank of the local group failed"); return -1; }   // peers may overwrite their send buffers now
        return 0;
    }
    set_error("halo", "partition has processor faces but no communicator (call hnumo_comm_init)");
    return -1;
}

inline int halo_exchange_traces(Solver& S, const Planes& tr, int nv) {
    if (S.nhalo == 0) return 0;
    size_t tot = (size_t)S.nhalo * nv * S.ngl;
    k_pack_traces<<<(tot + 255) / 256, 256, 0, S.stream>>>(tr.p, tr.stride, S.d_halo_slot, S.nhalo, nv, S.ngl, S.d_send);
    if (halo_sendrecv(S, (size_t)nv * S.ngl)) return -1;
    k_unpack_traces<<<(tot + 255) / 256, 256, 0, S.stream>>>(tr.p, tr.stride, S.nslots, S.nhalo, nv, S.ngl, S.d_recv);
    S.n_launches += 2;
    return 0;
}
// slot planes of any width (width = nq: face values at the face quadrature points, method_visc == 1; the reference's
// create_communicator_quad / bcl_create_communicator(.,4,nlayers,nq), src/create_rhs_communicator.F90:82-134)
inline int halo_exchange_slot_planes(Solver& S, double* p, size_t stride, int nv, int width) {
    if (S.nhalo == 0) return 0;
    size_t tot = (size_t)S.nhalo * nv * width;
    k_pack_traces<<<(tot + 255) / 256, 256, 0, S.stream>>>(p, stride, S.d_halo_slot, S.nhalo, nv, width, S.d_send);
    if (halo_sendrecv(S, (size_t)nv * width)) return -1;
    k_unpack_traces<<<(tot + 255) / 256, 256, 0, S.stream>>>(p, stride, S.nslots, S.nhalo, nv, width, S.d_recv);
    S.n_launches += 2;
    return 0;
}
// several sets of nodal planes in ONE message per neighbour: send[(h*NPT + p)*ngl + n] with p running over the planes of all
// segments (NPT in total)
struct HaloSeg { const double* planes; int np; size_t stride; double* hout; size_t hstride; };
struct HaloSegs { HaloSeg s[4]; int n, npt; };
__global__ void k_pack_nodal_multi(HaloSegs g, const int* halo_slot, int nhalo, int ngl, int npts, double* send) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t tot = (size_t)nhalo * g.npt * ngl;
    if (t >= tot) return;
    int n = t % ngl; size_t r = t / ngl; int p = r % g.npt; int h = r / g.npt;
    int slot = halo_slot[h], e = slot >> 2, s = slot & 3, k = 0;
    while (p >= g.s[k].np) { p -= g.s[k].np; ++k; }
    send[t] = g.s[k].planes[(size_t)p * g.s[k].stride + (size_t)e * npts + face_node(s, n, ngl)];
}
__global__ void k_unpack_nodal_multi(HaloSegs g, int nhalo, int ngl, const double* recv) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t tot = (size_t)nhalo * g.npt * ngl;
    if (t >= tot) return;
    int n = t % ngl; size_t r = t / ngl; int p = r % g.npt; int h = r / g.npt, k = 0;
    while (p >= g.s[k].np) { p -= g.s[k].np; ++k; }
    g.s[k].hout[(size_t)p * g.s[k].hstride + (size_t)h * ngl + n] = recv[t];
}
inline int halo_exchange_nodal_multi(Solver& S, const HaloSeg* seg, int nseg) {
    if (S.nhalo == 0) return 0;
    HaloSegs g; g.n = nseg; g.npt = 0;
    for (int i = 0; i < nseg; ++i) { g.s[i] = seg[i]; g.npt += seg[i].np; }
    for (int i = nseg; i < 4; ++i) g.s[i] = HaloSeg{nullptr, 1 << 30, 0, nullptr, 0};
    size_t tot = (size_t)S.nhalo * g.npt * S.ngl;
    k_pack_nodal_multi<<<(tot + 255) / 256, 256, 0, S.stream>>>(g, S.d_halo_slot, S.nhalo, S.ngl, S.npts, S.d_send);
    if (halo_sendrecv(S, (size_t)g.npt * S.ngl)) return -1;
    k_unpack_nodal_multi<<<(tot + 255) / 256, 256, 0, S.stream>>>(g, S.nhalo, S.ngl, S.d_recv);
    S.n_launches += 2;
    return 0;
}
inline int halo_exchange_nodal(Solver& S, const double* planes, int np, size_t stride, const Planes& hout) {
    if (S.nhalo == 0) return 0;
    size_t tot = (size_t)S.nhalo * np * S.ngl;
    k_pack_nodal<<<(tot + 255) / 256, 256, 0, S.stream>>>(planes, stride, S.d_halo_slot, S.nhalo, np, S.ngl, S.npts, S.d_send);
    if (halo_sendrecv(S, (size_t)np * S.ngl)) return -1;
    k_unpack_nodal<<<(tot + 255) / 256, 256, 0, S.stream>>>(hout.p, hout.stride, S.nhalo, np, S.ngl, S.d_recv);
    S.n_launches += 2;
    return 0;
}

}  // namespace hn
