// ============================================================================
// oracle_capi.cpp -- TEST INFRASTRUCTURE ONLY.  C entry points of the CPU oracle
// (loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs; never by the product).
// ============================================================================
#include <cstring>
#include <omp.h>

#include "hnumo_oracle.hpp"

using orc::Arr;
using orc::Oracle;

extern "C" {

// OpenMP team size of every later call (the launcher's OMP_NUM_THREADS is only the default)
void orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int orc_get_max_threads(void) { return omp_get_max_threads(); }
void* orc_create(const orc::Config* cfg) { return new Oracle(*cfg); }
void orc_destroy(void* h) { delete (Oracle*)h; }

// out[0..9] = ngl,nq,nelem,npoin,npoin_q,nface,nlayers,N_btp,kstages,error_flag
void orc_info(void* h, int* out) {
    Oracle* o = (Oracle*)h;
    out[0] = o->ngl; out[1] = o->nq; out[2] = o->nelem; out[3] = o->npoin; out[4] = o->npoin_q;
    out[5] = o->nface; out[6] = o->nl; out[7] = o->N_btp; out[8] = o->kstages; out[9] = o->error_flag;
}
// out[0..5] = dt, dt_btp, gravity, btp_seconds, stage_count, unused
void orc_scalars(void* h, double* out) {
    Oracle* o = (Oracle*)h;
    out[0] = o->dt; out[1] = o->dt_btp; out[2] = o->gravity; out[3] = o->btp_seconds; out[4] = (double)o->stage_count;
}
long orc_array_size(void* h, const char* name) {
    Oracle* o = (Oracle*)h;
    auto it = o->reg.find(name);
    if (it == o->reg.end()) {
        if (!std::strcmp(name, "wgl")) return o->ngl;
        if (!std::strcmp(name, "wnq")) return o->nq;
        if (!std::strcmp(name, "xgl")) return o->ngl;
        if (!std::strcmp(name, "xnq")) return o->nq;
        return -1;
    }
    return (long)it->second->size();
}
int orc_get_array(void* h, const char* name, double* out) {
    Oracle* o = (Oracle*)h;
    auto it = o->reg.find(name);
    if (it == o->reg.end()) {
        const std::vector<double>* v = nullptr;
        if (!std::strcmp(name, "wgl")) v = &o->wgl;
        else if (!std::strcmp(name, "wnq")) v = &o->wnq;
        else if (!std::strcmp(name, "xgl")) v = &o->xgl;
        else if (!std::strcmp(name, "xnq")) v = &o->xnq;
        if (!v) return -1;
        std::memcpy(out, v->data(), v->size() * sizeof(double));
        return 0;
    }
    std::memcpy(out, it->second->data(), it->second->size() * sizeof(double));
    return 0;
}
int orc_set_array(void* h, const char* name, const double* in) {
    Oracle* o = (Oracle*)h;
    auto it = o->reg.find(name);
    if (it == o->reg.end()) return -1;
    std::memcpy(it->second->data(), in, it->second->size() * sizeof(double));
    return 0;
}
void orc_get_face(void* h, int* out) {
    Oracle* o = (Oracle*)h;
    std::memcpy(out, o->face.data(), o->face.size() * sizeof(int));
}
// one or more baroclinic steps (ti_rk_bcl.F90); returns error flag
int orc_step(void* h, int nsteps) {
    Oracle* o = (Oracle*)h;
    for (int i = 0; i < nsteps; ++i) o->ti_rk_bcl();
    return o->error_flag;
}
// phase: face traces + dpprime_visc + btp_bcl_coeffs_qdf from the current qprime_df (ti_rk_bcl.F90:43-50)
void orc_btp_bcl_coeffs(void* h) {
    Oracle* o = (Oracle*)h;
    Arr qf; qf.alloc(3, 2, o->ngl, o->nface, o->nl);
    o->extract_qprime_df_face(qf, o->qprime_df);
    for (int k = 0; k < o->nl; ++k)
        for (int I = 0; I < o->npoin; ++I) o->dpprime_visc(I, k) = o->qprime_df(0, I, k);
    if (o->cfg.method_visc == 1) o->interpolate_dpp();
    o->btp_bcl_coeffs_qdf(qf, o->qprime_df);
}
// phase: one barotropic RHS evaluation on the current qb_df (mod_rhs_btp.F90:28-59); accumulators are updated
void orc_rhs_btp(void* h, double* rhs_out) {
    Oracle* o = (Oracle*)h;
    Arr rhs; rhs.alloc(3, o->npoin);
    o->create_rhs_btp(rhs, o->qb_df, o->qprime_df);
    std::memcpy(rhs_out, rhs.data(), rhs.size() * sizeof(double));
}
// phase: layer_mass_rhs (mod_create_rhs_mlswe.F90:53-78) of the current qprime_df with the current time averages: dp_advec(npoin,nl)
void orc_layer_mass_rhs(void* h, double* out) {
    Oracle* o = (Oracle*)h;
    Arr qf; qf.alloc(3, 2, o->ngl, o->nface, o->nl);
    o->extract_qprime_df_face(qf, o->qprime_df);
    Arr dp_advec; dp_advec.alloc(o->npoin, o->nl);
    o->layer_mass_rhs(dp_advec, o->qprime_df, qf);
    std::memcpy(out, dp_advec.data(), dp_advec.size() * sizeof(double));
}
// phase: layer_momentum_rhs (mod_create_rhs_mlswe.F90:28-51) of the current q_df, qprime_df: rhs_mom(2,npoin,nl)
void orc_layer_momentum_rhs(void* h, double* out) {
    Oracle* o = (Oracle*)h;
    Arr qf; qf.alloc(3, 2, o->ngl, o->nface, o->nl);
    o->extract_qprime_df_face(qf, o->qprime_df);
    Arr rhs; rhs.alloc(2, o->npoin, o->nl);
    o->rhs_momentum(rhs, o->qprime_df, o->q_df, qf);
    std::memcpy(out, rhs.data(), rhs.size() * sizeof(double));
}
// phase: rhs_layer_shear_stress of the current q_df (mod_create_rhs_mlswe.F90:146-279): rhs_stress(2,npoin,nl), no mass matrix
void orc_shear_stress(void* h, double* out) {
    Oracle* o = (Oracle*)h;
    Arr rs; rs.alloc(2, o->npoin, o->nl);
    o->rhs_layer_shear_stress(rs, o->q_df);
    std::memcpy(out, rs.data(), rs.size() * sizeof(double));
}
// phase: full barotropic substep loop on qb_df in place (mod_rk_mlswe.F90:19-151)
void orc_btp_substeps(void* h) {
    Oracle* o = (Oracle*)h;
    o->ti_barotropic_ssprk_mlswe(o->qb_df, o->qprime_df);
}
// diagnostics.F90:24-45 -> qout(5,npoin,nl); masses(nl) per compute_conserved.F90
void orc_diagnostics(void* h, double* qout, double* masses) {
    Oracle* o = (Oracle*)h;
    Arr q; o->diagnostics(q);
    std::memcpy(qout, q.data(), q.size() * sizeof(double));
    for (int k = 0; k < o->nl; ++k) masses[k] = o->layer_mass(q, k);
}
}
