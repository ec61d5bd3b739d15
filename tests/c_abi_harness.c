/* c_abi_harness.c -- a second, non-Python caller of the C-ABI (include/hnumo_b200.h).
 *
 * Plain C, linked against libhnumo_b200.so like the Fortran driver would be.  It fills hnumo_desc_t from column-major arrays
 * read from a deck file (written by tests/harness_util.py from the same brick deck the ctypes tests use), runs
 *     hnumo_init -> nsteps x hnumo_ti_rk_bcl(q_df, qb_df, qprime_df) -> hnumo_diagnostics -> hnumo_finalize
 * on host buffers it owns, and writes the final state.  tests/test_gpu_harness.py compares that state bitwise with the
 * ctypes path.  Built with gcc by __graft_entry__.build() (no CUDA headers needed: the ABI is plain C).
 *
 * usage: c_abi_harness <deck.bin> <out.bin> <nsteps>      exit code: 0 ok, 2 usage/IO, 3 hnumo_init failed (e.g. no GPU),
 *                                                          4 a step failed, 10+rc physics error code rc
 * deck file: int32 header[16] = {magic 0x484e554d, nelem, ngl, nq, nlayers, nface, kstages, N_btp, botfr, method_visc, rank,
 *            nranks, num_nbh, nsr (sum of num_send_recv), 0, 0}; double scalars[6] = {dt, dt_btp, gravity, cd, visc, ad};
 *            then the arrays in the order of hnumo_desc_t, then q_df, qb_df, qprime_df. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/hnumo_b200.h"

static void* rd(FILE* f, size_t bytes) {
    void* p = malloc(bytes ? bytes : 1);
    if (!p || (bytes && fread(p, 1, bytes, f) != bytes)) { fprintf(stderr, "harness: short read\n"); exit(2); }
    return p;
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s deck.bin out.bin nsteps\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("deck"); return 2; }
    int32_t h[16];
    double sc[7];
    if (fread(h, sizeof(int32_t), 16, f) != 16 || h[0] != 0x484e554d || fread(sc, sizeof(double), 7, f) != 7) { fprintf(stderr, "harness: bad deck header\n"); return 2; }
    hnumo_desc_t d;
    memset(&d, 0, sizeof(d));
    d.abi_version = HNUMO_ABI_VERSION;
    d.nelem = h[1]; d.ngl = h[2]; d.nq = h[3]; d.nlayers = h[4]; d.nface = h[5]; d.kstages = h[6]; d.N_btp = h[7];
    d.botfr = h[8]; d.method_visc = h[9]; d.rank = h[10]; d.nranks = h[11]; d.num_nbh = h[12];
    const int nsr = h[13];
    d.dt = sc[0]; d.dt_btp = sc[1]; d.gravity = sc[2]; d.cd_mlswe = sc[3]; d.visc_mlswe = sc[4]; d.ad_mlswe = sc[5]; d.max_shear_dz = sc[6];
    const size_t npoin = (size_t)d.nelem * d.ngl * d.ngl, D = sizeof(double), I = sizeof(int32_t);
    d.psiq = rd(f, D * d.ngl * d.nq); d.dpsiq = rd(f, D * d.ngl * d.nq); d.wnq = rd(f, D * d.nq); d.wgl = rd(f, D * d.ngl);
    d.dpsi = rd(f, D * d.ngl * d.ngl);
    d.face = rd(f, I * 8 * d.nface); d.elem_metrics = rd(f, D * 5 * d.nelem); d.face_geom = rd(f, D * 3 * d.nface);
    d.pbprime_df = rd(f, D * npoin); d.massinv = rd(f, D * npoin); d.coriolis_df = rd(f, D * npoin); d.tau_wind_df = rd(f, D * 2 * npoin);
    d.zbot_df = rd(f, D * npoin); d.alpha_mlswe = rd(f, D * d.nlayers); d.ssprk_a = rd(f, D * d.kstages * 3); d.ssprk_beta = rd(f, D * d.kstages);
    if (d.num_nbh > 0) { d.nbh_proc = rd(f, I * d.num_nbh); d.num_send_recv = rd(f, I * d.num_nbh); d.nbh_send_recv = rd(f, I * nsr); }
    double* q = rd(f, D * 3 * npoin * d.nlayers);
    double* qb = rd(f, D * 4 * npoin);
    double* qp = rd(f, D * 3 * npoin * d.nlayers);
    fclose(f);

    hnumo_handle_t H = NULL;
    int rc = hnumo_init(&d, &H);
    if (rc != 0) { fprintf(stderr, "harness: hnumo_init failed (%d): %s\n", rc, hnumo_last_error()); return 3; }
    const int nsteps = atoi(argv[3]);
    for (int s = 0; s < nsteps; ++s) {
        rc = hnumo_ti_rk_bcl(H, q, qb, qp);   /* the reference's call: call ti_rk_bcl(q_df, qb_df, qprime_df) */
        if (rc < 0) { fprintf(stderr, "harness: step failed (%d): %s\n", rc, hnumo_last_error()); return 4; }
        if (rc > 0) { fprintf(stderr, "harness: physics error %d: %s\n", rc, hnumo_last_error()); return 10 + rc; }
    }
    const int64_t nd = 11 * (int64_t)d.nlayers + 12;
    double* diag = malloc(D * nd);
    if (hnumo_upload_state(H, q, qb, qp) != 0 || hnumo_diagnostics(H, diag, nd) != nd) { fprintf(stderr, "harness: diagnostics failed: %s\n", hnumo_last_error()); return 4; }
    FILE* o = fopen(argv[2], "wb");
    if (!o) { perror("out"); return 2; }
    fwrite(q, D, 3 * npoin * d.nlayers, o); fwrite(qb, D, 4 * npoin, o); fwrite(qp, D, 3 * npoin * d.nlayers, o); fwrite(diag, D, nd, o);
    fclose(o);
    hnumo_finalize(H);
    printf("harness ok: %d steps, layer 1 mass %.17g\n", nsteps, diag[0]);
    return 0;
}
