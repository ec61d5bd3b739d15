// ============================================================================
// oracle_visc_q.cpp -- TEST INFRASTRUCTURE ONLY (CPU oracle).
// Quadrature-point LDG viscosity, method_visc == 1:
//   interpolate_dpp                    src/mod_layer_terms.F90:25-55
//   compute_gradient_uv_q              src/mod_barotropic_terms.F90:445-477
//   btp_create_laplacian_v2            src/mod_laplacian_quad.F90:125-223
//   bcl_create_laplacian_v2            src/mod_laplacian_quad.F90:252-355
//   compute_laplacian_quad             src/mod_laplacian_quad.F90:613-642
//   create_rhs_laplacian_flux_quad     src/mod_laplacian_quad.F90:644-722
// Statement by statement in the reference's loop structure; the face loops are threaded over elements with the
// per-element face lists of build_elem_faces (same additions in the same order: bit-identical to the serial face loop).
// ============================================================================
#include <algorithm>
#include <cmath>

#include "hnumo_oracle.hpp"

namespace orc {

// imapl_q / imapr_q (create_normals_quad.F90:227-355) on a conforming brick: quadrature point l of local face iloc
static inline int face_quad_index(int nq, int iloc, int el0, int l) {
    int iq, jq;
    switch (iloc) {
        case 3: iq = l; jq = 0; break;
        case 4: iq = l; jq = nq - 1; break;
        case 5: iq = 0; jq = l; break;
        default: iq = nq - 1; jq = l; break;
    }
    return el0 * nq * nq + jq * nq + iq;
}

// mod_layer_terms.F90:25-55
void Oracle::interpolate_dpp() {
    dpprime_visc_q.zero();
#pragma omp parallel for schedule(static)
    for (int Iq = 0; Iq < npoin_q; ++Iq)
        for (int ip = 0; ip < npts; ++ip) {
            int I = indexq[(size_t)Iq * npts + ip];
            double hi = psih(ip, Iq);
            for (int k = 0; k < nl; ++k) dpprime_visc_q(Iq, k) = dpprime_visc_q(Iq, k) + dpprime_visc(I, k) * hi;
        }
}

// mod_barotropic_terms.F90:445-477: grad_uv(2,2,npoin_q) stored as (4,npoin_q): (1,1) (1,2) (2,1) (2,2) -> 0 1 2 3
void Oracle::compute_gradient_uv_q(Arr& grad_uv, const Arr& uv) {
    grad_uv.zero();
#pragma omp parallel for schedule(static)
    for (int Iq = 0; Iq < npoin_q; ++Iq)
        for (int ip = 0; ip < npts; ++ip) {
            int I = indexq[(size_t)Iq * npts + ip];
            double dhdx = dpsidx(ip, Iq), dhdy = dpsidy(ip, Iq);
            grad_uv(0, Iq) = grad_uv(0, Iq) + dhdx * uv(0, I);
            grad_uv(1, Iq) = grad_uv(1, Iq) + dhdy * uv(0, I);
            grad_uv(2, Iq) = grad_uv(2, Iq) + dhdx * uv(1, I);
            grad_uv(3, Iq) = grad_uv(3, Iq) + dhdy * uv(1, I);
        }
}

// face values of a quadrature-point flux (mod_laplacian_quad.F90:161-208 / 283-333): left = the left element's value at the
// face quadrature point, right = the neighbour's, or the ghost (copy; mirrored about the normal on free-slip walls).
// Processor boundaries (er == 0) keep the copy, which the halo exchange overwrites in the reference (single rank here).
void Oracle::visc_flux_faces(Arr& ff, const Arr& flux) {
#pragma omp parallel for schedule(static)
    for (int f = 0; f < nface; ++f) {
        int ilocl = face[8 * f + 4], ilocr = face[8 * f + 5], iel = face[8 * f + 6] - 1, ier = face[8 * f + 7];
        for (int iquad = 0; iquad < nq; ++iquad) {
            int Iq = face_quad_index(nq, ilocl, iel, iquad);
            for (int v = 0; v < 4; ++v) ff(v, 0, iquad, f) = flux(v, Iq);
            if (ier > 0) {
                int Ir = face_quad_index(nq, ilocr, ier - 1, iquad);
                for (int v = 0; v < 4; ++v) ff(v, 1, iquad, f) = flux(v, Ir);
            } else {
                for (int v = 0; v < 4; ++v) ff(v, 1, iquad, f) = ff(v, 0, iquad, f);
                if (ier == -4) {
                    double nx = normal_vector_q(0, iquad, f), ny = normal_vector_q(1, iquad, f);
                    double un = flux(0, Iq) * nx + flux(1, Iq) * ny;
                    ff(0, 1, iquad, f) = flux(0, Iq) - 2.0 * un * nx;
                    ff(1, 1, iquad, f) = flux(1, Iq) - 2.0 * un * ny;
                    un = flux(2, Iq) * nx + flux(3, Iq) * ny;
                    ff(2, 1, iquad, f) = flux(2, Iq) - 2.0 * un * nx;
                    ff(3, 1, iquad, f) = flux(3, Iq) - 2.0 * un * ny;
                }
            }
        }
    }
}

// mod_laplacian_quad.F90:613-642
void Oracle::compute_laplacian_quad(Arr& lap_q, const Arr& grad_dpuvp) {
    lap_q.zero();
    const int nq2 = nq * nq;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < nelem; ++e)
        for (int Iq = e * nq2; Iq < (e + 1) * nq2; ++Iq) {
            double wq = wjac(Iq);
            for (int ip = 0; ip < npts; ++ip) {
                int I = indexq[(size_t)Iq * npts + ip];
                double u_visc = dpsidx(ip, Iq) * grad_dpuvp(0, Iq) + dpsidy(ip, Iq) * grad_dpuvp(1, Iq);
                double v_visc = dpsidx(ip, Iq) * grad_dpuvp(2, Iq) + dpsidy(ip, Iq) * grad_dpuvp(3, Iq);
                lap_q(0, I) = lap_q(0, I) - wq * u_visc;
                lap_q(1, I) = lap_q(1, I) - wq * v_visc;
            }
        }
}

// mod_laplacian_quad.F90:644-722 (beta = 0.5: central flux; the flux expression is kept as written)
void Oracle::create_rhs_laplacian_flux_quad(Arr& rhs, const Arr& gradq_face) {
    if (ef_ptr.empty()) build_elem_faces();
    const double beta = 0.5, alpha = 1.0 - beta;
    Arr fl; fl.alloc(2, nq, nface);   // wq * (flux_qu, flux_qv) per face quadrature point
#pragma omp parallel for schedule(static)
    for (int f = 0; f < nface; ++f)
        for (int iquad = 0; iquad < nq; ++iquad) {
            double qul[2] = {gradq_face(0, 0, iquad, f), gradq_face(1, 0, iquad, f)}, qvl[2] = {gradq_face(2, 0, iquad, f), gradq_face(3, 0, iquad, f)};
            double qur[2] = {gradq_face(0, 1, iquad, f), gradq_face(1, 1, iquad, f)}, qvr[2] = {gradq_face(2, 1, iquad, f), gradq_face(3, 1, iquad, f)};
            double qu_mean[2] = {alpha * qul[0] + beta * qur[0], alpha * qul[1] + beta * qur[1]};
            double qv_mean[2] = {alpha * qvl[0] + beta * qvr[0], alpha * qvl[1] + beta * qvr[1]};
            double nx = normal_vector_q(0, iquad, f), ny = normal_vector_q(1, iquad, f);
            fl(0, iquad, f) = (qu_mean[0] - qul[0] * nx) + (qu_mean[1] - qul[1] * ny);
            fl(1, iquad, f) = (qv_mean[0] - qvl[0] * nx) + (qv_mean[1] - qvl[1] * ny);
        }
#pragma omp parallel for schedule(static)
    for (int e = 0; e < nelem; ++e)
        for (int p = ef_ptr[e]; p < ef_ptr[e + 1]; ++p) {
            const int f = ef_face[p], side = ef_side[p];
            const std::vector<int>& fn = side == 0 ? fnodeL : fnodeR;
            const double sgn = side == 0 ? 1.0 : -1.0;
            for (int iquad = 0; iquad < nq; ++iquad) {
                double wq = jac_faceq(iquad, f);
                for (int i = 0; i < ngl; ++i) {
                    double hi = psiq(i, iquad);
                    int ip = fn[(size_t)f * ngl + i];
                    rhs(0, ip) = rhs(0, ip) + sgn * (wq * hi * fl(0, iquad, f));
                    rhs(1, ip) = rhs(1, ip) + sgn * (wq * hi * fl(1, iquad, f));
                }
            }
        }
}

// mod_laplacian_quad.F90:125-223
void Oracle::btp_create_laplacian_v2(Arr& rhs_lap, const Arr& qprime, const Arr& qb) {
    Arr Uk; Uk.alloc(2, npoin);
    Arr flux_uv_visc; flux_uv_visc.alloc(4, npoin_q);
    Arr graduv; graduv.alloc(4, npoin_q);
    Arr ff; ff.alloc(4, 2, nq, nface);
    Arr rhs_temp; rhs_temp.alloc(2, npoin);
    for (int k = 0; k < nl; ++k) {
#pragma omp parallel for schedule(static)
        for (int I = 0; I < npoin; ++I) {
            Uk(0, I) = qprime(1, I, k) + qb(2, I) / qb(0, I);
            Uk(1, I) = qprime(2, I, k) + qb(3, I) / qb(0, I);
        }
        compute_gradient_uv_q(graduv, Uk);
#pragma omp parallel for schedule(static)
        for (int Iq = 0; Iq < npoin_q; ++Iq)
            for (int v = 0; v < 4; ++v) flux_uv_visc(v, Iq) = flux_uv_visc(v, Iq) + dpprime_visc_q(Iq, k) * graduv(v, Iq);
    }
    visc_flux_faces(ff, flux_uv_visc);
    compute_laplacian_quad(rhs_temp, flux_uv_visc);
    create_rhs_laplacian_flux_quad(rhs_temp, ff);
#pragma omp parallel for schedule(static)
    for (int I = 0; I < npoin; ++I) {
        rhs_lap(0, I) = cfg.visc_mlswe * massinv(I) * rhs_temp(0, I);
        rhs_lap(1, I) = cfg.visc_mlswe * massinv(I) * rhs_temp(1, I);
    }
}

// mod_laplacian_quad.F90:252-355
void Oracle::bcl_create_laplacian_v2(Arr& rhs_lap, const Arr& qprime) {
    rhs_lap.zero();
    Arr Uk; Uk.alloc(2, npoin);
    Arr flux; flux.alloc(4, npoin_q);
    Arr graduv; graduv.alloc(4, npoin_q);
    Arr ff; ff.alloc(4, 2, nq, nface);
    Arr rhs_temp; rhs_temp.alloc(2, npoin);
    for (int k = 0; k < nl; ++k) {
#pragma omp parallel for schedule(static)
        for (int I = 0; I < npoin; ++I) {
            Uk(0, I) = qprime(1, I, k) + uvb_ave_df(0, I);
            Uk(1, I) = qprime(2, I, k) + uvb_ave_df(1, I);
        }
        compute_gradient_uv_q(graduv, Uk);
#pragma omp parallel for schedule(static)
        for (int Iq = 0; Iq < npoin_q; ++Iq)
            for (int v = 0; v < 4; ++v) flux(v, Iq) = dpprime_visc_q(Iq, k) * graduv(v, Iq);
        visc_flux_faces(ff, flux);
        compute_laplacian_quad(rhs_temp, flux);
        create_rhs_laplacian_flux_quad(rhs_temp, ff);
        for (int I = 0; I < npoin; ++I) {
            rhs_lap(0, I, k) = cfg.visc_mlswe * massinv(I) * rhs_temp(0, I);
            rhs_lap(1, I, k) = cfg.visc_mlswe * massinv(I) * rhs_temp(1, I);
        }
    }
}

}  // namespace orc
