// ============================================================================
// oracle_btp.cpp -- TEST INFRASTRUCTURE ONLY (CPU oracle, barotropic half).
// As-written restatement of the barotropic substep loop of h-NUMO:
//   ti_barotropic_ssprk_mlswe  src/mod_rk_mlswe.F90:19-151
//   create_rhs_btp             src/mod_rhs_btp.F90:28-59,102-370
//   btp_extract_df / btp_mom_boundary_df / btp_bcl_coeffs_qdf / compute_gradient_uv
//                              src/mod_barotropic_terms.F90:25-97,165-443
//   btp_create_laplacian       src/mod_laplacian_quad.F90:32-121,357-390,427-519
// Loop orders and operator precedence follow the reference.  OpenMP is used only
// where every output location is written by exactly one iteration.
// ============================================================================
#include <chrono>

#include "hnumo_oracle.hpp"

namespace orc {

// mod_barotropic_terms.F90:25-97
void Oracle::btp_extract_df(Arr& qb_df_face, const Arr& qb) {
    qb_df_face.zero();
#pragma omp parallel for schedule(static)
    for (int f = 0; f < nface; ++f) {
        int er = face[8 * f + 7];
        for (int n = 0; n < ngl; ++n) {
            int I = fnodeL[(size_t)f * ngl + n];
            for (int v = 0; v < 4; ++v) qb_df_face(v, 0, n, f) = qb(v, I);
            if (er > 0) {
                int Ir = fnodeR[(size_t)f * ngl + n];
                for (int v = 0; v < 4; ++v) qb_df_face(v, 1, n, f) = qb(v, Ir);
            } else {
                for (int v = 0; v < 4; ++v) qb_df_face(v, 1, n, f) = qb_df_face(v, 0, n, f);
                if (er == -4) {
                    double nx = normal_vector(0, n, f), ny = normal_vector(1, n, f);
                    double un = nx * qb_df_face(2, 0, n, f) + ny * qb_df_face(3, 0, n, f);
                    qb_df_face(2, 1, n, f) = qb_df_face(2, 0, n, f) - 2.0 * un * nx;
                    qb_df_face(3, 1, n, f) = qb_df_face(3, 0, n, f) - 2.0 * un * ny;
                } else if (er == -2) {
                    qb_df_face(2, 1, n, f) = -qb_df_face(2, 0, n, f);
                    qb_df_face(3, 1, n, f) = -qb_df_face(3, 0, n, f);
                }
            }
        }
    }
}

// mod_rhs_btp.F90:102-209
void Oracle::create_rhs_btp_volume_qdf(Arr& rhs, const Arr& qb, const Arr& qprime) {
    rhs.zero();
    const int botfr = cfg.botfr;
    const double cd_mlswe = cfg.cd_mlswe;
    const int nq2 = nq * nq;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < nelem; ++e) {
        double tb_u = 0.0, tb_v = 0.0;  // reference: set once before the loop, kept when botfr==0
        for (int Iq = e * nq2; Iq < (e + 1) * nq2; ++Iq) {
            double dp = 0, dpp = 0, udp = 0, vdp = 0, pp = 0, up = 0, vp = 0;
            for (int ip = 0; ip < npts; ++ip) {
                int I = indexq[(size_t)Iq * npts + ip];
                double hi = psih(ip, Iq);
                dp = dp + hi * qb(0, I);
                dpp = dpp + hi * qb(1, I);
                udp = udp + hi * qb(2, I);
                vdp = vdp + hi * qb(3, I);
                pp = pp + hi * qprime(0, I, nl - 1);
                up = up + hi * qprime(1, I, nl - 1);
                vp = vp + hi * qprime(2, I, nl - 1);
            }
            double wq = wjac(Iq);
            double ub = udp / dp, vb = vdp / dp;
            if (botfr == 1) {
                double ubot = up + ub, vbot = vp + vb;
                double spd = (cd_mlswe / gravity) * pp;
                tb_u = spd * ubot; tb_v = spd * vbot;
            } else if (botfr == 2) {
                double ubot = up + ub, vbot = vp + vb;
                double spd = (cd_mlswe / alpha_mlswe(nl - 1)) * std::sqrt(ubot * ubot + vbot * vbot);
                tb_u = spd * ubot; tb_v = spd * vbot;
            }
            double sc_x = coriolis_quad(Iq) * vdp + gravity * (tau_wind(0, Iq) - tb_u) - gravity * dp * grad_zbot_quad(0, Iq);
            double sc_y = -coriolis_quad(Iq) * udp + gravity * (tau_wind(1, Iq) - tb_v) - gravity * dp * grad_zbot_quad(1, Iq);
            double ope = 1.0 + dpp * one_over_pbprime(Iq);
            double Hq = (ope * ope) * H_bcl(Iq);
            double qu = ub * udp + ope * Q_uu_dp(Iq);
            double quv = ub * vdp + ope * Q_uv_dp(Iq);
            double qv = vb * vdp + ope * Q_vv_dp(Iq);

            H_ave(Iq) += Hq; Qu_ave(Iq) += qu; Qv_ave(Iq) += qv; Quv_ave(Iq) += quv;
            tau_bot_ave(0, Iq) += tb_u; tau_bot_ave(1, Iq) += tb_v;
            ope_ave(Iq) += ope; ope2_ave(Iq) += ope * ope;
            btp_mass_flux_ave(0, Iq) += udp; btp_mass_flux_ave(1, Iq) += vdp;
            uvb_ave(0, Iq) += ub; uvb_ave(1, Iq) += vb;

            for (int ip = 0; ip < npts; ++ip) {
                int I = indexq[(size_t)Iq * npts + ip];
                double hi = psih(ip, Iq), dhdx = dpsidx(ip, Iq), dhdy = dpsidy(ip, Iq);
                rhs(0, I) = rhs(0, I) + wq * (dhdx * udp + dhdy * vdp);
                rhs(1, I) = rhs(1, I) + wq * (hi * sc_x + dhdx * (Hq + qu) + quv * dhdy);
                rhs(2, I) = rhs(2, I) + wq * (hi * sc_y + dhdx * quv + dhdy * (Hq + qv));
            }
        }
    }
}

void Oracle::build_elem_faces() {
    ef_ptr.assign(nelem + 1, 0);
    for (int f = 0; f < nface; ++f) {
        ef_ptr[face[8 * f + 6]]++;                       // left element (1-based -> slot el+1)
        if (face[8 * f + 7] > 0) ef_ptr[face[8 * f + 7]]++;
    }
    for (int e = 0; e < nelem; ++e) ef_ptr[e + 1] += ef_ptr[e];
    ef_face.assign(ef_ptr[nelem], 0); ef_side.assign(ef_ptr[nelem], 0);
    std::vector<int> fill(ef_ptr.begin(), ef_ptr.end() - 1);
    for (int f = 0; f < nface; ++f) {                     // ascending f: every element's list ends up sorted by face number
        int el = face[8 * f + 6] - 1, er = face[8 * f + 7];
        ef_face[fill[el]] = f; ef_side[fill[el]++] = 0;
        if (er > 0) { ef_face[fill[er - 1]] = f; ef_side[fill[er - 1]++] = 1; }
    }
}

// mod_rhs_btp.F90:211-370.  Threaded as: (A) every face evaluated independently (its running sums are per face),
// (B) scatter of the reference gathered per element in the reference's face order (see ef_ptr in the header).
void Oracle::creat_btp_fluxes_qdf(Arr& rhs, const Arr& qb_df_face) {
    if (ef_ptr.empty()) build_elem_faces();
    static thread_local std::vector<double> quu, quv, qvu, qvv, H_face_temp, flux_edge_x, flux_edge_y, ul, ur, vl, vr, pbl, pbr, one_plus_eta_edge;
    Arr fl3; fl3.alloc(3, nq, nface);   // flux, H_kx + flux_x, H_ky + flux_y at the face quadrature points
#pragma omp parallel
    {
    quu.assign(nq, 0); quv.assign(nq, 0); qvu.assign(nq, 0); qvv.assign(nq, 0); H_face_temp.assign(nq, 0); flux_edge_x.assign(nq, 0);
    flux_edge_y.assign(nq, 0); ul.assign(nq, 0); ur.assign(nq, 0); vl.assign(nq, 0); vr.assign(nq, 0); pbl.assign(nq, 0); pbr.assign(nq, 0);
    one_plus_eta_edge.assign(nq, 0);
    Arr qbl, qbr; qbl.alloc(4, nq); qbr.alloc(4, nq);
#pragma omp for schedule(static)
    for (int f = 0; f < nface; ++f) {
        qbl.zero(); qbr.zero();
        std::fill(pbl.begin(), pbl.end(), 0.0); std::fill(pbr.begin(), pbr.end(), 0.0);
        for (int iq = 0; iq < nq; ++iq) {
            double nxl = normal_vector_q(0, iq, f), nyl = normal_vector_q(1, iq, f);
            double nxr = -nxl, nyr = -nyl;
            for (int n = 0; n < ngl; ++n) {
                double hi = psiq(n, iq);
                for (int v = 0; v < 4; ++v) {
                    qbl(v, iq) = qbl(v, iq) + hi * qb_df_face(v, 0, n, f);
                    qbr(v, iq) = qbr(v, iq) + hi * qb_df_face(v, 1, n, f);
                }
                pbl[iq] = pbl[iq] + hi * pbprime_df_face(0, n, f);
                pbr[iq] = pbr[iq] + hi * pbprime_df_face(1, n, f);
            }
            double pU_L = nxl * qbl(2, iq) + nyl * qbl(3, iq);
            double pU_R = nxr * qbr(2, iq) + nyr * qbr(3, iq);
            double pbpert_edge = coeff_pbpert_L(iq, f) * qbl(1, iq) + coeff_pbpert_R(iq, f) * qbr(1, iq) +
                                 coeff_pbub_LR(iq, f) * (pU_L + pU_R);
            one_plus_eta_edge[iq] = 1.0 + pbpert_edge * one_over_pbprime_edge(iq, f);
            flux_edge_x[iq] = coeff_mass_pbub_L(iq, f) * qbl(2, iq) + coeff_mass_pbub_R(iq, f) * qbr(2, iq) +
                              coeff_mass_pbpert_LR(iq, f) * (nxl * qbl(1, iq) + nxr * qbr(1, iq));
            flux_edge_y[iq] = coeff_mass_pbub_L(iq, f) * qbl(3, iq) + coeff_mass_pbub_R(iq, f) * qbr(3, iq) +
                              coeff_mass_pbpert_LR(iq, f) * (nyl * qbl(1, iq) + nyr * qbr(1, iq));
        }
        for (int iq = 0; iq < nq; ++iq) {
            ul[iq] = qbl(2, iq) / qbl(0, iq); ur[iq] = qbr(2, iq) / qbr(0, iq);
            vl[iq] = qbl(3, iq) / qbl(0, iq); vr[iq] = qbr(3, iq) / qbr(0, iq);
            quu[iq] = 0.5 * (ul[iq] * qbl(2, iq) + ur[iq] * qbr(2, iq)) + one_plus_eta_edge[iq] * Q_uu_dp_edge(iq, f);
            quv[iq] = 0.5 * (vl[iq] * qbl(2, iq) + vr[iq] * qbr(2, iq)) + one_plus_eta_edge[iq] * Q_uv_dp_edge(iq, f);
            qvu[iq] = 0.5 * (ul[iq] * qbl(3, iq) + ur[iq] * qbr(3, iq)) + one_plus_eta_edge[iq] * Q_uv_dp_edge(iq, f);
            qvv[iq] = 0.5 * (vl[iq] * qbl(3, iq) + vr[iq] * qbr(3, iq)) + one_plus_eta_edge[iq] * Q_vv_dp_edge(iq, f);
            H_face_temp[iq] = (one_plus_eta_edge[iq] * one_plus_eta_edge[iq]) * H_bcl_edge(iq, f);

            btp_mass_flux_face_ave(0, iq, f) += flux_edge_x[iq];
            btp_mass_flux_face_ave(1, iq, f) += flux_edge_y[iq];
            H_face_ave(iq, f) += H_face_temp[iq];
            Qu_face_ave(0, iq, f) += quu[iq]; Qu_face_ave(1, iq, f) += quv[iq];
            Qv_face_ave(0, iq, f) += qvu[iq]; Qv_face_ave(1, iq, f) += qvv[iq];
            double ol = 1.0 + (qbl(1, iq) / pbl[iq]), orr = 1.0 + (qbr(1, iq) / pbr[iq]);
            ope_face_ave(0, iq, f) += ol; ope_face_ave(1, iq, f) += orr;
            ope2_face_ave(0, iq, f) += ol * ol; ope2_face_ave(1, iq, f) += orr * orr;
            one_plus_eta_edge_2_ave(iq, f) += one_plus_eta_edge[iq] * one_plus_eta_edge[iq];
            uvb_face_ave(0, 0, iq, f) += ul[iq]; uvb_face_ave(0, 1, iq, f) += ur[iq];
            uvb_face_ave(1, 0, iq, f) += vl[iq]; uvb_face_ave(1, 1, iq, f) += vr[iq];
        }
        for (int iq = 0; iq < nq; ++iq) {
            double nxl = normal_vector_q(0, iq, f), nyl = normal_vector_q(1, iq, f);
            double H_kx = nxl * H_face_temp[iq], H_ky = nyl * H_face_temp[iq];
            double lamb = coeff_mass_pbpert_LR(iq, f);
            double dispu = 0.5 * lamb * (qbr(2, iq) - qbl(2, iq));
            double dispv = 0.5 * lamb * (qbr(3, iq) - qbl(3, iq));
            double flux_x = nxl * quu[iq] + nyl * quv[iq] - dispu;
            double flux_y = nxl * qvu[iq] + nyl * qvv[iq] - dispv;
            double flux = nxl * flux_edge_x[iq] + nyl * flux_edge_y[iq];
            fl3(0, iq, f) = flux; fl3(1, iq, f) = H_kx + flux_x; fl3(2, iq, f) = H_ky + flux_y;
        }
    }
    // (B) the reference's scatter `rhs(:,il) -= wq*hi*(...)`, `rhs(:,ir) += ...` (mod_rhs_btp.F90:332-366), per element
#pragma omp for schedule(static)
    for (int e = 0; e < nelem; ++e)
        for (int t = ef_ptr[e]; t < ef_ptr[e + 1]; ++t) {
            const int f = ef_face[t], side = ef_side[t];
            const std::vector<int>& fn = side ? fnodeR : fnodeL;
            for (int iq = 0; iq < nq; ++iq) {
                double wq = jac_faceq(iq, f);
                double flux = fl3(0, iq, f), fx = fl3(1, iq, f), fy = fl3(2, iq, f);
                for (int n = 0; n < ngl; ++n) {
                    double hi = psiq(n, iq);
                    int I = fn[(size_t)f * ngl + n];
                    if (side == 0) {
                        rhs(0, I) = rhs(0, I) - wq * hi * flux;
                        rhs(1, I) = rhs(1, I) - wq * hi * fx;
                        rhs(2, I) = rhs(2, I) - wq * hi * fy;
                    } else {
                        rhs(0, I) = rhs(0, I) + wq * hi * flux;
                        rhs(1, I) = rhs(1, I) + wq * hi * fx;
                        rhs(2, I) = rhs(2, I) + wq * hi * fy;
                    }
                }
            }
        }
#pragma omp for schedule(static)
    for (int I = 0; I < npoin; ++I) {
        rhs(0, I) = massinv(I) * rhs(0, I);
        rhs(1, I) = massinv(I) * rhs(1, I);
        rhs(2, I) = massinv(I) * rhs(2, I);
    }
    }
}

// mod_barotropic_terms.F90:411-443.  uv points at a (2,npoin) slice with leading dimension ld
void Oracle::compute_gradient_uv(Arr& grad_uv, const double* uv, int ld) {
    grad_uv.zero();
#pragma omp parallel for schedule(static)
    for (int Iq = 0; Iq < npoin; ++Iq)
        for (int ip = 0; ip < npts; ++ip) {
            int I = index_df[(size_t)Iq * npts + ip];
            double dhdx = dpsidx_df(ip, Iq), dhdy = dpsidy_df(ip, Iq);
            grad_uv(0, Iq) = grad_uv(0, Iq) + dhdx * uv[(size_t)ld * I + 0];
            grad_uv(1, Iq) = grad_uv(1, Iq) + dhdy * uv[(size_t)ld * I + 0];
            grad_uv(2, Iq) = grad_uv(2, Iq) + dhdx * uv[(size_t)ld * I + 1];
            grad_uv(3, Iq) = grad_uv(3, Iq) + dhdy * uv[(size_t)ld * I + 1];
        }
}

// mod_laplacian_quad.F90:32-121 (+ btp_compute_laplacian :357-390, create_rhs_laplacian_flux :427-519)
void Oracle::btp_create_laplacian(Arr& rhs_lap, const Arr& qb) {
    Arr Uk; Uk.alloc(2, npoin);
    Arr graduv; graduv.alloc(4, npoin);
    Arr graduv_face; graduv_face.alloc(4, 2, ngl, nface);
    if (ef_ptr.empty()) build_elem_faces();
#pragma omp parallel for schedule(static)
    for (int I = 0; I < npoin; ++I) { Uk(0, I) = qb(2, I) / qb(0, I); Uk(1, I) = qb(3, I) / qb(0, I); }
    compute_gradient_uv(graduv, Uk.data(), 2);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < graduvb_ave.size(); ++i) graduvb_ave.v[i] += graduv.v[i];
#pragma omp parallel for schedule(static)
    for (int f = 0; f < nface; ++f) {
        int ier = face[8 * f + 7];
        for (int n = 0; n < ngl; ++n) {
            int Iq = fnodeL[(size_t)f * ngl + n];
            for (int v = 0; v < 4; ++v) graduv_face(v, 0, n, f) = graduv(v, Iq);
            if (ier > 0) {
                int Ir = fnodeR[(size_t)f * ngl + n];
                for (int v = 0; v < 4; ++v) graduv_face(v, 1, n, f) = graduv(v, Ir);
            } else {
                for (int v = 0; v < 4; ++v) graduv_face(v, 1, n, f) = graduv_face(v, 0, n, f);
                if (ier == -4) {
                    double nx = normal_vector(0, n, f), ny = normal_vector(1, n, f);
                    double un = graduv_face(0, 0, n, f) * nx + graduv_face(1, 0, n, f) * ny;
                    graduv_face(0, 1, n, f) = graduv_face(0, 0, n, f) - 2.0 * un * nx;
                    graduv_face(1, 1, n, f) = graduv_face(1, 0, n, f) - 2.0 * un * ny;
                    un = graduv_face(2, 0, n, f) * nx + graduv_face(3, 0, n, f) * ny;
                    graduv_face(2, 1, n, f) = graduv_face(2, 0, n, f) - 2.0 * un * nx;
                    graduv_face(3, 1, n, f) = graduv_face(3, 0, n, f) - 2.0 * un * ny;
                }
            }
        }
    }
    // btp_compute_laplacian
    rhs_lap.zero();
#pragma omp parallel for schedule(static)
    for (int e = 0; e < nelem; ++e)
        for (int Iq = e * npts; Iq < (e + 1) * npts; ++Iq) {
            double wq = wjac_df(Iq);
            double qq[4];
            for (int v = 0; v < 4; ++v) qq[v] = pbprime_visc(Iq) * graduv(v, Iq) + btp_dpp_graduv(v, Iq);
            for (int ip = 0; ip < npts; ++ip) {
                int I = index_df[(size_t)Iq * npts + ip];
                rhs_lap(0, I) = rhs_lap(0, I) - wq * (dpsidx_df(ip, Iq) * qq[0] + dpsidy_df(ip, Iq) * qq[1]);
                rhs_lap(1, I) = rhs_lap(1, I) - wq * (dpsidx_df(ip, Iq) * qq[2] + dpsidy_df(ip, Iq) * qq[3]);
            }
        }
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < graduvb_face_ave.size(); ++i) graduvb_face_ave.v[i] += graduv_face.v[i];
    // create_rhs_laplacian_flux (as-written flux form, hazard 2): faces in parallel, then the reference's scatter gathered
    // per element in the reference's face order
    const double beta = 0.5, alpha = 1.0 - beta;
    Arr lfl; lfl.alloc(2, ngl, nface);
#pragma omp parallel for schedule(static)
    for (int f = 0; f < nface; ++f)
        for (int iq = 0; iq < ngl; ++iq) {
            double fv[4][2];
            for (int v = 0; v < 4; ++v) {
                fv[v][0] = btp_graduv_dpp_face(4, 0, iq, f) * graduv_face(v, 0, iq, f) + btp_graduv_dpp_face(v, 0, iq, f);
                fv[v][1] = btp_graduv_dpp_face(4, 1, iq, f) * graduv_face(v, 1, iq, f) + btp_graduv_dpp_face(v, 1, iq, f);
            }
            double nx = normal_vector(0, iq, f), ny = normal_vector(1, iq, f);
            double qu_mean0 = alpha * fv[0][0] + beta * fv[0][1], qu_mean1 = alpha * fv[1][0] + beta * fv[1][1];
            double qv_mean0 = alpha * fv[2][0] + beta * fv[2][1], qv_mean1 = alpha * fv[3][0] + beta * fv[3][1];
            lfl(0, iq, f) = (qu_mean0 - fv[0][0] * nx) + (qu_mean1 - fv[1][0] * ny);
            lfl(1, iq, f) = (qv_mean0 - fv[2][0] * nx) + (qv_mean1 - fv[3][0] * ny);
        }
#pragma omp parallel for schedule(static)
    for (int e = 0; e < nelem; ++e)
        for (int t = ef_ptr[e]; t < ef_ptr[e + 1]; ++t) {
            const int f = ef_face[t], side = ef_side[t];
            const std::vector<int>& fn = side ? fnodeR : fnodeL;
            for (int iq = 0; iq < ngl; ++iq) {
                double wq = jac_face(iq, f), flux_qu = lfl(0, iq, f), flux_qv = lfl(1, iq, f);
                for (int i = 0; i < ngl; ++i) {
                    double hi = psi(i, iq);
                    int ip = fn[(size_t)f * ngl + i];
                    if (side == 0) {
                        rhs_lap(0, ip) = rhs_lap(0, ip) + wq * hi * flux_qu;
                        rhs_lap(1, ip) = rhs_lap(1, ip) + wq * hi * flux_qv;
                    } else {
                        rhs_lap(0, ip) = rhs_lap(0, ip) - wq * hi * flux_qu;
                        rhs_lap(1, ip) = rhs_lap(1, ip) - wq * hi * flux_qv;
                    }
                }
            }
        }
#pragma omp parallel for schedule(static)
    for (int I = 0; I < npoin; ++I) {
        rhs_lap(0, I) = cfg.visc_mlswe * massinv(I) * rhs_lap(0, I);
        rhs_lap(1, I) = cfg.visc_mlswe * massinv(I) * rhs_lap(1, I);
    }
}

// mod_rhs_btp.F90:28-59 (single rank: the pre/post communicators are no-ops)
void Oracle::create_rhs_btp(Arr& rhs, const Arr& qb, const Arr& qprime) {
    Arr qb_df_face; qb_df_face.alloc(4, 2, ngl, nface);
    Arr rhs_visc_btp; rhs_visc_btp.alloc(2, npoin);
    btp_extract_df(qb_df_face, qb);
    create_rhs_btp_volume_qdf(rhs, qb, qprime);
    creat_btp_fluxes_qdf(rhs, qb_df_face);
    if (cfg.method_visc == 1) btp_create_laplacian_v2(rhs_visc_btp, qprime, qb);
    else btp_create_laplacian(rhs_visc_btp, qb);
#pragma omp parallel for schedule(static)
    for (int I = 0; I < npoin; ++I) {
        rhs(1, I) = rhs(1, I) + rhs_visc_btp(0, I);
        rhs(2, I) = rhs(2, I) + rhs_visc_btp(1, I);
    }
}

// mod_barotropic_terms.F90:165-217.  qb is the full (4,npoin) array; slice 3:4 is modified.
void Oracle::btp_mom_boundary_df(Arr& qb) {
    for (int f = 0; f < nface; ++f) {
        int er = face[8 * f + 7];
        if (er == -4) {
            for (int n = 0; n < ngl; ++n) {
                int I = fnodeL[(size_t)f * ngl + n];
                double nx = normal_vector(0, n, f), ny = normal_vector(1, n, f);
                double unl = qb(2, I) * nx + qb(3, I) * ny;
                qb(2, I) = qb(2, I) - unl * nx;
                qb(3, I) = qb(3, I) - unl * ny;
            }
        } else if (er == -2) {
            for (int n = 0; n < ngl; ++n) {
                int I = fnodeL[(size_t)f * ngl + n];
                qb(2, I) = 0.0; qb(3, I) = 0.0;
            }
        }
    }
}

// threaded array assignment (the reference's whole-array `qb1 = qb`)
static void pcopy(Arr& dst, const Arr& src) {
    if (dst.size() != src.size()) { dst = src; return; }
    const size_t n = src.size();
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) dst.v[i] = src.v[i];
}

// mod_rk_mlswe.F90:19-151
void Oracle::ti_barotropic_ssprk_mlswe(Arr& qb, const Arr& qprime) {
    auto t0 = std::chrono::steady_clock::now();
    one_plus_eta_edge_2_ave.zero(); uvb_ave.zero(); uvb_ave_df.zero(); ope_ave.zero(); btp_mass_flux_ave.zero();
    H_ave.zero(); Qu_ave.zero(); Qv_ave.zero(); Quv_ave.zero();
    ope2_ave_df.zero(); uvb_face_ave.zero(); ope_face_ave.zero(); ope2_face_ave.zero(); btp_mass_flux_face_ave.zero();
    H_face_ave.zero(); Qu_face_ave.zero(); Qv_face_ave.zero(); Quv_face_ave.zero();
    tau_wind_ave.zero(); tau_bot_ave.zero(); ope2_ave.zero();
    graduvb_face_ave.zero(); graduvb_ave.zero();

    Arr rhs; rhs.alloc(3, npoin);
    Arr qb0, qb1, qb2; qb2.alloc(4, npoin); qb0.alloc(4, npoin); qb1.alloc(4, npoin);
    for (int mstep = 1; mstep <= N_btp; ++mstep) {
        pcopy(qb0, qb); pcopy(qb1, qb);
        for (int ik = 1; ik <= kstages; ++ik) {
            double dtt = dt_btp * ssprk_beta(ik - 1);
#pragma omp parallel for schedule(static)
            for (int I = 0; I < npoin; ++I) {
                double t = 1.0 + qb1(1, I) * one_over_pbprime_df(I);
                ope2_ave_df(I) = ope2_ave_df(I) + t * t;
                uvb_ave_df(0, I) = uvb_ave_df(0, I) + qb1(2, I) / qb1(0, I);
                uvb_ave_df(1, I) = uvb_ave_df(1, I) + qb1(3, I) / qb1(0, I);
            }
            create_rhs_btp(rhs, qb1, qprime);
            ++stage_count;
            const double a1 = ssprk_a(ik - 1, 0), a2 = ssprk_a(ik - 1, 1), a3 = ssprk_a(ik - 1, 2);
#pragma omp parallel for schedule(static)
            for (int I = 0; I < npoin; ++I) {
                qb(1, I) = a1 * qb0(1, I) + a2 * qb1(1, I) + a3 * qb2(1, I) + dtt * rhs(0, I);
                qb(2, I) = a1 * qb0(2, I) + a2 * qb1(2, I) + a3 * qb2(2, I) + dtt * rhs(1, I);
                qb(3, I) = a1 * qb0(3, I) + a2 * qb1(3, I) + a3 * qb2(3, I) + dtt * rhs(2, I);
                qb(0, I) = qb(1, I) + pbprime_df(I);
            }
            btp_mom_boundary_df(qb);
            pcopy(qb1, qb);
            if (kstages == 5 && ik == 2) pcopy(qb2, qb);
        }
        for (size_t i = 0; i < tau_wind_ave.size(); ++i) tau_wind_ave.v[i] += tau_wind.v[i];
    }
    double N_inv = 1.0 / (double)(kstages * N_btp);
    auto scale = [&](Arr& a, double s) { for (auto& x : a.v) x = s * x; };
    scale(uvb_ave_df, N_inv); scale(graduvb_face_ave, N_inv); scale(graduvb_ave, N_inv);
    for (auto& x : tau_wind_ave.v) x = x / (double)N_btp;
    scale(ope2_ave_df, N_inv); scale(ope2_ave, N_inv);
    scale(ope_ave, N_inv); scale(H_ave, N_inv); scale(Qu_ave, N_inv); scale(Qv_ave, N_inv); scale(Quv_ave, N_inv);
    scale(btp_mass_flux_ave, N_inv); scale(tau_bot_ave, N_inv);
    scale(ope_face_ave, N_inv); scale(ope2_face_ave, N_inv); scale(H_face_ave, N_inv); scale(Qu_face_ave, N_inv);
    scale(Qv_face_ave, N_inv); scale(btp_mass_flux_face_ave, N_inv); scale(one_plus_eta_edge_2_ave, N_inv);
    scale(uvb_ave, N_inv); scale(uvb_face_ave, N_inv);
    btp_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// mod_barotropic_terms.F90:219-409
void Oracle::btp_bcl_coeffs_qdf(const Arr& qprime_df_face, const Arr& qprime) {
    Q_uu_dp.zero(); Q_uv_dp.zero(); Q_vv_dp.zero(); H_bcl.zero();
    Q_uu_dp_edge.zero(); Q_uv_dp_edge.zero(); Q_vv_dp_edge.zero(); H_bcl_edge.zero();
    btp_dpp_graduv.zero(); pbprime_visc.zero();
#pragma omp parallel for schedule(static)
    for (int Iq = 0; Iq < npoin_q; ++Iq) {
        std::vector<double> pprime(nl + 1);
        pprime[0] = 0.0;
        for (int k = 0; k < nl; ++k) {
            double qq[3] = {0, 0, 0};
            for (int ip = 0; ip < npts; ++ip) {
                int I = indexq[(size_t)Iq * npts + ip];
                double hi = psih(ip, Iq);
                for (int v = 0; v < 3; ++v) qq[v] = qq[v] + hi * qprime(v, I, k);
            }
            Q_uu_dp(Iq) = Q_uu_dp(Iq) + qq[1] * (qq[1] * qq[0]);
            Q_uv_dp(Iq) = Q_uv_dp(Iq) + qq[2] * (qq[1] * qq[0]);
            Q_vv_dp(Iq) = Q_vv_dp(Iq) + qq[2] * (qq[2] * qq[0]);
            pprime[k + 1] = pprime[k] + qq[0];
            H_bcl(Iq) = H_bcl(Iq) + 0.5 * alpha_mlswe(k) * (pprime[k + 1] * pprime[k + 1] - pprime[k] * pprime[k]);
        }
    }
    Arr graduv; graduv.alloc(4, npoin);
    for (int k = 0; k < nl; ++k) {
        compute_gradient_uv(graduv, &qprime.v[(size_t)3 * npoin * k + 1], 3);
        for (int I = 0; I < npoin; ++I) {
            for (int v = 0; v < 4; ++v) {
                dpp_graduv(v, I, k) = dpprime_visc(I, k) * graduv(v, I);
                btp_dpp_graduv(v, I) = btp_dpp_graduv(v, I) + dpp_graduv(v, I, k);
            }
            pbprime_visc(I) = pbprime_visc(I) + dpprime_visc(I, k);
        }
    }
#pragma omp parallel for schedule(static)
    for (int f = 0; f < nface; ++f) {
        int ier = face[8 * f + 7];
        std::vector<double> pprime_l(nl + 1), pprime_r(nl + 1);
        for (int iq = 0; iq < nq; ++iq) {
            pprime_l[0] = 0.0; pprime_r[0] = 0.0;
            for (int k = 0; k < nl; ++k) {
                double ql[3] = {0, 0, 0}, qr[3] = {0, 0, 0};
                for (int n = 0; n < ngl; ++n) {
                    double hi = psiq(n, iq);
                    for (int v = 0; v < 3; ++v) {
                        ql[v] = ql[v] + hi * qprime_df_face(v, 0, n, f, k);
                        qr[v] = qr[v] + hi * qprime_df_face(v, 1, n, f, k);
                    }
                }
                Q_uu_dp_edge(iq, f) = Q_uu_dp_edge(iq, f) + 0.5 * ((ql[1] * ql[1] * ql[0]) + (qr[1] * qr[1] * qr[0]));
                Q_uv_dp_edge(iq, f) = Q_uv_dp_edge(iq, f) + 0.5 * ((ql[2] * ql[1] * ql[0]) + (qr[2] * qr[1] * qr[0]));
                Q_vv_dp_edge(iq, f) = Q_vv_dp_edge(iq, f) + 0.5 * ((ql[2] * ql[2] * ql[0]) + (qr[2] * qr[2] * qr[0]));
                pprime_l[k + 1] = pprime_l[k] + ql[0];
                double left_dp = 0.5 * alpha_mlswe(k) * (pprime_l[k + 1] * pprime_l[k + 1] - pprime_l[k] * pprime_l[k]);
                pprime_r[k + 1] = pprime_r[k] + qr[0];
                double right_dp = 0.5 * alpha_mlswe(k) * (pprime_r[k + 1] * pprime_r[k + 1] - pprime_r[k] * pprime_r[k]);
                H_bcl_edge(iq, f) = H_bcl_edge(iq, f) + 0.5 * (left_dp + right_dp);
            }
        }
        for (int n = 0; n < ngl; ++n) {
            int Iq = fnodeL[(size_t)f * ngl + n];
            for (int k = 0; k < nl; ++k) {
                for (int v = 0; v < 4; ++v) graduv_dpp_face(v, 0, n, f, k) = dpp_graduv(v, Iq, k);
                graduv_dpp_face(4, 0, n, f, k) = dpprime_visc(Iq, k);
            }
            if (ier > 0) {
                int Ir = fnodeR[(size_t)f * ngl + n];
                for (int k = 0; k < nl; ++k) {
                    for (int v = 0; v < 4; ++v) graduv_dpp_face(v, 1, n, f, k) = dpp_graduv(v, Ir, k);
                    graduv_dpp_face(4, 1, n, f, k) = dpprime_visc(Ir, k);
                }
            } else {
                for (int k = 0; k < nl; ++k)
                    for (int v = 0; v < 5; ++v) graduv_dpp_face(v, 1, n, f, k) = graduv_dpp_face(v, 0, n, f, k);
                if (ier == -4) {
                    double nx = normal_vector(0, n, f), ny = normal_vector(1, n, f);
                    for (int k = 0; k < nl; ++k) {  // Iq is still the left node here (hazard 6)
                        double un = dpp_graduv(0, Iq, k) * nx + dpp_graduv(1, Iq, k) * ny;
                        graduv_dpp_face(0, 1, n, f, k) = dpp_graduv(0, Iq, k) - 2.0 * un * nx;
                        graduv_dpp_face(1, 1, n, f, k) = dpp_graduv(1, Iq, k) - 2.0 * un * ny;
                        un = dpp_graduv(2, Iq, k) * nx + dpp_graduv(3, Iq, k) * ny;
                        graduv_dpp_face(2, 1, n, f, k) = dpp_graduv(2, Iq, k) - 2.0 * un * nx;
                        graduv_dpp_face(3, 1, n, f, k) = dpp_graduv(3, Iq, k) - 2.0 * un * ny;
                    }
                }
            }
        }
    }
    btp_graduv_dpp_face.zero();
    for (int f = 0; f < nface; ++f)
        for (int n = 0; n < ngl; ++n)
            for (int k = 0; k < nl; ++k)
                for (int v = 0; v < 5; ++v) {
                    btp_graduv_dpp_face(v, 0, n, f) = btp_graduv_dpp_face(v, 0, n, f) + graduv_dpp_face(v, 0, n, f, k);
                    btp_graduv_dpp_face(v, 1, n, f) = btp_graduv_dpp_face(v, 1, n, f) + graduv_dpp_face(v, 1, n, f, k);
                }
}

}  // namespace orc
