// ============================================================================
// oracle_bcl.cpp -- TEST INFRASTRUCTURE ONLY (CPU oracle, baroclinic half).
// As-written restatement of
//   ti_rk_bcl                      src/ti_rk_bcl.F90:9-87
//   thickness/momentum/momentum_mass/rhs_momentum/apply_consistency
//                                  src/mod_splitting.F90:25-366
//   mod_layer_terms                src/mod_layer_terms.F90:57-137,198-320,354-465,529-584
//   mod_create_rhs_mlswe           src/mod_create_rhs_mlswe.F90:28-101,281-1115
//   bcl_create_laplacian           src/mod_laplacian_quad.F90:227-248,392-425,521-611
// Single rank: bcl_create_communicator calls are no-ops.
// ============================================================================
#include <algorithm>

#include "hnumo_oracle.hpp"

namespace orc {

// mod_layer_terms.F90:354-415
void Oracle::extract_qprime_df_face(Arr& qf, const Arr& qprime) {
    qf.zero();
    for (int f = 0; f < nface; ++f) {
        int er = face[8 * f + 7];
        for (int n = 0; n < ngl; ++n) {
            int I = fnodeL[(size_t)f * ngl + n];
            for (int k = 0; k < nl; ++k)
                for (int v = 0; v < 3; ++v) qf(v, 0, n, f, k) = qprime(v, I, k);
            if (er > 0) {
                int Ir = fnodeR[(size_t)f * ngl + n];
                for (int k = 0; k < nl; ++k)
                    for (int v = 0; v < 3; ++v) qf(v, 1, n, f, k) = qprime(v, Ir, k);
            } else {
                for (int k = 0; k < nl; ++k)
                    for (int v = 0; v < 3; ++v) qf(v, 1, n, f, k) = qf(v, 0, n, f, k);
                if (er == -4) {
                    double nx = normal_vector(0, n, f), ny = normal_vector(1, n, f);
                    for (int k = 0; k < nl; ++k) {
                        double un = qprime(1, I, k) * nx + qprime(2, I, k) * ny;
                        qf(1, 1, n, f, k) = qprime(1, I, k) - 2.0 * un * nx;
                        qf(2, 1, n, f, k) = qprime(2, I, k) - 2.0 * un * ny;
                    }
                } else if (er == -2) {
                    for (int k = 0; k < nl; ++k) {
                        qf(1, 1, n, f, k) = -qf(1, 0, n, f, k);
                        qf(2, 1, n, f, k) = -qf(2, 0, n, f, k);
                    }
                }
            }
        }
    }
}

// mod_layer_terms.F90:417-465 applied to component 1 of qprime_df_face (mod_splitting.F90:89)
void Oracle::extract_dprime_df_face(Arr& qf, const Arr& qprime) {
    for (int f = 0; f < nface; ++f) {
        int er = face[8 * f + 7];
        for (int n = 0; n < ngl; ++n) {
            int I = fnodeL[(size_t)f * ngl + n];
            for (int k = 0; k < nl; ++k) qf(0, 0, n, f, k) = qprime(0, I, k);
            if (er > 0) {
                int Ir = fnodeR[(size_t)f * ngl + n];
                for (int k = 0; k < nl; ++k) qf(0, 1, n, f, k) = qprime(0, Ir, k);
            } else {
                for (int k = 0; k < nl; ++k) qf(0, 1, n, f, k) = qf(0, 0, n, f, k);
            }
        }
    }
}

// mod_create_rhs_mlswe.F90:53-78 (volume :822-877, flux :922-1034)
void Oracle::layer_mass_rhs(Arr& dp_advec, const Arr& qprime, const Arr& qf) {
    dp_advec.zero();
    sum_layer_mass_flux.zero();
    const int nq2 = nq * nq;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < nelem; ++e)
        for (int Iq = e * nq2; Iq < (e + 1) * nq2; ++Iq) {
            double qb0 = ope_ave(Iq), qb1 = uvb_ave(0, Iq), qb2 = uvb_ave(1, Iq);
            double wq = wjac(Iq);
            for (int k = 0; k < nl; ++k) {
                double qp[3] = {0, 0, 0};
                for (int ip = 0; ip < npts; ++ip) {
                    int I = indexq[(size_t)Iq * npts + ip];
                    double hi = psih(ip, Iq);
                    for (int v = 0; v < 3; ++v) qp[v] = qp[v] + hi * qprime(v, I, k);
                }
                double dp_temp = qp[0] * qb0;
                double udp = (qp[1] + qb1) * dp_temp;
                double vdp = (qp[2] + qb2) * dp_temp;
                sum_layer_mass_flux(0, Iq) = sum_layer_mass_flux(0, Iq) + udp;
                sum_layer_mass_flux(1, Iq) = sum_layer_mass_flux(1, Iq) + vdp;
                for (int ip = 0; ip < npts; ++ip) {
                    int I = indexq[(size_t)Iq * npts + ip];
                    dp_advec(I, k) = dp_advec(I, k) + wq * (dpsidx(ip, Iq) * udp + dpsidy(ip, Iq) * vdp);
                }
            }
        }
    // create_layer_mass_flux
    sum_layer_mass_flux_face.zero();
    std::vector<double> flux_edge_u(nq), flux_edge_v(nq);
    for (int f = 0; f < nface; ++f) {
        int er = face[8 * f + 7];
        for (int k = 0; k < nl; ++k) {
            for (int iq = 0; iq < nq; ++iq) {
                double ql[3] = {0, 0, 0}, qr[3] = {0, 0, 0};
                for (int n = 0; n < ngl; ++n) {
                    double hi = psiq(n, iq);
                    for (int v = 0; v < 3; ++v) {
                        ql[v] = ql[v] + hi * qf(v, 0, n, f, k);
                        qr[v] = qr[v] + hi * qf(v, 1, n, f, k);
                    }
                }
                double nxl = normal_vector_q(0, iq, f), nyl = normal_vector_q(1, iq, f);
                double uu = 0.5 * ((ql[1] + uvb_face_ave(0, 0, iq, f)) + (qr[1] + uvb_face_ave(0, 1, iq, f)));
                double vv = 0.5 * ((ql[2] + uvb_face_ave(1, 0, iq, f)) + (qr[2] + uvb_face_ave(1, 1, iq, f)));
                double dpl = ope_face_ave(0, iq, f) * ql[0];
                double dpr = ope_face_ave(1, iq, f) * qr[0];
                flux_edge_u[iq] = (uu * nxl > 0.0) ? uu * dpl : uu * dpr;
                flux_edge_v[iq] = (vv * nyl > 0.0) ? vv * dpl : vv * dpr;
            }
            for (int iq = 0; iq < nq; ++iq) {
                sum_layer_mass_flux_face(0, iq, f) += flux_edge_u[iq];
                sum_layer_mass_flux_face(1, iq, f) += flux_edge_v[iq];
            }
            for (int iq = 0; iq < nq; ++iq) {
                double wq = jac_faceq(iq, f);
                double nxl = normal_vector_q(0, iq, f), nyl = normal_vector_q(1, iq, f);
                double flux = nxl * flux_edge_u[iq] + nyl * flux_edge_v[iq];
                for (int n = 0; n < ngl; ++n) {
                    double hi = psiq(n, iq);
                    int I = fnodeL[(size_t)f * ngl + n];
                    dp_advec(I, k) = dp_advec(I, k) - wq * hi * flux;
                    if (er > 0) {
                        int Ir = fnodeR[(size_t)f * ngl + n];
                        dp_advec(Ir, k) = dp_advec(Ir, k) + wq * hi * flux;
                    }
                }
            }
        }
    }
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) dp_advec(I, k) = massinv(I) * dp_advec(I, k);
}

// mod_splitting.F90:324-366 with evaluate_consistency_face (mod_layer_terms.F90:57-137) and
// consistency_mass_rhs (mod_create_rhs_mlswe.F90:80-101, 879-920, 1036-1115)
void Oracle::apply_consistency(Arr& q) {
    Arr dpprime_df; dpprime_df.alloc(npoin, nl);
    Arr dp_advec; dp_advec.alloc(npoin, nl);
    Arr deficit; deficit.alloc(2, 2, nq, nface, nl);
    std::vector<double> ope(npoin, 0.0);
    for (int I = 0; I < npoin; ++I) {
        double s = 0.0;
        for (int k = 0; k < nl; ++k) s += q(0, I, k);
        ope[I] = s / pbprime_df(I);
    }
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) dpprime_df(I, k) = q(0, I, k) / ope[I];
    // evaluate_consistency_face
    for (int k = 0; k < nl; ++k)
        for (int f = 0; f < nface; ++f) {
            int er = face[8 * f + 7];
            for (int iq = 0; iq < nq; ++iq) {
                double qprime_l = 0.0, qprime_r = 0.0;
                for (int n = 0; n < ngl; ++n) qprime_l = qprime_l + psiq(n, iq) * dpprime_df(fnodeL[(size_t)f * ngl + n], k);
                if (er > 0) {
                    for (int n = 0; n < ngl; ++n) qprime_r = qprime_r + psiq(n, iq) * dpprime_df(fnodeR[(size_t)f * ngl + n], k);
                } else qprime_r = qprime_l;
                double wl = qprime_l / pbprime_face(0, iq, f), wr = qprime_r / pbprime_face(1, iq, f);
                double d0 = btp_mass_flux_face_ave(0, iq, f) - sum_layer_mass_flux_face(0, iq, f);
                double d1 = btp_mass_flux_face_ave(1, iq, f) - sum_layer_mass_flux_face(1, iq, f);
                deficit(0, 0, iq, f, k) = wl * d0; deficit(1, 0, iq, f, k) = wl * d1;
                deficit(0, 1, iq, f, k) = wr * d0; deficit(1, 1, iq, f, k) = wr * d1;
            }
        }
    // create_consistency_volume_mass
    const int nq2 = nq * nq;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < nelem; ++e)
        for (int k = 0; k < nl; ++k)
            for (int Iq = e * nq2; Iq < (e + 1) * nq2; ++Iq) {
                double dp = 0.0;
                for (int ip = 0; ip < npts; ++ip) dp = dp + psih(ip, Iq) * dpprime_df(indexq[(size_t)Iq * npts + ip], k);
                double weight = dp / pbprime(Iq);
                double udp = weight * (btp_mass_flux_ave(0, Iq) - sum_layer_mass_flux(0, Iq));
                double vdp = weight * (btp_mass_flux_ave(1, Iq) - sum_layer_mass_flux(1, Iq));
                double wq = wjac(Iq);
                for (int ip = 0; ip < npts; ++ip) {
                    int I = indexq[(size_t)Iq * npts + ip];
                    dp_advec(I, k) = dp_advec(I, k) + wq * (dpsidx(ip, Iq) * udp + dpsidy(ip, Iq) * vdp);
                }
            }
    // create_consistency_mass_flux
    std::vector<double> flux_edge_u(nq), flux_edge_v(nq);
    for (int k = 0; k < nl; ++k)
        for (int f = 0; f < nface; ++f) {
            int er = face[8 * f + 7];
            for (int iq = 0; iq < nq; ++iq) {
                double nxl = normal_vector_q(0, iq, f), nyl = normal_vector_q(1, iq, f);
                flux_edge_u[iq] = (deficit(0, 0, iq, f, k) * nxl > 0.0) ? deficit(0, 0, iq, f, k) : deficit(0, 1, iq, f, k);
                flux_edge_v[iq] = (deficit(1, 0, iq, f, k) * nyl > 0.0) ? deficit(1, 0, iq, f, k) : deficit(1, 1, iq, f, k);
            }
            for (int iq = 0; iq < nq; ++iq) {
                double wq = jac_faceq(iq, f);
                double nxl = normal_vector_q(0, iq, f), nyl = normal_vector_q(1, iq, f);
                double flux = nxl * flux_edge_u[iq] + nyl * flux_edge_v[iq];
                for (int n = 0; n < ngl; ++n) {
                    double hi = psiq(n, iq);
                    int I = fnodeL[(size_t)f * ngl + n];
                    dp_advec(I, k) = dp_advec(I, k) - wq * hi * flux;
                    if (er > 0) {
                        int Ir = fnodeR[(size_t)f * ngl + n];
                        dp_advec(Ir, k) = dp_advec(Ir, k) + wq * hi * flux;
                    }
                }
            }
        }
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) q(0, I, k) = q(0, I, k) + dt * massinv(I) * dp_advec(I, k);
}

// mod_laplacian_quad.F90:227-248 (+ :392-425 volume, :521-611 flux)
void Oracle::bcl_create_laplacian(Arr& rhs_lap) {
    rhs_lap.zero();
    Arr rhs_temp; rhs_temp.alloc(2, npoin);
    const double beta = 0.5, alpha = 1.0 - beta;
    for (int k = 0; k < nl; ++k) {
        rhs_temp.zero();
#pragma omp parallel for schedule(static)
        for (int e = 0; e < nelem; ++e)
            for (int Iq = e * npts; Iq < (e + 1) * npts; ++Iq) {
                double wq = wjac_df(Iq);
                double qq[4];
                for (int v = 0; v < 4; ++v) qq[v] = dpprime_visc(Iq, k) * graduvb_ave(v, Iq) + dpp_graduv(v, Iq, k);
                for (int ip = 0; ip < npts; ++ip) {
                    int I = index_df[(size_t)Iq * npts + ip];
                    rhs_temp(0, I) = rhs_temp(0, I) - wq * (dpsidx_df(ip, Iq) * qq[0] + dpsidy_df(ip, Iq) * qq[1]);
                    rhs_temp(1, I) = rhs_temp(1, I) - wq * (dpsidx_df(ip, Iq) * qq[2] + dpsidy_df(ip, Iq) * qq[3]);
                }
            }
        for (int f = 0; f < nface; ++f) {
            int ier = face[8 * f + 7];
            for (int iq = 0; iq < ngl; ++iq) {
                double fv[4][2];
                for (int v = 0; v < 4; ++v) {
                    fv[v][0] = graduv_dpp_face(4, 0, iq, f, k) * graduvb_face_ave(v, 0, iq, f) + graduv_dpp_face(v, 0, iq, f, k);
                    fv[v][1] = graduv_dpp_face(4, 1, iq, f, k) * graduvb_face_ave(v, 1, iq, f) + graduv_dpp_face(v, 1, iq, f, k);
                }
                double nx = normal_vector(0, iq, f), ny = normal_vector(1, iq, f);
                double qu_mean0 = alpha * fv[0][0] + beta * fv[0][1], qu_mean1 = alpha * fv[1][0] + beta * fv[1][1];
                double qv_mean0 = alpha * fv[2][0] + beta * fv[2][1], qv_mean1 = alpha * fv[3][0] + beta * fv[3][1];
                double wq = jac_face(iq, f);
                double flux_qu = (qu_mean0 - fv[0][0] * nx) + (qu_mean1 - fv[1][0] * ny);
                double flux_qv = (qv_mean0 - fv[2][0] * nx) + (qv_mean1 - fv[3][0] * ny);
                for (int i = 0; i < ngl; ++i) {
                    double hi = psi(i, iq);
                    int ip = fnodeL[(size_t)f * ngl + i];
                    rhs_temp(0, ip) = rhs_temp(0, ip) + wq * hi * flux_qu;
                    rhs_temp(1, ip) = rhs_temp(1, ip) + wq * hi * flux_qv;
                    if (ier > 0) {
                        int ipr = fnodeR[(size_t)f * ngl + i];
                        rhs_temp(0, ipr) = rhs_temp(0, ipr) - wq * hi * flux_qu;
                        rhs_temp(1, ipr) = rhs_temp(1, ipr) - wq * hi * flux_qv;
                    }
                }
            }
        }
        for (int I = 0; I < npoin; ++I) {
            rhs_lap(0, I, k) = cfg.visc_mlswe * massinv(I) * rhs_temp(0, I);
            rhs_lap(1, I, k) = cfg.visc_mlswe * massinv(I) * rhs_temp(1, I);
        }
    }
}

// mod_create_rhs_mlswe.F90:281-456 (create_rhs_dynamics_volume_layers)
void Oracle::layer_momentum_volume(Arr& rhs_mom, const Arr& qprime, const Arr& q) {
    const double eps1 = 1.0e-20;
    rhs_mom.zero();
    const double Pstress = (gravity / alpha_mlswe(0)) * 50.0;
    const double Pbstress = (gravity / alpha_mlswe(nl - 1)) * 10.0;
    Arr z_elv; z_elv.alloc(npoin, nl + 1);
    for (int I = 0; I < npoin; ++I) z_elv(I, nl) = zbot_df(I);
    for (int k = nl - 1; k >= 0; --k)
        for (int I = 0; I < npoin; ++I)
            z_elv(I, k) = z_elv(I, k + 1) + (alpha_mlswe(k) / gravity) * (std::sqrt(ope2_ave_df(I)) * qprime(0, I, k));
    const int nq2 = nq * nq;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < nelem; ++e) {
        std::vector<double> p_tmp(nl + 1), temp_uu(nl), temp_vv(nl), H_tmp(nl), u_udp(nl), v_vdp(nl), u_vdp1(nl), u_vdp2(nl);
        std::vector<double> gradz1(nl + 1), gradz2(nl + 1), pprime_temp(nl + 1), dpq(nl);
        double qp[3] = {0, 0, 0};
        for (int Iq = e * nq2; Iq < (e + 1) * nq2; ++Iq) {
            p_tmp[0] = 0.0;
            std::fill(temp_uu.begin(), temp_uu.end(), 0.0); std::fill(temp_vv.begin(), temp_vv.end(), 0.0);
            for (int k = 0; k < nl; ++k) {
                qp[0] = qp[1] = qp[2] = 0.0;
                for (int ip = 0; ip < npts; ++ip) {
                    int I = indexq[(size_t)Iq * npts + ip];
                    double hi = psih(ip, Iq);
                    for (int v = 0; v < 3; ++v) qp[v] = qp[v] + hi * qprime(v, I, k);
                    temp_uu[k] = temp_uu[k] + hi * q(1, I, k);
                    temp_vv[k] = temp_vv[k] + hi * q(2, I, k);
                }
                dpq[k] = qp[0];
                double qb0 = ope_ave(Iq), qb1 = uvb_ave(0, Iq), qb2 = uvb_ave(1, Iq);
                p_tmp[k + 1] = p_tmp[k] + std::sqrt(ope2_ave(Iq)) * qp[0];
                H_tmp[k] = 0.5 * alpha_mlswe(k) * (p_tmp[k + 1] * p_tmp[k + 1] - p_tmp[k] * p_tmp[k]);
                double dp = qp[0] * qb0, u = qp[1] + qb1, v = qp[2] + qb2;
                u_udp[k] = dp * u * u;
                v_vdp[k] = dp * v * v;
                u_vdp1[k] = u * v * dp;
                u_vdp2[k] = v * u * dp;
                temp_uu[k] = std::fabs(temp_uu[k]) + eps1;
                temp_vv[k] = std::fabs(temp_vv[k]) + eps1;
            }
            std::fill(gradz1.begin(), gradz1.end(), 0.0); std::fill(gradz2.begin(), gradz2.end(), 0.0);
            for (int ip = 0; ip < npts; ++ip) {
                int I = indexq[(size_t)Iq * npts + ip];
                for (int k = 0; k <= nl; ++k) {
                    gradz1[k] = gradz1[k] + dpsidx(ip, Iq) * z_elv(I, k);
                    gradz2[k] = gradz2[k] + dpsidy(ip, Iq) * z_elv(I, k);
                }
            }
            double s_uu = 0, s_uv = 0, s_vv = 0, s_tu = 0, s_tv = 0, s_H = 0;
            for (int k = 0; k < nl; ++k) { s_uu += u_udp[k]; s_uv += u_vdp1[k]; s_vv += v_vdp[k]; s_tu += temp_uu[k]; s_tv += temp_vv[k]; s_H += H_tmp[k]; }
            double uu_dp_deficitq = Qu_ave(Iq) - s_uu;
            double uv_dp_deficitq = Quv_ave(Iq) - s_uv;
            double vv_dp_deficitq = Qv_ave(Iq) - s_vv;
            double one_over_sumuq = 1.0 / s_tu, one_over_sumvq = 1.0 / s_tv;
            double wq = wjac(Iq);
            std::fill(pprime_temp.begin(), pprime_temp.end(), 0.0);
            for (int k = 0; k < nl; ++k) {
                // hazard 1 (mod_create_rhs_mlswe.F90:382): qp(k) indexes the (dp',u',v') 3-vector of the LAST layer by
                // layer number.  Defined for nl<=3; for nl>3 the reference reads out of bounds, there we use the
                // evident intent dp'_k ("reference-undefined, intent semantics", SURVEY 8 hazard 1).
                double inc = (nl <= 3) ? qp[k] : dpq[k];
                pprime_temp[k + 1] = pprime_temp[k] + inc;
                double weightq = temp_uu[k] * one_over_sumuq;
                u_udp[k] = u_udp[k] + weightq * uu_dp_deficitq;
                u_vdp1[k] = u_vdp1[k] + weightq * uv_dp_deficitq;
                weightq = temp_vv[k] * one_over_sumvq;
                u_vdp2[k] = u_vdp2[k] + weightq * uv_dp_deficitq;
                v_vdp[k] = v_vdp[k] + weightq * vv_dp_deficitq;
                double Hq = H_tmp[k];
                double weight = 1.0;
                double acceleration = s_H;
                if (acceleration > 0.0) weight = H_ave(Iq) / acceleration;
                Hq = Hq * weight;
                double var_uu = u_udp[k], var_uv = u_vdp1[k], var_vu = u_vdp2[k], var_vv = v_vdp[k];
                double temp1 = (std::min(pprime_temp[k + 1], Pstress) - std::min(pprime_temp[k], Pstress)) / Pstress;
                double tau_wind_u = temp1 * tau_wind(0, Iq), tau_wind_v = temp1 * tau_wind(1, Iq);
                double tempbot = std::min(Pbstress, pbprime(Iq) - pprime_temp[k + 1]) - std::min(Pbstress, pbprime(Iq) - pprime_temp[k]);
                tempbot = tempbot / Pbstress;
                double source_x = gravity * (tau_wind_u - tempbot * tau_bot_ave(0, Iq) + p_tmp[k] * gradz1[k] - p_tmp[k + 1] * gradz1[k + 1]);
                double source_y = gravity * (tau_wind_v - tempbot * tau_bot_ave(1, Iq) + p_tmp[k] * gradz2[k] - p_tmp[k + 1] * gradz2[k + 1]);
                for (int ip = 0; ip < npts; ++ip) {
                    int I = indexq[(size_t)Iq * npts + ip];
                    double hi = psih(ip, Iq), dhdx = dpsidx(ip, Iq), dhdy = dpsidy(ip, Iq);
                    rhs_mom(0, I, k) = rhs_mom(0, I, k) + wq * (hi * source_x + dhdx * (Hq + var_uu) + var_uv * dhdy);
                    rhs_mom(1, I, k) = rhs_mom(1, I, k) + wq * (hi * source_y + var_vu * dhdx + dhdy * (Hq + var_vv));
                }
            }
        }
    }
}

// mod_create_rhs_mlswe.F90:458-820
void Oracle::apply_layers_fluxes(Arr& rhs_mom, const Arr& qf) {
    const double eps1 = 1.0e-20;
    std::vector<double> alpha_over_g(nl), g_over_alpha(nl);
    for (int k = 0; k < nl; ++k) { alpha_over_g[k] = alpha_mlswe(k) / gravity; g_over_alpha[k] = gravity / alpha_mlswe(k); }
    Arr p_face, z_face; p_face.alloc(2, nl + 1); z_face.alloc(2, nl + 1);
    std::vector<double> p_edge_plus(nl + 1), p_edge_minus(nl + 1), p2l(nl + 1), p2r(nl + 1), z_edge_plus(nl + 1), z_edge_minus(nl + 1);
    Arr ql, qr; ql.alloc(3, nq, nl); qr.alloc(3, nq, nl);
    Arr udpl, udpr, vdpl, vdpr; udpl.alloc(nq, nl); udpr.alloc(nq, nl); vdpl.alloc(nq, nl); vdpr.alloc(nq, nl);
    Arr H_face, udp_flux, vdp_flux; H_face.alloc(2, nq, nl); udp_flux.alloc(2, nq, nl); vdp_flux.alloc(2, nq, nl);
    for (int f = 0; f < nface; ++f) {
        int er = face[8 * f + 7];
        ql.zero(); qr.zero();
        for (int iq = 0; iq < nq; ++iq) {
            double qbl0 = ope_face_ave(0, iq, f), qbl1 = uvb_face_ave(0, 0, iq, f), qbl2 = uvb_face_ave(1, 0, iq, f);
            double qbr0 = ope_face_ave(1, iq, f), qbr1 = uvb_face_ave(0, 1, iq, f), qbr2 = uvb_face_ave(1, 1, iq, f);
            double nxl = normal_vector_q(0, iq, f), nyl = normal_vector_q(1, iq, f);
            for (int k = 0; k < nl; ++k) {
                for (int n = 0; n < ngl; ++n) {
                    double hi = psiq(n, iq);
                    for (int v = 0; v < 3; ++v) {
                        ql(v, iq, k) = ql(v, iq, k) + hi * qf(v, 0, n, f, k);
                        qr(v, iq, k) = qr(v, iq, k) + hi * qf(v, 1, n, f, k);
                    }
                }
                double dpl = qbl0 * ql(0, iq, k), dpr = qbr0 * qr(0, iq, k);
                double ul = ql(1, iq, k) + qbl1, ur = qr(1, iq, k) + qbr1;
                double vl = ql(2, iq, k) + qbl2, vr = qr(2, iq, k) + qbr2;
                double uu = 0.5 * (ul + ur), vv = 0.5 * (vl + vr);
                udpl(iq, k) = ul * dpl; udpr(iq, k) = ur * dpr; vdpl(iq, k) = vl * dpl; vdpr(iq, k) = vr * dpr;
                if (uu * nxl > 0.0) { udp_flux(0, iq, k) = uu * (ul * dpl); vdp_flux(0, iq, k) = uu * (vl * dpl); }
                else { udp_flux(0, iq, k) = uu * (ur * dpr); vdp_flux(0, iq, k) = uu * (vr * dpr); }
                if (vv * nyl > 0.0) { udp_flux(1, iq, k) = vv * (ul * dpl); vdp_flux(1, iq, k) = vv * (vl * dpl); }
                else { udp_flux(1, iq, k) = vv * (ur * dpr); vdp_flux(1, iq, k) = vv * (vr * dpr); }
            }
            double su0 = 0, su1 = 0, sv0 = 0, sv1 = 0;
            for (int k = 0; k < nl; ++k) { su0 += udp_flux(0, iq, k); su1 += udp_flux(1, iq, k); sv0 += vdp_flux(0, iq, k); sv1 += vdp_flux(1, iq, k); }
            double uu_dp_flux_deficit = Qu_face_ave(0, iq, f) - su0;
            double uv_dp_flux_deficit = Qu_face_ave(1, iq, f) - su1;
            double vu_dp_flux_deficit = Qv_face_ave(0, iq, f) - sv0;
            double vv_dp_flux_deficit = Qv_face_ave(1, iq, f) - sv1;
            double sl = 0, sr = 0;
            for (int k = 0; k < nl; ++k) { sl += std::fabs(udpl(iq, k)) + eps1; sr += std::fabs(udpr(iq, k)) + eps1; }
            double one_over_sum_l = 1.0 / sl, one_over_sum_r = 1.0 / sr;
            for (int k = 0; k < nl; ++k) {
                double w = (uu_dp_flux_deficit * nxl > 0.0) ? std::fabs(udpl(iq, k)) * one_over_sum_l : std::fabs(udpr(iq, k)) * one_over_sum_r;
                udp_flux(0, iq, k) = udp_flux(0, iq, k) + w * uu_dp_flux_deficit;
            }
            for (int k = 0; k < nl; ++k) {
                double w = (uv_dp_flux_deficit * nyl > 0.0) ? std::fabs(udpl(iq, k)) * one_over_sum_l : std::fabs(udpr(iq, k)) * one_over_sum_r;
                udp_flux(1, iq, k) = udp_flux(1, iq, k) + w * uv_dp_flux_deficit;
            }
            sl = 0; sr = 0;
            for (int k = 0; k < nl; ++k) { sl += std::fabs(vdpl(iq, k)) + eps1; sr += std::fabs(vdpr(iq, k)) + eps1; }
            one_over_sum_l = 1.0 / sl; one_over_sum_r = 1.0 / sr;
            for (int k = 0; k < nl; ++k) {
                double w = (vu_dp_flux_deficit * nxl > 0.0) ? std::fabs(vdpl(iq, k)) * one_over_sum_l : std::fabs(vdpr(iq, k)) * one_over_sum_r;
                vdp_flux(0, iq, k) = vdp_flux(0, iq, k) + w * vu_dp_flux_deficit;
            }
            for (int k = 0; k < nl; ++k) {
                double w = (vv_dp_flux_deficit * nyl > 0.0) ? std::fabs(vdpl(iq, k)) * one_over_sum_l : std::fabs(vdpr(iq, k)) * one_over_sum_r;
                vdp_flux(1, iq, k) = vdp_flux(1, iq, k) + w * vv_dp_flux_deficit;
            }

            z_face.zero(); p_face.zero();
            std::fill(z_edge_plus.begin(), z_edge_plus.end(), 0.0); std::fill(z_edge_minus.begin(), z_edge_minus.end(), 0.0);
            std::fill(p_edge_plus.begin(), p_edge_plus.end(), 0.0); std::fill(p_edge_minus.begin(), p_edge_minus.end(), 0.0);
            double ope_l = std::sqrt(ope2_face_ave(0, iq, f)), ope_r = std::sqrt(ope2_face_ave(1, iq, f));
            for (int k = 0; k < nl; ++k) {
                p_face(0, k + 1) = p_face(0, k) + ope_l * ql(0, iq, k);
                p_face(1, k + 1) = p_face(1, k) + ope_r * qr(0, iq, k);
            }
            double one_plus_eta_edge = std::sqrt(one_plus_eta_edge_2_ave(iq, f));
            z_face(0, nl) = zbot_face(0, iq, f); z_face(1, nl) = zbot_face(1, iq, f);
            z_edge_plus[nl] = zbot_face(0, iq, f); z_edge_minus[nl] = zbot_face(1, iq, f);
            for (int k = nl - 1; k >= 0; --k) {
                z_face(0, k) = z_face(0, k + 1) + alpha_over_g[k] * (ope_l * ql(0, iq, k));
                z_face(1, k) = z_face(1, k + 1) + alpha_over_g[k] * (ope_r * qr(0, iq, k));
                z_edge_plus[k] = z_edge_plus[k + 1] + alpha_over_g[k] * (one_plus_eta_edge * ql(0, iq, k));
                z_edge_minus[k] = z_edge_minus[k + 1] + alpha_over_g[k] * (one_plus_eta_edge * qr(0, iq, k));
            }
            p_edge_plus[1] = one_plus_eta_edge * ql(0, iq, 0);
            p_edge_minus[1] = one_plus_eta_edge * qr(0, iq, 0);
            for (int k = 1; k < nl; ++k) {
                p_edge_plus[k + 1] = p_edge_plus[k] + one_plus_eta_edge * ql(0, iq, k);
                p_edge_minus[k + 1] = p_edge_minus[k] + one_plus_eta_edge * qr(0, iq, k);
            }
            for (int k = 0; k < nl; ++k) {
                double H_r_plus = 0.5 * alpha_mlswe(k) * (p_edge_plus[k + 1] * p_edge_plus[k + 1] - p_edge_plus[k] * p_edge_plus[k]);
                double H_r_minus = 0.0;
                for (int kt = 0; kt < nl; ++kt) {
                    double z_top = std::min(z_edge_minus[kt], z_edge_plus[k]);
                    double z_bot = std::max(z_edge_minus[kt + 1], z_edge_plus[k + 1]);
                    double dz = z_top - z_bot;
                    if (dz > 0.0) {
                        double p_bot = p_edge_minus[kt + 1] - g_over_alpha[kt] * (z_bot - z_edge_minus[kt + 1]);
                        double p_top = p_edge_minus[kt + 1] - g_over_alpha[kt] * (z_top - z_edge_minus[kt + 1]);
                        H_r_minus = H_r_minus + 0.5 * alpha_mlswe(kt) * (p_bot * p_bot - p_top * p_top);
                    }
                }
                H_face(0, iq, k) = 0.5 * (H_r_plus + H_r_minus);
                H_r_minus = 0.5 * alpha_mlswe(k) * (p_edge_minus[k + 1] * p_edge_minus[k + 1] - p_edge_minus[k] * p_edge_minus[k]);
                H_r_plus = 0.0;
                for (int kt = 0; kt < nl; ++kt) {
                    double z_top = std::min(z_edge_plus[kt], z_edge_minus[k]);
                    double z_bot = std::max(z_edge_plus[kt + 1], z_edge_minus[k + 1]);
                    double dz = z_top - z_bot;
                    if (dz > 0.0) {
                        double p_bot = p_edge_plus[kt + 1] - g_over_alpha[kt] * (z_bot - z_edge_plus[kt + 1]);
                        double p_top = p_edge_plus[kt + 1] - g_over_alpha[kt] * (z_top - z_edge_plus[kt + 1]);
                        H_r_plus = H_r_plus + 0.5 * alpha_mlswe(kt) * (p_bot * p_bot - p_top * p_top);
                    }
                }
                H_face(1, iq, k) = 0.5 * (H_r_plus + H_r_minus);
            }
            if (er == -4) {
                std::fill(p2l.begin(), p2l.end(), 0.0); std::fill(p2r.begin(), p2r.end(), 0.0);
                for (int k = 0; k < nl; ++k) {
                    p2l[k + 1] = p_face(0, k + 1);
                    H_face(0, iq, k) = 0.5 * alpha_mlswe(k) * (p2l[k + 1] * p2l[k + 1] - p2l[k] * p2l[k]);
                    p2r[k + 1] = p_face(1, k + 1);
                    H_face(1, iq, k) = 0.5 * alpha_mlswe(k) * (p2r[k + 1] * p2r[k + 1] - p2r[k] * p2r[k]);
                }
            }
            if (er != -4) {
                for (int k = 0; k < nl - 1; ++k) {
                    double p_inc1 = g_over_alpha[k] * (z_face(0, k + 1) - z_edge_plus[k + 1]);
                    double H_corr1 = 0.5 * alpha_mlswe(k) * ((p_face(0, k + 1) + p_inc1) * (p_face(0, k + 1) + p_inc1) - p_face(0, k + 1) * p_face(0, k + 1));
                    H_face(0, iq, k) = H_face(0, iq, k) - H_corr1;
                    H_face(0, iq, k + 1) = H_face(0, iq, k + 1) + H_corr1;
                    double p_inc2 = g_over_alpha[k] * (z_face(1, k + 1) - z_edge_minus[k + 1]);
                    double H_corr2 = 0.5 * alpha_mlswe(k) * ((p_face(1, k + 1) + p_inc2) * (p_face(1, k + 1) + p_inc2) - p_face(1, k + 1) * p_face(1, k + 1));
                    H_face(1, iq, k) = H_face(1, iq, k) - H_corr2;
                    H_face(1, iq, k + 1) = H_face(1, iq, k + 1) + H_corr2;
                }
            }
            for (int s = 0; s < 2; ++s) {
                double weight = 1.0, acceleration = 0.0;
                for (int k = 0; k < nl; ++k) acceleration += H_face(s, iq, k);
                if (acceleration > 0.0) weight = H_face_ave(iq, f) / acceleration;
                for (int k = 0; k < nl; ++k) H_face(s, iq, k) = H_face(s, iq, k) * weight;
            }
        }
        for (int k = 0; k < nl; ++k)
            for (int iq = 0; iq < nq; ++iq) {
                double wq = jac_faceq(iq, f);
                double nxl = normal_vector_q(0, iq, f), nyl = normal_vector_q(1, iq, f);
                double hlx_k = nxl * H_face(0, iq, k), hrx_k = nxl * H_face(1, iq, k);
                double hly_k = nyl * H_face(0, iq, k), hry_k = nyl * H_face(1, iq, k);
                double flux_x = nxl * udp_flux(0, iq, k) + nyl * udp_flux(1, iq, k);
                double flux_y = nxl * vdp_flux(0, iq, k) + nyl * vdp_flux(1, iq, k);
                for (int n = 0; n < ngl; ++n) {
                    double hi = psiq(n, iq);
                    int I = fnodeL[(size_t)f * ngl + n];
                    rhs_mom(0, I, k) = rhs_mom(0, I, k) - wq * hi * (hlx_k + flux_x);
                    rhs_mom(1, I, k) = rhs_mom(1, I, k) - wq * hi * (hly_k + flux_y);
                    if (er > 0) {
                        int Ir = fnodeR[(size_t)f * ngl + n];
                        rhs_mom(0, Ir, k) = rhs_mom(0, Ir, k) + wq * hi * (hrx_k + flux_x);
                        rhs_mom(1, Ir, k) = rhs_mom(1, Ir, k) + wq * hi * (hry_k + flux_y);
                    }
                }
            }
    }
}

// mod_splitting.F90:289-322 + mod_create_rhs_mlswe.F90:28-51
void Oracle::rhs_momentum(Arr& rhs_mom, const Arr& qprime, const Arr& q, const Arr& qf) {
    Arr rhs_visc_bcl; rhs_visc_bcl.alloc(2, npoin, nl);
    if (cfg.method_visc == 1) bcl_create_laplacian_v2(rhs_visc_bcl, qprime);
    else bcl_create_laplacian(rhs_visc_bcl);
    layer_momentum_volume(rhs_mom, qprime, q);
    apply_layers_fluxes(rhs_mom, qf);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) {
            rhs_mom(0, I, k) = massinv(I) * rhs_mom(0, I, k) + rhs_visc_bcl(0, I, k);
            rhs_mom(1, I, k) = massinv(I) * rhs_mom(1, I, k) + rhs_visc_bcl(1, I, k);
        }
}

// mod_layer_terms.F90:529-584.  q is the full (3,npoin,nl) array; components 2:3 are modified.
void Oracle::layer_mom_boundary_df(Arr& q) {
    for (int f = 0; f < nface; ++f) {
        int er = face[8 * f + 7];
        if (er == -4) {
            for (int n = 0; n < ngl; ++n) {
                int I = fnodeL[(size_t)f * ngl + n];
                double nx = normal_vector(0, n, f), ny = normal_vector(1, n, f);
                for (int k = 0; k < nl; ++k) {
                    double upnl = q(1, I, k) * nx + q(2, I, k) * ny;
                    q(1, I, k) = q(1, I, k) - upnl * nx;
                    q(2, I, k) = q(2, I, k) - upnl * ny;
                }
            }
        } else if (er == -2) {
            for (int n = 0; n < ngl; ++n) {
                int I = fnodeL[(size_t)f * ngl + n];
                for (int k = 0; k < nl; ++k) { q(1, I, k) = 0.0; q(2, I, k) = 0.0; }
            }
        }
    }
}

// mod_layer_terms.F90:272-320
void Oracle::extract_velocity(Arr& uv, const Arr& q, const Arr& qb) {
    uv.zero();
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) { uv(0, I, k) = q(1, I, k) / q(0, I, k); uv(1, I, k) = q(2, I, k) / q(0, I, k); }
    for (int I = 0; I < npoin; ++I) {
        double ubar = 0.0, vbar = 0.0;
        for (int k = 0; k < nl; ++k) { ubar = ubar + uv(0, I, k) * q(0, I, k); vbar = vbar + uv(1, I, k) * q(0, I, k); }
        if (qb(0, I) > 0.0) {
            ubar = ubar / qb(0, I); vbar = vbar / qb(0, I);
            for (int k = 0; k < nl; ++k) {
                uv(0, I, k) = uv(0, I, k) - ubar + qb(2, I) / qb(0, I);
                uv(1, I, k) = uv(1, I, k) - vbar + qb(3, I) / qb(0, I);
            }
        } else {
            for (int k = 0; k < nl; ++k) { uv(0, I, k) = 0.0; uv(1, I, k) = 0.0; }
        }
    }
}

// mod_layer_terms.F90:198-238
void Oracle::evaluate_bcl(Arr& qf, Arr& q, Arr& qprime, const Arr& qb) {
    Arr uv; uv.alloc(2, npoin, nl);
    std::vector<double> ope(npoin, 0.0);
    extract_velocity(uv, q, qb);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) {
            q(1, I, k) = uv(0, I, k) * q(0, I, k);
            q(2, I, k) = uv(1, I, k) * q(0, I, k);
            ope[I] = ope[I] + q(0, I, k);
        }
    for (int I = 0; I < npoin; ++I) ope[I] = ope[I] / pbprime_df(I);
    extract_velocity(uv, q, qb);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) {
            qprime(0, I, k) = q(0, I, k) / ope[I];
            qprime(1, I, k) = uv(0, I, k) - qb(2, I) / qb(0, I);
            qprime(2, I, k) = uv(1, I, k) - qb(3, I) / qb(0, I);
        }
    extract_qprime_df_face(qf, qprime);
}

// mod_layer_terms.F90:240-270
void Oracle::evaluate_bcl_v1(Arr& q, Arr& qprime, const Arr& qb) {
    Arr uv; uv.alloc(2, npoin, nl);
    extract_velocity(uv, q, qb);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) { q(1, I, k) = uv(0, I, k) * q(0, I, k); q(2, I, k) = uv(1, I, k) * q(0, I, k); }
    extract_velocity(uv, q, qb);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) {
            qprime(1, I, k) = uv(0, I, k) - qb(2, I) / qb(0, I);
            qprime(2, I, k) = uv(1, I, k) - qb(3, I) / qb(0, I);
        }
}

// mod_layer_terms.F90:139-196: velocities reconciled with the barotropic velocity, momentum rebuilt from them
void Oracle::velocity_df(Arr& q, const Arr& qb) {
    Arr uv; uv.alloc(2, npoin, nl);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) { uv(0, I, k) = q(1, I, k) / q(0, I, k); uv(1, I, k) = q(2, I, k) / q(0, I, k); }
    for (int I = 0; I < npoin; ++I) {
        double ubar = 0.0, vbar = 0.0;
        for (int k = 0; k < nl; ++k) { ubar = ubar + uv(0, I, k) * q(0, I, k); vbar = vbar + uv(1, I, k) * q(0, I, k); }
        if (qb(0, I) > 0.0) {
            ubar = ubar / qb(0, I); vbar = vbar / qb(0, I);
            for (int k = 0; k < nl; ++k) {
                uv(0, I, k) = uv(0, I, k) - ubar + qb(2, I) / qb(0, I);
                uv(1, I, k) = uv(1, I, k) - vbar + qb(3, I) / qb(0, I);
            }
        } else {
            for (int k = 0; k < nl; ++k) { uv(0, I, k) = 0.0; uv(1, I, k) = 0.0; }
        }
    }
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) { q(1, I, k) = uv(0, I, k) * q(0, I, k); q(2, I, k) = uv(1, I, k) * q(0, I, k); }
}

// mod_create_rhs_mlswe.F90:146-279: implicit vertical shear stress between the layers, one tridiagonal system per quadrature
// point; rhs_stress(2,npoin,nl) WITHOUT the inverse mass matrix (the caller applies it, mod_splitting.F90:160-163).
// Statement by statement as written (sub-diagonal -coeff, super-diagonal -coeff1, right-hand side u = udp/dp), except for one
// value the reference never defines: tau_u(nlayers+1), tau_v(nlayers+1) are read at :253-254 but never assigned
// (parity hazard 2): the evident intent -- no shear stress through the bottom, which the bottom drag handles -- is 0.
void Oracle::rhs_layer_shear_stress(Arr& rhs_stress, const Arr& q) {
    rhs_stress.zero();
    const int nq2 = nq * nq;
    const double ad = cfg.ad_mlswe;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < nelem; ++e) {
        std::vector<double> tau_u(nl + 1), tau_v(nl + 1), a(nl), b(nl), c(nl), dp(nl), udp(nl), vdp(nl), r1(nl), r2(nl), u(nl), v(nl);
        for (int Iq = e * nq2; Iq < (e + 1) * nq2; ++Iq) {
            std::fill(dp.begin(), dp.end(), 0.0); std::fill(udp.begin(), udp.end(), 0.0); std::fill(vdp.begin(), vdp.end(), 0.0);
            for (int ip = 0; ip < npts; ++ip) {
                int I = indexq[(size_t)Iq * npts + ip];
                double hi = psih(ip, Iq);
                for (int k = 0; k < nl; ++k) {
                    dp[k] = dp[k] + hi * q(0, I, k);
                    udp[k] = udp[k] + hi * q(1, I, k);
                    vdp[k] = vdp[k] + hi * q(2, I, k);
                }
            }
            double coeff = std::fmax(std::sqrt(0.5 * coriolis_quad(Iq) * ad) / alpha_mlswe(0), ad / (alpha_mlswe(0) * cfg.max_shear_dz));
            double coeff1 = gravity * dt * coeff;
            for (int k = 0; k < nl; ++k) {
                a[k] = -coeff;
                b[k] = dp[k] + 2.0 * coeff1;
                c[k] = -coeff1;
                r1[k] = udp[k] / dp[k];
                r2[k] = vdp[k] / dp[k];
            }
            b[0] = dp[0] + coeff1;
            b[nl - 1] = dp[nl - 1] + coeff1;
            a[0] = 0.0;
            c[nl - 1] = 0.0;
            for (int k = 1; k < nl; ++k) {
                double mult = a[k] / b[k - 1];
                b[k] = b[k] - mult * c[k - 1];
                r1[k] = r1[k] - mult * r1[k - 1];
                r2[k] = r2[k] - mult * r2[k - 1];
            }
            r1[nl - 1] = r1[nl - 1] / b[nl - 1];
            r2[nl - 1] = r2[nl - 1] / b[nl - 1];
            u[nl - 1] = r1[nl - 1]; v[nl - 1] = r2[nl - 1];
            for (int k = nl - 2; k >= 0; --k) {
                r1[k] = (r1[k] - c[k] * r1[k + 1]) / b[k];
                r2[k] = (r2[k] - c[k] * r2[k + 1]) / b[k];
                u[k] = r1[k]; v[k] = r2[k];
            }
            tau_u[0] = 0.0; tau_v[0] = 0.0;
            for (int k = 1; k < nl; ++k) { tau_u[k] = coeff * (u[k - 1] - u[k]); tau_v[k] = coeff * (v[k - 1] - v[k]); }
            tau_u[nl] = 0.0; tau_v[nl] = 0.0;   // hazard 2: undefined in the reference
            double wq = wjac(Iq);
            for (int k = 0; k < nl; ++k) {
                double tau_u_q = gravity * (tau_u[k] - tau_u[k + 1]);
                double tau_v_q = gravity * (tau_v[k] - tau_v[k + 1]);
                for (int ip = 0; ip < npts; ++ip) {
                    int I = indexq[(size_t)Iq * npts + ip];
                    double hi = psih(ip, Iq);
                    rhs_stress(0, I, k) = rhs_stress(0, I, k) + wq * hi * tau_u_q;
                    rhs_stress(1, I, k) = rhs_stress(1, I, k) + wq * hi * tau_v_q;
                }
            }
        }
    }
}

// mod_splitting.F90:139-164 / 247-271: the shear-stress block of momentum / momentum_mass.  q_df_temp(2,npoin,nl) holds
// q_df(2:3) + dt*rhs_mom on entry.  In `momentum` the reference passes an unset local (`uv`, of another shape) to
// rhs_layer_shear_stress (:158) where `momentum_mass` passes q_df3 (:265): parity hazard 3; q_df3 -- the evident intent -- is used in both.
void Oracle::add_shear_stress(Arr& q_df_temp, const Arr& q, const Arr& qb) {
    Arr q3; q3.alloc(3, npoin, nl);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) {
            double tempu = q_df_temp(0, I, k) + fdt2_bcl(I) * q(2, I, k);
            double tempv = q_df_temp(1, I, k) - fdt2_bcl(I) * q(1, I, k);
            q3(1, I, k) = a_bcl(I) * tempu + b_bcl(I) * tempv;
            q3(2, I, k) = -b_bcl(I) * tempu + a_bcl(I) * tempv;
            q3(0, I, k) = q(0, I, k);
        }
    velocity_df(q3, qb);
    Arr rhs_stress; rhs_stress.alloc(2, npoin, nl);
    rhs_layer_shear_stress(rhs_stress, q3);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) {
            q_df_temp(0, I, k) = q_df_temp(0, I, k) + dt * (massinv(I) * rhs_stress(0, I, k));
            q_df_temp(1, I, k) = q_df_temp(1, I, k) + dt * (massinv(I) * rhs_stress(1, I, k));
        }
}

// mod_splitting.F90:182-287
void Oracle::momentum_mass(Arr& q, Arr& qf, Arr& qprime, const Arr& qb) {
    Arr dp_advec; dp_advec.alloc(npoin, nl);
    Arr rhs_mom; rhs_mom.alloc(2, npoin, nl);
    layer_mass_rhs(dp_advec, qprime, qf);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) {
            q(0, I, k) = q(0, I, k) + dt * dp_advec(I, k);
            if (q(0, I, k) < 0.0) error_flag = 1;
        }
    apply_consistency(q);
    rhs_momentum(rhs_mom, qprime, q, qf);
    Arr q_df_temp; q_df_temp.alloc(2, npoin, nl);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) {
            q_df_temp(0, I, k) = q(1, I, k) + dt * rhs_mom(0, I, k);
            q_df_temp(1, I, k) = q(2, I, k) + dt * rhs_mom(1, I, k);
        }
    if (cfg.ad_mlswe > 0.0) add_shear_stress(q_df_temp, q, qb);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) {
            double tempu = q_df_temp(0, I, k) + fdt2_bcl(I) * q(2, I, k);
            double tempv = q_df_temp(1, I, k) - fdt2_bcl(I) * q(1, I, k);
            q(1, I, k) = a_bcl(I) * tempu + b_bcl(I) * tempv;
            q(2, I, k) = -b_bcl(I) * tempu + a_bcl(I) * tempv;
        }
    layer_mom_boundary_df(q);
    evaluate_bcl(qf, q, qprime, qb);
}

// mod_splitting.F90:25-91
void Oracle::thickness(Arr& qprime, Arr& q, const Arr& qb, Arr& qf) {
    (void)qb;
    Arr dp_advec; dp_advec.alloc(npoin, nl);
    layer_mass_rhs(dp_advec, qprime, qf);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) {
            q(0, I, k) = q(0, I, k) + dt * dp_advec(I, k);
            if (q(0, I, k) < 0.0) error_flag = 1;
        }
    apply_consistency(q);
    for (int I = 0; I < npoin; ++I) {
        double s = 0.0;
        for (int k = 0; k < nl; ++k) s += q(0, I, k);
        double ope = s / pbprime_df(I);
        for (int k = 0; k < nl; ++k) qprime(0, I, k) = q(0, I, k) / ope;
    }
    extract_dprime_df_face(qf, qprime);
}

// mod_splitting.F90:94-180
void Oracle::momentum(Arr& q, Arr& qprime, const Arr& qb, const Arr& qf) {
    Arr rhs_mom; rhs_mom.alloc(2, npoin, nl);
    rhs_momentum(rhs_mom, qprime, q, qf);
    Arr q_df_temp; q_df_temp.alloc(2, npoin, nl);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) {
            q_df_temp(0, I, k) = q(1, I, k) + dt * rhs_mom(0, I, k);
            q_df_temp(1, I, k) = q(2, I, k) + dt * rhs_mom(1, I, k);
        }
    if (cfg.ad_mlswe > 0.0) add_shear_stress(q_df_temp, q, qb);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) {
            double tempu = q_df_temp(0, I, k) + fdt2_bcl(I) * q(2, I, k);
            double tempv = q_df_temp(1, I, k) - fdt2_bcl(I) * q(1, I, k);
            q(1, I, k) = a_bcl(I) * tempu + b_bcl(I) * tempv;
            q(2, I, k) = -b_bcl(I) * tempu + a_bcl(I) * tempv;
        }
    layer_mom_boundary_df(q);
    evaluate_bcl_v1(q, qprime, qb);
}

// ti_rk_bcl.F90:9-87
void Oracle::ti_rk_bcl() {
    Arr qf, qf2; qf.alloc(3, 2, ngl, nface, nl);
    Arr qbp, qprime2, q2, dpprime_df2;
    // ---- prediction
    extract_qprime_df_face(qf, qprime_df);
    qbp = qb_df;
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) dpprime_visc(I, k) = qprime_df(0, I, k);
    if (cfg.method_visc == 1) interpolate_dpp();
    btp_bcl_coeffs_qdf(qf, qprime_df);
    ti_barotropic_ssprk_mlswe(qbp, qprime_df);
    q2 = q_df; qprime2 = qprime_df; qf2 = qf;
    momentum_mass(q2, qf2, qprime2, qbp);
    // ---- correction
    for (size_t i = 0; i < qprime2.size(); ++i) qprime2.v[i] = 0.5 * (qprime2.v[i] + qprime_df.v[i]);
    for (size_t i = 0; i < qf2.size(); ++i) qf2.v[i] = 0.5 * (qf.v[i] + qf2.v[i]);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) dpprime_visc(I, k) = qprime2(0, I, k);
    if (cfg.method_visc == 1) interpolate_dpp();
    btp_bcl_coeffs_qdf(qf2, qprime2);
    ti_barotropic_ssprk_mlswe(qb_df, qprime2);
    thickness(qprime2, q_df, qb_df, qf2);
    dpprime_df2.alloc(npoin, nl);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) {
            dpprime_df2(I, k) = qprime2(0, I, k);
            qprime2(0, I, k) = 0.5 * (qprime_df(0, I, k) + dpprime_df2(I, k));
        }
    for (int k = 0; k < nl; ++k)
        for (int f = 0; f < nface; ++f)
            for (int n = 0; n < ngl; ++n)
                for (int s = 0; s < 2; ++s) qf2(0, s, n, f, k) = 0.5 * (qf(0, s, n, f, k) + qf2(0, s, n, f, k));
    momentum(q_df, qprime2, qb_df, qf2);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) {
            qprime_df(0, I, k) = dpprime_df2(I, k);
            qprime_df(1, I, k) = qprime2(1, I, k);
            qprime_df(2, I, k) = qprime2(2, I, k);
        }
}

// diagnostics.F90:24-45: q(1)=h, q(2)=u, q(3)=v, q(4)=dp, q(5)=interface elevation ("ssh")
void Oracle::diagnostics(Arr& qout) const {
    qout.alloc(5, npoin, nl);
    Arr elev; elev.alloc(npoin, nl + 1);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) {
            qout(0, I, k) = (alpha_mlswe(k) / gravity) * q_df(0, I, k);
            qout(1, I, k) = q_df(1, I, k) / q_df(0, I, k);
            qout(2, I, k) = q_df(2, I, k) / q_df(0, I, k);
            qout(3, I, k) = q_df(0, I, k);
        }
    for (int I = 0; I < npoin; ++I) elev(I, nl) = zbot_df(I);
    for (int k = nl - 1; k >= 0; --k)
        for (int I = 0; I < npoin; ++I) elev(I, k) = elev(I, k + 1) + qout(0, I, k);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) qout(4, I, k) = elev(I, k);
}

// compute_conserved.F90:29-41 applied to q(1,:) = h of layer k
double Oracle::layer_mass(const Arr& qout, int k) const {
    double mass = 0.0;
    for (int Iq = 0; Iq < npoin; ++Iq) {
        double wq = wjac_df(Iq);
        for (int ip = 0; ip < npts; ++ip) {
            int I = index_df[(size_t)Iq * npts + ip];
            mass = mass + wq * psih_df(ip, Iq) * qout(0, I, k);
        }
    }
    return mass;
}

}  // namespace orc
