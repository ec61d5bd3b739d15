// ============================================================================
// oracle_setup.cpp -- TEST INFRASTRUCTURE ONLY (CPU oracle, set-up half).
// Restates the one-time set-up the Fortran driver performs before the time loop
// (reference src/amain.F90:31-34): basis, brick grid, metrics, faces, dense
// tensor-product tables, initial conditions and static coefficients.
// Parity status: see hnumo_oracle.hpp (weakly pinned).
// ============================================================================
#include "hnumo_oracle.hpp"

#include <algorithm>
#include <cstdlib>

namespace orc {

static const double PI_TRIG = 3.1415926535897932346;  // mod_constants.F90:23

// mod_legendre.F90:189-236 (legendre_poly_loc): P_n, P'_n, P''_n by recurrence.
static void legendre_poly(int n, double x, double& p0, double& p0_1, double& p0_2) {
    double p2 = 0, p2_1 = 0, p2_2 = 0, p1 = 0, p1_1 = 0, p1_2 = 0;
    p0 = 1.0; p0_1 = 0.0; p0_2 = 0.0;
    for (int j = 1; j <= n; ++j) {
        p2 = p1; p2_1 = p1_1; p2_2 = p1_2;
        p1 = p0; p1_1 = p0_1; p1_2 = p0_2;
        double a = (2.0 * j - 1.0) / j, b = (j - 1.0) / j;
        p0 = a * x * p1 - b * p2;
        p0_1 = a * (p1 + x * p1_1) - b * p2_1;
        p0_2 = a * (2.0 * p1_1 + x * p1_2) - b * p2_2;
    }
}

// mod_legendre.F90:54-111 (legendre_gauss_lobatto)
static void legendre_gauss_lobatto(int ngl, std::vector<double>& xgl, std::vector<double>& wgl) {
    xgl.assign(ngl, 0.0); wgl.assign(ngl, 0.0);
    if (ngl == 1) { xgl[0] = 0; wgl[0] = 2; return; }
    const double thres = 2.220446049250313e-16;
    double pi = 4.0 * std::atan(1.0);
    int n = ngl - 1, nh = (n + 1) / 2;
    for (int i = 1; i <= nh; ++i) {
        double x = std::cos((2.0 * i - 1.0) / (2.0 * n + 1.0) * pi);
        double p0 = 0, p0_1 = 0, p0_2 = 0;
        for (int k = 1; k <= 20; ++k) {
            legendre_poly(n, x, p0, p0_1, p0_2);
            double dx = -(1.0 - x * x) * p0_1 / (-2.0 * x * p0_1 + (1.0 - x * x) * p0_2);
            x = x + dx;
            if (std::fabs(dx) < thres) break;
        }
        // note: as in the reference, p0 is the value from the last evaluation (before the final update)
        xgl[n + 1 - i] = x;
        wgl[n + 1 - i] = 2.0 / ((double)(n * (n + 1)) * p0 * p0);
    }
    if (n + 1 != 2 * nh) {
        double p0, a, b;
        legendre_poly(n, 0.0, p0, a, b);
        xgl[nh] = 0.0;
        wgl[nh] = 2.0 / ((double)(n * (n + 1)) * p0 * p0);
    }
    for (int i = 1; i <= nh; ++i) {
        xgl[i - 1] = -xgl[n + 1 - i];
        wgl[i - 1] = wgl[n + 1 - i];
    }
}

// mod_legendre.F90:248-322 (legendre_basis, reduce_round_off=.true. branch)
static void legendre_basis(int ngl, const std::vector<double>& xgl, Arr& psi, Arr& dpsi) {
    psi.alloc(ngl, ngl); dpsi.alloc(ngl, ngl);
    std::vector<double> bb(ngl, 0.0), cc(ngl, 0.0);
    for (int j = 0; j < ngl; ++j)
        for (int i = 0; i < ngl; ++i) {
            if (i == j) psi(i, j) = 1.0;
            else { bb[j] += std::log(std::fabs(xgl[j] - xgl[i])); psi(i, j) = 0.0; }
        }
    for (int j = 0; j < ngl; ++j)
        for (int i = 0; i < ngl; ++i)
            if (i != j) {
                double sgn = ((i + j) % 2 == 0) ? 1.0 : -1.0;
                dpsi(i, j) = sgn * std::exp(bb[j] - bb[i]) / (xgl[j] - xgl[i]);
                cc[j] += dpsi(i, j);
            }
    for (int j = 0; j < ngl; ++j) dpsi(j, j) = -cc[j];
}

// mod_legendre.F90:387-433 (lagrange_basis)
static void lagrange_basis(int ngl, const std::vector<double>& xgl, int nq, std::vector<double>& xnq,
                           std::vector<double>& wnq, Arr& psiq, Arr& dpsiq) {
    legendre_gauss_lobatto(nq, xnq, wnq);
    psiq.alloc(ngl, nq); dpsiq.alloc(ngl, nq);
    for (int l = 0; l < nq; ++l) {
        double xl = xnq[l];
        for (int i = 0; i < ngl; ++i) {
            double ksi = xgl[i];
            psiq(i, l) = 1.0; dpsiq(i, l) = 0.0;
            for (int j = 0; j < ngl; ++j) {
                double xj = xgl[j];
                if (j != i) psiq(i, l) = psiq(i, l) * (xl - xj) / (ksi - xj);
                double ddpsi = 1.0;
                if (j != i) {
                    for (int k = 0; k < ngl; ++k) {
                        double xk = xgl[k];
                        if (k != i && k != j) ddpsi = ddpsi * (xl - xk) / (ksi - xk);
                    }
                    dpsiq(i, l) = dpsiq(i, l) + ddpsi / (ksi - xj);
                }
            }
        }
    }
}

// mod_basis.F90:60-186
void Oracle::build_basis() {
    ngl = cfg.nop + 1;
    nq = cfg.dg_integ_exact ? 2 * cfg.nop + 1 : 2 * cfg.nop - 1;
    npts = ngl * ngl;
    legendre_gauss_lobatto(ngl, xgl, wgl);
    legendre_basis(ngl, xgl, psi, dpsi);
    lagrange_basis(ngl, xgl, nq, xnq, wnq, psiq, dpsiq);
}

// Brick of nelx x nely trees, one element each (p4est.c:157-242 fill_coordinates,
// mod_p4est.F90:344-359 rescale; intma_dg mod_grid.F90:230-237).  Element order here is
// row-major (e = ex + nelx*ey); the reference uses p4est's space-filling order, which only
// changes the order of rhs scatter sums.
void Oracle::build_grid() {
    nelem = cfg.nelx * cfg.nely;
    npoin = nelem * npts;
    npoin_q = nelem * nq * nq;
    coord.alloc(2, npoin);
    for (int ey = 0; ey < cfg.nely; ++ey)
        for (int ex = 0; ex < cfg.nelx; ++ex) {
            int e = ex + cfg.nelx * ey;
            double vx[4] = {(double)ex, (double)ex + 1, (double)ex, (double)ex + 1};
            double vy[4] = {(double)ey, (double)ey, (double)ey + 1, (double)ey + 1};
            for (int m = 0; m < ngl; ++m)
                for (int l = 0; l < ngl; ++l) {
                    double w[4] = {(1 - xgl[l]) * (1 - xgl[m]), (1 + xgl[l]) * (1 - xgl[m]),
                                   (1 - xgl[l]) * (1 + xgl[m]), (1 + xgl[l]) * (1 + xgl[m])};
                    double tx = 0, ty = 0;
                    for (int c = 0; c < 4; ++c) { tx += w[c] * vx[c]; ty += w[c] * vy[c]; }
                    tx /= 4; ty /= 4;
                    int I = e * npts + m * ngl + l;
                    double x = tx / cfg.nelx, y = ty / cfg.nely;
                    coord(0, I) = x * (cfg.xdims[1] - cfg.xdims[0]) + cfg.xdims[0];
                    coord(1, I) = y * (cfg.ydims[1] - cfg.ydims[0]) + cfg.ydims[0];
                }
        }
    if (cfg.mesh_warp != 0.0) {   // test aid, see Config::mesh_warp
        const double pi = PI_TRIG, Lx = cfg.xdims[1] - cfg.xdims[0], Ly = cfg.ydims[1] - cfg.ydims[0];
        const double kx = std::max(1, cfg.nelx / 2), ky = std::max(1, cfg.nely / 2);
        for (int I = 0; I < npoin; ++I) {
            const double X = (coord(0, I) - cfg.xdims[0]) / Lx, Y = (coord(1, I) - cfg.ydims[0]) / Ly;
            const double ddx = cfg.mesh_warp * (Lx / cfg.nelx) * std::sin(pi * kx * X) * std::cos(pi * ky * Y);
            const double ddy = cfg.mesh_warp * (Ly / cfg.nely) * std::cos(pi * kx * X) * std::sin(pi * ky * Y);
            coord(0, I) += ddx; coord(1, I) += ddy;
        }
    }
    // faces: p4est.c:1560-1735.  p4est face f=0..3 (-x,+x,-y,+y) -> numa local face 5,6,3,4;
    // left element = lower element number; boundary: face(8) = -bc flag.
    static const int transform[4] = {4, 5, 2, 3};
    face.clear();
    auto push_face = [&](int ilocl, int ilocr, int el, int er) {
        int f[8] = {0, 0, 0, 0, ilocl, ilocr, el, er};
        face.insert(face.end(), f, f + 8);
    };
    for (int ey = 0; ey < cfg.nely; ++ey)
        for (int ex = 0; ex < cfg.nelx; ++ex) {
            int q = ex + cfg.nelx * ey;
            for (int f = 0; f < 4; ++f) {
                int nx = ex + (f == 0 ? -1 : f == 1 ? 1 : 0), ny = ey + (f == 2 ? -1 : f == 3 ? 1 : 0);
                bool bdy = nx < 0 || nx >= cfg.nelx || ny < 0 || ny >= cfg.nely;
                if (bdy) {
                    int bc = (f == 0) ? cfg.x_boundary[0] : (f == 1) ? cfg.x_boundary[1]
                           : (f == 2) ? cfg.y_boundary[0] : cfg.y_boundary[1];
                    push_face(transform[f] + 1, 0, q + 1, -bc);
                } else {
                    int nq_ = nx + cfg.nelx * ny;
                    int nf = f ^ 1;
                    if (q < nq_) push_face(transform[f] + 1, transform[nf] + 1, q + 1, nq_ + 1);
                }
            }
        }
    nface = (int)face.size() / 8;
}

// metrics.F90:40-127, metrics_quad.F90:20-127, create_mass.F90:24-31, mod_metrics.F90:112
void Oracle::build_metrics() {
    ksi_x.alloc(npoin); ksi_y.alloc(npoin); eta_x.alloc(npoin); eta_y.alloc(npoin); jac.alloc(npoin);
    ksiq_x.alloc(npoin_q); ksiq_y.alloc(npoin_q); etaq_x.alloc(npoin_q); etaq_y.alloc(npoin_q); jacq.alloc(npoin_q);
    massinv.alloc(npoin);
    for (int e = 0; e < nelem; ++e) {
        // nodal metrics: x_ksi(i,j) = sum_n dpsi(n,i) x(n,j)  (mod_gradient.F90:110-166)
        for (int j = 0; j < ngl; ++j)
            for (int i = 0; i < ngl; ++i) {
                double x_ksi = 0, x_eta = 0, y_ksi = 0, y_eta = 0;
                for (int n = 0; n < ngl; ++n) {
                    x_ksi += dpsi(n, i) * coord(0, e * npts + j * ngl + n);
                    y_ksi += dpsi(n, i) * coord(1, e * npts + j * ngl + n);
                    x_eta += coord(0, e * npts + n * ngl + i) * dpsi(n, j);
                    y_eta += coord(1, e * npts + n * ngl + i) * dpsi(n, j);
                }
                double z_zeta = 1.0;
                double xj = (x_ksi * y_eta * z_zeta) - (y_ksi * x_eta * z_zeta);
                int I = e * npts + j * ngl + i;
                ksi_x(I) = (y_eta * z_zeta) / xj;
                ksi_y(I) = -(x_eta * z_zeta) / xj;
                eta_x(I) = -(y_ksi * z_zeta) / xj;
                eta_y(I) = (x_ksi * z_zeta) / xj;
                jac(I) = wgl[i] * wgl[j] * 1.0 * std::fabs(xj);
            }
        // quadrature metrics (compute_local_gradient_quad_v3, mod_gradient.F90:175-260)
        for (int jq = 0; jq < nq; ++jq)
            for (int iq = 0; iq < nq; ++iq) {
                double x_ksi = 0, x_eta = 0, y_ksi = 0, y_eta = 0;
                for (int m = 0; m < ngl; ++m)
                    for (int n = 0; n < ngl; ++n) {
                        double hix = dpsiq(n, iq) * psiq(m, jq);
                        double hiy = psiq(n, iq) * dpsiq(m, jq);
                        double x = coord(0, e * npts + m * ngl + n), y = coord(1, e * npts + m * ngl + n);
                        x_ksi += hix * x; y_ksi += hix * y;
                        x_eta += hiy * x; y_eta += hiy * y;
                    }
                double xj = x_ksi * y_eta - y_ksi * x_eta;
                int Iq = e * nq * nq + jq * nq + iq;
                ksiq_x(Iq) = y_eta / xj;
                ksiq_y(Iq) = -x_eta / xj;
                etaq_x(Iq) = -y_ksi / xj;
                etaq_y(Iq) = x_ksi / xj;
                jacq(Iq) = wnq[iq] * wnq[jq] * 1.0 * std::fabs(xj);
            }
    }
    for (int I = 0; I < npoin; ++I) massinv(I) = 1.0 / jac(I);  // DG: mass(ip)=jac, massinv=1/mass
}

// create_normals.F90:17-215, create_normals_quad.F90, create_imaplr create_normals.F90:229-397
void Oracle::build_faces() {
    normal_vector.alloc(2, ngl, nface); jac_face.alloc(ngl, nface);
    normal_vector_q.alloc(2, nq, nface); jac_faceq.alloc(nq, nface);
    fnodeL.assign((size_t)ngl * nface, -1); fnodeR.assign((size_t)ngl * nface, -1);
    auto local_node = [&](int iloc, int l, int& i, int& j) {
        switch (iloc) {
            case 3: i = l; j = 0; break;
            case 4: i = l; j = ngl - 1; break;
            case 5: i = 0; j = l; break;
            default: i = ngl - 1; j = l; break;  // 6
        }
    };
    for (int f = 0; f < nface; ++f) {
        int ilocl = face[8 * f + 4], ilocr = face[8 * f + 5], iel = face[8 * f + 6] - 1, ier = face[8 * f + 7];
        // nodal normals
        for (int l = 0; l < ngl; ++l) {
            int i, j; local_node(ilocl, l, i, j);
            double x_ksi = 0, x_eta = 0, y_ksi = 0, y_eta = 0;
            for (int n = 0; n < ngl; ++n) {
                x_ksi += dpsi(n, i) * coord(0, iel * npts + j * ngl + n);
                y_ksi += dpsi(n, i) * coord(1, iel * npts + j * ngl + n);
                x_eta += coord(0, iel * npts + n * ngl + i) * dpsi(n, j);
                y_eta += coord(1, iel * npts + n * ngl + i) * dpsi(n, j);
            }
            double nx, ny, z_zeta = 1.0;
            switch (ilocl) {
                case 3: nx = +y_ksi * z_zeta; ny = -x_ksi * z_zeta; break;
                case 4: nx = -y_ksi * z_zeta; ny = +x_ksi * z_zeta; break;
                case 5: nx = -y_eta * z_zeta; ny = +x_eta * z_zeta; break;
                default: nx = +y_eta * z_zeta; ny = -x_eta * z_zeta; break;
            }
            double nlen = std::sqrt(nx * nx + ny * ny);
            jac_face(l, f) = wgl[l] * 1.0 * nlen;
            normal_vector(0, l, f) = nx / nlen;
            normal_vector(1, l, f) = ny / nlen;
            fnodeL[(size_t)f * ngl + l] = iel * npts + j * ngl + i;
            if (ier > 0) {
                int ir, jr; local_node(ilocr, l, ir, jr);
                fnodeR[(size_t)f * ngl + l] = (ier - 1) * npts + jr * ngl + ir;
            }
        }
        // quadrature normals: the quadrature grid of the left element on the face
        for (int l = 0; l < nq; ++l) {
            int iq, jq;
            switch (ilocl) {
                case 3: iq = l; jq = 0; break;
                case 4: iq = l; jq = nq - 1; break;
                case 5: iq = 0; jq = l; break;
                default: iq = nq - 1; jq = l; break;
            }
            double x_ksi = 0, x_eta = 0, y_ksi = 0, y_eta = 0;
            for (int m = 0; m < ngl; ++m)
                for (int n = 0; n < ngl; ++n) {
                    double hix = dpsiq(n, iq) * psiq(m, jq), hiy = psiq(n, iq) * dpsiq(m, jq);
                    double x = coord(0, iel * npts + m * ngl + n), y = coord(1, iel * npts + m * ngl + n);
                    x_ksi += hix * x; y_ksi += hix * y; x_eta += hiy * x; y_eta += hiy * y;
                }
            double nx, ny;
            switch (ilocl) {
                case 3: nx = +y_ksi; ny = -x_ksi; break;
                case 4: nx = -y_ksi; ny = +x_ksi; break;
                case 5: nx = -y_eta; ny = +x_eta; break;
                default: nx = +y_eta; ny = -x_eta; break;
            }
            double nlen = std::sqrt(nx * nx + ny * ny);
            jac_faceq(l, f) = wnq[l] * 1.0 * nlen;
            normal_vector_q(0, l, f) = nx / nlen;
            normal_vector_q(1, l, f) = ny / nlen;
        }
    }
}

// Tensor_product.F90:50-125
void Oracle::build_tensor_tables() {
    psih.alloc(npts, npoin_q); dpsidx.alloc(npts, npoin_q); dpsidy.alloc(npts, npoin_q); wjac.alloc(npoin_q);
    psih_df.alloc(npts, npoin); dpsidx_df.alloc(npts, npoin); dpsidy_df.alloc(npts, npoin); wjac_df.alloc(npoin);
    indexq.assign((size_t)npts * npoin_q, 0); index_df.assign((size_t)npts * npoin, 0);
    for (int e = 0; e < nelem; ++e) {
        for (int jq = 0; jq < nq; ++jq)
            for (int iq = 0; iq < nq; ++iq) {
                int Iq = e * nq * nq + jq * nq + iq;
                wjac(Iq) = jacq(Iq);
                double e_x = ksiq_x(Iq), e_y = ksiq_y(Iq), n_x = etaq_x(Iq), n_y = etaq_y(Iq);
                int ip = 0;
                for (int m = 0; m < ngl; ++m)
                    for (int n = 0; n < ngl; ++n, ++ip) {
                        indexq[(size_t)Iq * npts + ip] = e * npts + m * ngl + n;
                        psih(ip, Iq) = psiq(n, iq) * psiq(m, jq);
                        double h_e = dpsiq(n, iq) * psiq(m, jq);
                        double h_n = psiq(n, iq) * dpsiq(m, jq);
                        dpsidx(ip, Iq) = h_e * e_x + h_n * n_x;
                        dpsidy(ip, Iq) = h_e * e_y + h_n * n_y;
                    }
            }
        for (int jq = 0; jq < ngl; ++jq)
            for (int iq = 0; iq < ngl; ++iq) {
                int Iq = e * npts + jq * ngl + iq;
                wjac_df(Iq) = jac(Iq);
                double e_x = ksi_x(Iq), e_y = ksi_y(Iq), n_x = eta_x(Iq), n_y = eta_y(Iq);
                int ip = 0;
                for (int m = 0; m < ngl; ++m)
                    for (int n = 0; n < ngl; ++n, ++ip) {
                        index_df[(size_t)Iq * npts + ip] = e * npts + m * ngl + n;
                        psih_df(ip, Iq) = psi(n, iq) * psi(m, jq);
                        double h_e = dpsi(n, iq) * psi(m, jq);
                        double h_n = psi(n, iq) * dpsi(m, jq);
                        dpsidx_df(ip, Iq) = h_e * e_x + h_n * n_x;
                        dpsidy_df(ip, Iq) = h_e * e_y + h_n * n_y;
                    }
            }
    }
}

// mod_initial.F90:157-183 -> initial_conditions.F90:86-416, mod_initial_mlswe.F90
void Oracle::build_initial() {
    nl = cfg.nlayers;
    gravity = 9.806;  // initial_conditions.F90:97,132,173; mod_initial_mlswe.F90:306
    const double pi = PI_TRIG;
    alpha_mlswe.alloc(nl);
    zbot_df.alloc(npoin); z_interface.alloc(npoin, nl + 1); tau_wind_df.alloc(2, npoin);
    pbprime_df.alloc(npoin);
    q_df.alloc(3, npoin, nl); qprime_df.alloc(3, npoin, nl); qb_df.alloc(4, npoin);

    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    for (int I = 0; I < npoin; ++I) {
        xmin = std::min(xmin, coord(0, I)); xmax = std::max(xmax, coord(0, I));
        ymin = std::min(ymin, coord(1, I)); ymax = std::max(ymax, coord(1, I));
    }
    double Ly = cfg.ydims[1] - cfg.ydims[0];
    double Lx = cfg.xdims[1] - cfg.xdims[0];

    switch (cfg.test_case) {
        case TC_BUMP: {  // initial_conditions.F90:95-128
            double H_bot = 40.0;
            for (int I = 0; I < npoin; ++I) zbot_df(I) = -H_bot;
            for (int k = 0; k <= nl; ++k)
                for (int I = 0; I < npoin; ++I) z_interface(I, k) = -(double)k * H_bot / (double)nl;
            double xm = 0.5 * (xmax + xmin), yl = 0.5 * (ymax + ymin), L = 250.0, amp = 1.0;
            for (int I = 0; I < npoin; ++I) {
                double x = coord(0, I), y = coord(1, I);
                double r = std::sqrt((x - xm) * (x - xm) + (y - yl) * (y - yl));
                if (r < L) z_interface(I, 1) = z_interface(I, 1) + 0.5 * amp * (1.0 + std::cos(pi * r / L));
            }
            alpha_mlswe(0) = 0.9737e-3;
            if (nl > 1) alpha_mlswe(1) = 0.9735e-3;
            for (int k = 2; k < nl; ++k) alpha_mlswe(k) = 0.9735e-3;  // reference leaves k>2 unset
        } break;
        case TC_LAKE: {  // initial_conditions.F90:130-169
            double H_bot = 40.0;
            for (int I = 0; I < npoin; ++I) zbot_df(I) = -H_bot;
            double xm = 0.5 * (cfg.xdims[0] + cfg.xdims[1]), yl = 0.5 * (cfg.ydims[0] + cfg.ydims[1]), L = 250.0;
            for (int I = 0; I < npoin; ++I) {
                double x = coord(0, I), y = coord(1, I);
                double r = std::sqrt((x - xm) * (x - xm) + (y - yl) * (y - yl));
                if (r < L) zbot_df(I) = zbot_df(I) + 3.0 * (1.0 + std::cos(pi * r / L));
            }
            for (int k = 0; k <= nl; ++k) {
                for (int I = 0; I < npoin; ++I) {
                    if (nl < 5) z_interface(I, k) = -(double)k * H_bot / (double)nl;
                    else { z_interface(I, k) = -(double)k * 32 / (double)(nl - 1); z_interface(I, nl) = -H_bot; }
                }
            }
            double rho_0 = 1027.01037;
            alpha_mlswe(0) = 1.0 / rho_0;
            for (int k = 2; k <= nl; ++k) alpha_mlswe(k - 1) = 1.0 / (rho_0 + k * 0.2110 / (double)nl);
        } break;
        case TC_DOUBLE_GYRE: {  // initial_conditions.F90:171-191 (2 layers)
            double H_bot = 9928.0;
            for (int I = 0; I < npoin; ++I) {
                zbot_df(I) = -H_bot;
                z_interface(I, 1) = -1489.5;
                z_interface(I, 2) = -H_bot;
                tau_wind_df(0, I) = -0.1 * std::cos(2.0 * pi * coord(1, I) / Ly);
            }
            alpha_mlswe(0) = 9.7370e-04; alpha_mlswe(1) = 9.7350e-04;
        } break;
        case TC_DOUBLE_GYRE_SYNTH: {  // explicit nl-layer extension; not in the reference
            double H_bot = -cfg.synth_z[nl];
            for (int I = 0; I < npoin; ++I) {
                zbot_df(I) = -H_bot;
                for (int k = 0; k <= nl; ++k) z_interface(I, k) = cfg.synth_z[k];
                tau_wind_df(0, I) = -0.1 * std::cos(2.0 * pi * coord(1, I) / Ly);
            }
            for (int k = 0; k < nl; ++k) alpha_mlswe(k) = cfg.synth_alpha[k];
        } break;
        default: std::fprintf(stderr, "oracle: unknown test case\n"); std::abort();
    }
    // clip interfaces (initial_conditions.F90:310-317)
    for (int I = 0; I < npoin; ++I)
        for (int k = 0; k <= nl; ++k) z_interface(I, k) = std::max(zbot_df(I), z_interface(I, k));
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I)
            pbprime_df(I) = pbprime_df(I) + (gravity / alpha_mlswe(k)) * (z_interface(I, k) - z_interface(I, k + 1));

    // interpolate_pbprime_init (mod_initial_mlswe.F90:170-262)
    pbprime.alloc(npoin_q); pbprime_face.alloc(2, nq, nface); pbprime_df_face.alloc(2, ngl, nface);
    for (int e = 0; e < nelem; ++e)
        for (int jq = 0; jq < nq; ++jq)
            for (int iq = 0; iq < nq; ++iq) {
                int Iq = e * nq * nq + jq * nq + iq;
                for (int m = 0; m < ngl; ++m)
                    for (int n = 0; n < ngl; ++n)
                        pbprime(Iq) = pbprime(Iq) + pbprime_df(e * npts + m * ngl + n) * (psiq(n, iq) * psiq(m, jq));
            }
    auto face_quad_index = [&](int iloc, int el0, int l) {
        int iq, jq;
        switch (iloc) {
            case 3: iq = l; jq = 0; break;
            case 4: iq = l; jq = nq - 1; break;
            case 5: iq = 0; jq = l; break;
            default: iq = nq - 1; jq = l; break;
        }
        return el0 * nq * nq + jq * nq + iq;
    };
    for (int f = 0; f < nface; ++f) {
        int ilocl = face[8 * f + 4], ilocr = face[8 * f + 5], el = face[8 * f + 6] - 1, er = face[8 * f + 7];
        for (int iq = 0; iq < nq; ++iq) {
            pbprime_face(0, iq, f) = pbprime(face_quad_index(ilocl, el, iq));
            if (er > 0) pbprime_face(1, iq, f) = pbprime(face_quad_index(ilocr, er - 1, iq));
            else pbprime_face(1, iq, f) = pbprime_face(0, iq, f);
        }
        for (int n = 0; n < ngl; ++n) {
            pbprime_df_face(0, n, f) = pbprime_df(fnodeL[(size_t)f * ngl + n]);
            if (er > 0) pbprime_df_face(1, n, f) = pbprime_df(fnodeR[(size_t)f * ngl + n]);
            else pbprime_df_face(1, n, f) = pbprime_df_face(0, n, f);
        }
    }
    // reciprocals with >0 guards (initial_conditions.F90:337-372)
    pbprime_edge.alloc(nq, nface); one_over_pbprime_edge.alloc(nq, nface); one_over_pbprime_face.alloc(2, nq, nface);
    one_over_pbprime_df_face.alloc(2, ngl, nface); one_over_pbprime_df.alloc(npoin); one_over_pbprime.alloc(npoin_q);
    for (int f = 0; f < nface; ++f) {
        for (int iq = 0; iq < nq; ++iq) {
            pbprime_edge(iq, f) = pbprime_face(0, iq, f);
            if (pbprime_edge(iq, f) > 0.0) one_over_pbprime_edge(iq, f) = 1.0 / pbprime_edge(iq, f);
            for (int s = 0; s < 2; ++s)
                if (pbprime_face(s, iq, f) > 0.0) one_over_pbprime_face(s, iq, f) = 1.0 / pbprime_face(s, iq, f);
        }
        for (int n = 0; n < ngl; ++n)
            for (int s = 0; s < 2; ++s)
                if (pbprime_df_face(s, n, f) > 0.0) one_over_pbprime_df_face(s, n, f) = 1.0 / pbprime_df_face(s, n, f);
    }
    for (int I = 0; I < npoin; ++I) if (pbprime_df(I) > 0.0) one_over_pbprime_df(I) = 1.0 / pbprime_df(I);
    for (int Iq = 0; Iq < npoin_q; ++Iq) if (pbprime(Iq) > 0.0) one_over_pbprime(Iq) = 1.0 / pbprime(Iq);

    // layer and barotropic state (initial_conditions.F90:374-416); one_plus_eta_temp starts at 0 (hazard 5)
    std::vector<double> ope(npoin, 0.0);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) {
            q_df(0, I, k) = (gravity / alpha_mlswe(k)) * (z_interface(I, k) - z_interface(I, k + 1));
            ope[I] = ope[I] + q_df(0, I, k) / pbprime_df(I);
        }
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) qprime_df(0, I, k) = q_df(0, I, k) / ope[I];
    if (cfg.test_case == TC_DOUBLE_GYRE_SYNTH && cfg.synth_perturb != 0.0) {
        // optional deterministic perturbation of the layer thickness; column total preserved
        for (int I = 0; I < npoin; ++I) {
            double s = cfg.synth_perturb * std::sin(2.0 * pi * coord(0, I) / Lx) * std::sin(2.0 * pi * coord(1, I) / Ly);
            double d = q_df(0, I, 0) * s;
            q_df(0, I, 0) += d; q_df(0, I, 1) -= d;
            qprime_df(0, I, 0) = q_df(0, I, 0) / ope[I]; qprime_df(0, I, 1) = q_df(0, I, 1) / ope[I];
        }
    }
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) { q_df(1, I, k) = 0.0 * q_df(0, I, k); q_df(2, I, k) = 0.0 * q_df(0, I, k); }
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) {
            qb_df(0, I) += q_df(0, I, k); qb_df(2, I) += q_df(1, I, k); qb_df(3, I) += q_df(2, I, k);
        }
    for (int I = 0; I < npoin; ++I) qb_df(1, I) = qb_df(0, I) - pbprime_df(I);
    for (int k = 0; k < nl; ++k)
        for (int I = 0; I < npoin; ++I) {
            qprime_df(1, I, k) = q_df(1, I, k) / q_df(0, I, k) - qb_df(2, I) / qb_df(0, I);
            qprime_df(2, I, k) = q_df(2, I, k) / q_df(0, I, k) - qb_df(3, I) / qb_df(0, I);
        }

    // compute_reference_edge_variables (mod_initial_mlswe.F90:355-401)
    coeff_pbpert_L.alloc(nq, nface); coeff_pbpert_R.alloc(nq, nface); coeff_pbub_LR.alloc(nq, nface);
    coeff_mass_pbub_L.alloc(nq, nface); coeff_mass_pbub_R.alloc(nq, nface); coeff_mass_pbpert_LR.alloc(nq, nface);
    for (int f = 0; f < nface; ++f)
        for (int iq = 0; iq < nq; ++iq) {
            double c_minus = std::sqrt(alpha_mlswe(nl - 1) * pbprime_face(1, iq, f));
            double c_plus = std::sqrt(alpha_mlswe(nl - 1) * pbprime_face(0, iq, f));
            if (c_minus > 0.0 || c_plus > 0.0) {
                coeff_pbpert_L(iq, f) = c_minus / (c_minus + c_plus);
                coeff_pbpert_R(iq, f) = c_plus / (c_minus + c_plus);
                coeff_pbub_LR(iq, f) = 1.0 / (c_minus + c_plus);
                coeff_mass_pbub_L(iq, f) = c_plus / (c_minus + c_plus);
                coeff_mass_pbub_R(iq, f) = c_minus / (c_minus + c_plus);
                coeff_mass_pbpert_LR(iq, f) = c_minus * c_plus / (c_minus + c_plus);
            }
        }
    // bot_topo_derivatives (mod_initial_mlswe.F90:29-120); zbot starts at 0 (hazard 5)
    zbot.alloc(npoin_q); zbot_face.alloc(2, nq, nface);
    for (int e = 0; e < nelem; ++e)
        for (int jq = 0; jq < nq; ++jq)
            for (int iq = 0; iq < nq; ++iq) {
                int Iq = e * nq * nq + jq * nq + iq;
                for (int m = 0; m < ngl; ++m)
                    for (int n = 0; n < ngl; ++n)
                        zbot(Iq) = zbot(Iq) + zbot_df(e * npts + m * ngl + n) * (psiq(n, iq) * psiq(m, jq));
            }
    for (int f = 0; f < nface; ++f) {
        int ilocl = face[8 * f + 4], ilocr = face[8 * f + 5], el = face[8 * f + 6] - 1, er = face[8 * f + 7];
        for (int iq = 0; iq < nq; ++iq) {
            zbot_face(0, iq, f) = zbot(face_quad_index(ilocl, el, iq));
            if (er > 0) zbot_face(1, iq, f) = zbot(face_quad_index(ilocr, er - 1, iq));
            else zbot_face(1, iq, f) = zbot_face(0, iq, f);
        }
    }
    // compute_gradient_quad (mod_Tensorproduct.F90:57-111)
    grad_zbot_quad.alloc(2, npoin_q);
    for (int e = 0; e < nelem; ++e)
        for (int jq = 0; jq < nq; ++jq)
            for (int iq = 0; iq < nq; ++iq) {
                int Iq = e * nq * nq + jq * nq + iq;
                double e_x = ksiq_x(Iq), e_y = ksiq_y(Iq), n_x = etaq_x(Iq), n_y = etaq_y(Iq);
                for (int m = 0; m < ngl; ++m)
                    for (int n = 0; n < ngl; ++n) {
                        double h_e = dpsiq(n, iq) * psiq(m, jq), h_n = psiq(n, iq) * dpsiq(m, jq);
                        double qv = zbot_df(e * npts + m * ngl + n);
                        grad_zbot_quad(0, Iq) += (h_e * e_x + h_n * n_x) * qv;
                        grad_zbot_quad(1, Iq) += (h_e * e_y + h_n * n_y) * qv;
                    }
            }
    // N_btp (mod_initial.F90:176-177)
    N_btp = (int)std::ceil(cfg.dt / cfg.dt_btp);
    dt = cfg.dt;
    dt_btp = cfg.dt / (double)N_btp;
    kstages = cfg.kstages;
    // wind_stress_coriolis (mod_initial_mlswe.F90:287-350)
    tau_wind.alloc(2, npoin_q); coriolis_df.alloc(npoin); coriolis_quad.alloc(npoin_q);
    fdt_bcl.alloc(npoin); fdt2_bcl.alloc(npoin); a_bcl.alloc(npoin); b_bcl.alloc(npoin);
    {
        double Lyy = cfg.ydims[1], ym = 0.5 * Lyy;
        for (int I = 0; I < npoin; ++I) coriolis_df(I) = cfg.f0 + cfg.beta * (coord(1, I) - ym);
        for (int e = 0; e < nelem; ++e)
            for (int jq = 0; jq < nq; ++jq)
                for (int iq = 0; iq < nq; ++iq) {
                    int Iq = e * nq * nq + jq * nq + iq;
                    for (int m = 0; m < ngl; ++m)
                        for (int n = 0; n < ngl; ++n) {
                            int I = e * npts + m * ngl + n;
                            double hi = psiq(n, iq) * psiq(m, jq);
                            coriolis_quad(Iq) += coriolis_df(I) * hi;
                            tau_wind(0, Iq) += tau_wind_df(0, I) * hi;
                            tau_wind(1, Iq) += tau_wind_df(1, I) * hi;
                        }
                }
        for (int I = 0; I < npoin; ++I) {
            fdt_bcl(I) = dt * coriolis_df(I);
            fdt2_bcl(I) = 0.5 * fdt_bcl(I);
            a_bcl(I) = 1.0 / (1.0 + fdt2_bcl(I) * fdt2_bcl(I));
            b_bcl(I) = fdt2_bcl(I) / (1.0 + fdt2_bcl(I) * fdt2_bcl(I));
        }
    }
    // ssprk_coefficients (mod_initial_mlswe.F90:652-678), non-'lsrk' branch
    ssprk_a.alloc(kstages, 3); ssprk_beta.alloc(kstages);
    auto seta = [&](int ik, double a1, double a2, double a3, double b) {
        ssprk_a(ik - 1, 0) = a1; ssprk_a(ik - 1, 1) = a2; ssprk_a(ik - 1, 2) = a3; ssprk_beta(ik - 1) = b;
    };
    switch (kstages) {
        case 1: seta(1, 1.0, 0.0, 0.0, 1.0); break;
        case 2: seta(1, 1.0, 0.0, 0.0, 1.0); seta(2, 0.5, 0.5, 0.0, 0.5); break;
        case 3: seta(1, 1.0, 0.0, 0.0, 1.0); seta(2, 3.0 / 4.0, 1.0 / 4.0, 0.0, 1.0 / 4.0);
                seta(3, 1.0 / 3.0, 2.0 / 3.0, 0.0, 2.0 / 3.0); break;
        case 4: seta(1, 1.0, 0.0, 0.0, 0.5); seta(2, 0.0, 1.0, 0.0, 0.5);
                seta(3, 2.0 / 3.0, 1.0 / 3.0, 0.0, 1.0 / 6.0); seta(4, 0.0, 1.0, 0.0, 0.5); break;
        case 5:
            seta(1, 1.0, 0.0, 0.0, 0.377268915331368);
            seta(2, 0.0, 1.0, 0.0, 0.377268915331368);
            seta(3, 0.355909775063326, 0.644090224936674, 0.0, 0.242995220537396);
            seta(4, 0.367933791638137, 0.632066208361863, 0.0, 0.238458932846290);
            seta(5, 0.0, 0.762406163401431, 0.237593836598569, 0.287632146308408);
            break;
        default: std::fprintf(stderr, "oracle: kstages must be 1..5\n"); std::abort();
    }
}

// mod_variables.F90:51-107
void Oracle::allocate_variables() {
    Q_uu_dp.alloc(npoin_q); Q_uv_dp.alloc(npoin_q); Q_vv_dp.alloc(npoin_q); H_bcl.alloc(npoin_q);
    Q_uu_dp_edge.alloc(nq, nface); Q_uv_dp_edge.alloc(nq, nface); Q_vv_dp_edge.alloc(nq, nface); H_bcl_edge.alloc(nq, nface);
    ope_ave.alloc(npoin_q); H_ave.alloc(npoin_q); Qu_ave.alloc(npoin_q); Qv_ave.alloc(npoin_q); Quv_ave.alloc(npoin_q);
    ope2_ave.alloc(npoin_q); btp_mass_flux_ave.alloc(2, npoin_q); uvb_ave.alloc(2, npoin_q); ope2_ave_df.alloc(npoin);
    uvb_face_ave.alloc(2, 2, nq, nface); btp_mass_flux_face_ave.alloc(2, nq, nface); ope_face_ave.alloc(2, nq, nface);
    H_face_ave.alloc(nq, nface); Qu_face_ave.alloc(2, nq, nface); Qv_face_ave.alloc(2, nq, nface);
    Quv_face_ave.alloc(2, nq, nface); tau_wind_ave.alloc(2, npoin_q); tau_bot_ave.alloc(2, npoin_q);
    one_plus_eta_edge_2_ave.alloc(nq, nface); uvb_ave_df.alloc(2, npoin); ope2_face_ave.alloc(2, nq, nface);
    dpprime_visc.alloc(npoin, nl); dpprime_visc_q.alloc(npoin_q, nl); pbprime_visc.alloc(npoin); btp_dpp_graduv.alloc(4, npoin);
    dpp_graduv.alloc(4, npoin, nl); graduv_dpp_face.alloc(5, 2, ngl, nface, nl); btp_graduv_dpp_face.alloc(5, 2, ngl, nface);
    graduvb_face_ave.alloc(4, 2, ngl, nface); graduvb_ave.alloc(4, npoin);
    sum_layer_mass_flux.alloc(2, npoin_q); sum_layer_mass_flux_face.alloc(2, nq, nface);
}

// Test aid (not in the reference): per-element constant geometry, see Config::affine_metrics.
void Oracle::make_metrics_affine() {
    const int nq2 = nq * nq;
    for (int e = 0; e < nelem; ++e) {
        int I0 = e * nq2;
        double kx = ksiq_x(I0), ky = ksiq_y(I0), ex = etaq_x(I0), ey = etaq_y(I0), J = jacq(I0) / (wnq[0] * wnq[0]);
        for (int j = 0; j < nq; ++j)
            for (int i = 0; i < nq; ++i) {
                int Iq = I0 + j * nq + i;
                ksiq_x(Iq) = kx; ksiq_y(Iq) = ky; etaq_x(Iq) = ex; etaq_y(Iq) = ey; jacq(Iq) = wnq[i] * wnq[j] * J;
            }
        for (int j = 0; j < ngl; ++j)
            for (int i = 0; i < ngl; ++i) {
                int I = e * npts + j * ngl + i;
                ksi_x(I) = kx; ksi_y(I) = ky; eta_x(I) = ex; eta_y(I) = ey; jac(I) = wgl[i] * wgl[j] * J;
                massinv(I) = 1.0 / jac(I);
            }
    }
    for (int f = 0; f < nface; ++f) {
        double nx = normal_vector_q(0, 0, f), ny = normal_vector_q(1, 0, f), nlen = jac_faceq(0, f) / wnq[0];
        for (int iq = 0; iq < nq; ++iq) { normal_vector_q(0, iq, f) = nx; normal_vector_q(1, iq, f) = ny; jac_faceq(iq, f) = wnq[iq] * nlen; }
        for (int n = 0; n < ngl; ++n) { normal_vector(0, n, f) = nx; normal_vector(1, n, f) = ny; jac_face(n, f) = wgl[n] * nlen; }
    }
}

Oracle::Oracle(const Config& c) : cfg(c) {
    build_basis();
    build_grid();
    build_metrics();
    build_faces();
    if (cfg.affine_metrics && cfg.mesh_warp == 0.0) make_metrics_affine();
    build_tensor_tables();
    build_initial();
    allocate_variables();
#define REG(a) reg[#a] = &a
    REG(psi); REG(dpsi); REG(psiq); REG(dpsiq); REG(coord);
    REG(ksi_x); REG(ksi_y); REG(eta_x); REG(eta_y); REG(jac); REG(ksiq_x); REG(ksiq_y); REG(etaq_x); REG(etaq_y); REG(jacq);
    REG(massinv); REG(normal_vector); REG(jac_face); REG(normal_vector_q); REG(jac_faceq);
    REG(wjac); REG(wjac_df);
    REG(alpha_mlswe); REG(pbprime); REG(pbprime_df); REG(pbprime_face); REG(one_over_pbprime); REG(one_over_pbprime_face);
    REG(pbprime_edge); REG(one_over_pbprime_edge); REG(one_over_pbprime_df); REG(one_over_pbprime_df_face); REG(pbprime_df_face);
    REG(tau_wind); REG(tau_wind_df); REG(coriolis_df); REG(coriolis_quad); REG(coeff_pbpert_L); REG(coeff_pbpert_R);
    REG(coeff_pbub_LR); REG(coeff_mass_pbub_L); REG(coeff_mass_pbub_R); REG(coeff_mass_pbpert_LR); REG(zbot); REG(zbot_df);
    REG(zbot_face); REG(grad_zbot_quad); REG(fdt_bcl); REG(fdt2_bcl); REG(a_bcl); REG(b_bcl); REG(ssprk_a); REG(ssprk_beta);
    REG(z_interface); REG(q_df); REG(qb_df); REG(qprime_df);
    REG(Q_uu_dp); REG(Q_uv_dp); REG(Q_vv_dp); REG(H_bcl); REG(Q_uu_dp_edge); REG(Q_uv_dp_edge); REG(Q_vv_dp_edge); REG(H_bcl_edge);
    REG(ope_ave); REG(H_ave); REG(Qu_ave); REG(Qv_ave); REG(Quv_ave); REG(ope2_ave); REG(btp_mass_flux_ave); REG(uvb_ave);
    REG(ope2_ave_df); REG(uvb_face_ave); REG(btp_mass_flux_face_ave); REG(ope_face_ave); REG(H_face_ave); REG(Qu_face_ave);
    REG(Qv_face_ave); REG(Quv_face_ave); REG(tau_wind_ave); REG(tau_bot_ave); REG(one_plus_eta_edge_2_ave); REG(uvb_ave_df);
    REG(ope2_face_ave); REG(dpprime_visc); REG(dpprime_visc_q); REG(pbprime_visc); REG(btp_dpp_graduv); REG(dpp_graduv); REG(graduv_dpp_face);
    REG(btp_graduv_dpp_face); REG(graduvb_face_ave); REG(graduvb_ave); REG(sum_layer_mass_flux); REG(sum_layer_mass_flux_face);
#undef REG
}

}  // namespace orc
