"""Brick decks: the arrays the Fortran driver of h-NUMO hands to the hot path.

This module stands in for the parts of the reference that stay Fortran (input deck, p4est brick mesh, basis,
metrics, initial conditions) when the library is driven from Python (tests, bench.py).  It produces exactly the
members of ``hnumo_desc_t`` (include/hnumo_b200.h) plus the initial state, for one rank of a row-block partition.

Reference formulas restated here (set-up only, not on the hot path):
  LGL nodes/weights, Lagrange tables     src/mod_legendre.F90:54-111,248-322,387-433, src/mod_basis.F90:60-186
  brick coordinates, face table          src/p4est.c:157-242,1560-1735, src/mod_p4est.F90:344-359
  initial conditions, statics            src/initial_conditions.F90:86-416, src/mod_initial_mlswe.F90:287-350,652-678
"""
import math

import numpy as np

GRAVITY = 9.806  # initial_conditions.F90:97,132,173
PI_TRIG = 3.1415926535897932346


# ---------------------------------------------------------------------------------------------------------------
def _legendre_poly(n, x):
    p2 = p2_1 = p2_2 = 0.0
    p1 = p1_1 = p1_2 = 0.0
    p0, p0_1, p0_2 = 1.0, 0.0, 0.0
    for j in range(1, n + 1):
        p2, p2_1, p2_2 = p1, p1_1, p1_2
        p1, p1_1, p1_2 = p0, p0_1, p0_2
        a = (2.0 * j - 1.0) / j
        b = (j - 1.0) / j
        p0 = a * x * p1 - b * p2
        p0_1 = a * (p1 + x * p1_1) - b * p2_1
        p0_2 = a * (2.0 * p1_1 + x * p1_2) - b * p2_2
    return p0, p0_1, p0_2


def lgl(ngl):
    """Legendre-Gauss-Lobatto points and weights (mod_legendre.F90:54-111)."""
    x = np.zeros(ngl)
    w = np.zeros(ngl)
    if ngl == 1:
        return np.array([0.0]), np.array([2.0])
    n = ngl - 1
    nh = (n + 1) // 2
    eps = np.finfo(float).eps
    pi = 4.0 * math.atan(1.0)
    for i in range(1, nh + 1):
        xx = math.cos((2.0 * i - 1.0) / (2.0 * n + 1.0) * pi)
        p0 = 0.0
        for _ in range(20):
            p0, p0_1, p0_2 = _legendre_poly(n, xx)
            dx = -(1.0 - xx * xx) * p0_1 / (-2.0 * xx * p0_1 + (1.0 - xx * xx) * p0_2)
            xx = xx + dx
            if abs(dx) < eps:
                break
        x[n + 1 - i] = xx
        w[n + 1 - i] = 2.0 / (n * (n + 1) * p0 * p0)
    if n + 1 != 2 * nh:
        p0, _, _ = _legendre_poly(n, 0.0)
        x[nh] = 0.0
        w[nh] = 2.0 / (n * (n + 1) * p0 * p0)
    for i in range(1, nh + 1):
        x[i - 1] = -x[n + 1 - i]
        w[i - 1] = w[n + 1 - i]
    return x, w


def basis(nop, dg_integ_exact=True):
    """1-D operator tables of mod_basis: psiq(ngl,nq), dpsiq(ngl,nq), dpsi(ngl,ngl), weights."""
    ngl = nop + 1
    nq = 2 * nop + 1 if dg_integ_exact else 2 * nop - 1
    xgl, wgl = lgl(ngl)
    xnq, wnq = lgl(nq)
    # legendre_basis, reduce_round_off branch
    bb = np.zeros(ngl)
    for j in range(ngl):
        for i in range(ngl):
            if i != j:
                bb[j] += math.log(abs(xgl[j] - xgl[i]))
    dpsi = np.zeros((ngl, ngl))  # dpsi[i, j] = l_i'(x_j)
    cc = np.zeros(ngl)
    for j in range(ngl):
        for i in range(ngl):
            if i != j:
                sgn = 1.0 if (i + j) % 2 == 0 else -1.0
                dpsi[i, j] = sgn * math.exp(bb[j] - bb[i]) / (xgl[j] - xgl[i])
                cc[j] += dpsi[i, j]
    for j in range(ngl):
        dpsi[j, j] = -cc[j]
    # lagrange_basis
    psiq = np.zeros((ngl, nq))
    dpsiq = np.zeros((ngl, nq))
    for l in range(nq):
        xl = xnq[l]
        for i in range(ngl):
            ksi = xgl[i]
            p = 1.0
            dp = 0.0
            for j in range(ngl):
                if j != i:
                    p = p * (xl - xgl[j]) / (ksi - xgl[j])
                    dd = 1.0
                    for k in range(ngl):
                        if k != i and k != j:
                            dd = dd * (xl - xgl[k]) / (ksi - xgl[k])
                    dp = dp + dd / (ksi - xgl[j])
            psiq[i, l] = p
            dpsiq[i, l] = dp
    return dict(ngl=ngl, nq=nq, xgl=xgl, wgl=wgl, xnq=xnq, wnq=wnq, psiq=psiq, dpsiq=dpsiq, dpsi=dpsi)


# ---------------------------------------------------------------------------------------------------------------
SHIPPED = {
    # Examples/bump/numo3d.in == CI/bump/numo3d.in
    "bump": dict(nelx=10, nely=10, nop=4, nlayers=2, xdims=(0.0, 2e3), ydims=(0.0, 2e3), x_boundary=(4, 4),
                 y_boundary=(4, 4), dt=100.0, dt_btp=1.8, kstages=5, botfr=0, cd_mlswe=0.0, method_visc=0,
                 visc_mlswe=0.0, f0=0.0, beta=0.0, test_case="bump"),
    # Examples/lake/numo3d.in
    "lake": dict(nelx=10, nely=10, nop=4, nlayers=2, xdims=(0.0, 2e3), ydims=(0.0, 2e3), x_boundary=(4, 4),
                 y_boundary=(4, 4), dt=100.0, dt_btp=1.8, kstages=5, botfr=0, cd_mlswe=0.0, method_visc=0,
                 visc_mlswe=0.0, f0=0.0, beta=0.0, test_case="lakeAtrest"),
    # Examples/double_gyre/numo3d.in
    "double_gyre": dict(nelx=25, nely=25, nop=4, nlayers=2, xdims=(0.0, 2e6), ydims=(0.0, 2e6), x_boundary=(4, 4),
                        y_boundary=(4, 4), dt=500.0, dt_btp=25.0, kstages=5, botfr=1, cd_mlswe=1.0e-7, method_visc=3,
                        visc_mlswe=50.0, f0=0.93e-4, beta=2.0e-11, test_case="double-gyre"),
}


def synthetic_double_gyre(nelx, nely, nop=4, nlayers=3, dt=None, dt_btp=None, perturb=1.0e-3):
    """Synthetic double gyre of SURVEY.md 8(d) / BASELINE.md configs 4 and 5 (CFL-rescaled time steps)."""
    p = dict(SHIPPED["double_gyre"])
    p.update(nelx=nelx, nely=nely, nop=nop, nlayers=nlayers, test_case="double-gyre-synth", synth_perturb=perturb)
    H = 9928.0
    if nlayers == 2:
        z = [0.0, -1489.5, -H]
        alpha = [9.7370e-4, 9.7350e-4]
    elif nlayers == 3:
        z = [0.0, -1489.5, -4000.0, -H]
        alpha = [9.7370e-4, 9.7360e-4, 9.7350e-4]
    else:
        z = [0.0] + [-round(2.0 * H * (k / nlayers) ** 1.5) / 2.0 for k in range(1, nlayers)] + [-H]
        alpha = [9.7370e-4 + (9.7350e-4 - 9.7370e-4) * k / (nlayers - 1) for k in range(nlayers)]
    p["synth_z"] = z
    p["synth_alpha"] = alpha
    # keep the shipped barotropic CFL (~0.56): dt_btp scales with the minimum LGL spacing of the element
    xg, _ = lgl(nop + 1)
    dmin = (xg[1] - xg[0]) / 2.0 * (2.0e6 / max(nelx, nely))
    xg4, _ = lgl(5)
    dmin_ref = (xg4[1] - xg4[0]) / 2.0 * (2.0e6 / 25.0)
    if dt_btp is None:
        dt_btp = 25.0 * dmin / dmin_ref
    if dt is None:
        dt = 20.0 * dt_btp
    p["dt"], p["dt_btp"] = dt, dt_btp
    return p


# ---------------------------------------------------------------------------------------------------------------
def element_owner(p, nranks):
    """Owning rank of every element of the nelx x nely brick (global row-major numbering).

    p["partition"]: "rows" (default: nranks row blocks), "blocks:PXxPY" (2-D blocks: ranks with up to 4 neighbours),
    "morton" (contiguous chunks of the Morton / z-order curve, the shape of a uniform-weight p4est partition,
    p4est.c:1178)."""
    nelx, nely = p["nelx"], p["nely"]
    kind = p.get("partition", "rows")
    g = np.arange(nelx * nely)
    ex, ey = g % nelx, g // nelx
    if nranks == 1:
        return np.zeros(nelx * nely, dtype=np.int64)
    if kind == "rows":
        if nely % nranks != 0:
            raise ValueError("row-block partition needs nely divisible by the number of ranks")
        return ey // (nely // nranks)
    if kind.startswith("blocks:"):
        px, py = (int(t) for t in kind.split(":")[1].split("x"))
        if px * py != nranks or nelx % px or nely % py:
            raise ValueError("blocks partition: PX*PY must equal the number of ranks and divide the brick")
        return (ey // (nely // py)) * px + ex // (nelx // px)
    if kind == "morton":
        code = np.zeros(nelx * nely, dtype=np.int64)
        for b in range(16):
            code |= ((ex >> b) & 1) << (2 * b)
            code |= ((ey >> b) & 1) << (2 * b + 1)
        order = np.argsort(code, kind="stable")
        owner = np.empty(nelx * nely, dtype=np.int64)
        owner[order] = (np.arange(nelx * nely) * nranks) // (nelx * nely)
        return owner
    raise ValueError("unknown partition " + kind)


def build_deck(params, rank=0, nranks=1):
    """Arrays for one rank of a partition of the nelx x nely brick (element_owner: row blocks, 2-D blocks, Morton chunks).

    Returns a dict whose keys match hnumo_desc_t members, plus 'q_df', 'qb_df', 'qprime_df' (initial state in the
    reference AoS layout), 'coord', 'elem_global' (global row-major element number of each local element).
    """
    p = params
    nelx, nely, nop, nl = p["nelx"], p["nely"], p["nop"], p["nlayers"]
    B = basis(nop, p.get("dg_integ_exact", True))
    ngl, nq = B["ngl"], B["nq"]
    npts = ngl * ngl
    xg = B["xgl"]
    Lx = p["xdims"][1] - p["xdims"][0]
    Ly = p["ydims"][1] - p["ydims"][0]
    owner = element_owner(p, nranks)                 # owning rank of every global (row-major) element
    gl = np.nonzero(owner == rank)[0]                # local elements keep the global order: left element = lower number
    nelem = gl.size
    if nelem == 0:
        raise ValueError("partition leaves rank %d without elements" % rank)
    npoin = nelem * npts
    ex = gl % nelx
    ey = gl // nelx
    g2l = np.full(nelx * nely, -1, dtype=np.int64)
    g2l[gl] = np.arange(nelem)
    # coordinates (p4est.c:204-236): bilinear blend of the tree corners, then rescaled to xdims/ydims
    rl = xg[None, :]
    rm = xg[:, None]
    w0 = (1 - rl) * (1 - rm); w1 = (1 + rl) * (1 - rm); w2 = (1 - rl) * (1 + rm); w3 = (1 + rl) * (1 + rm)
    exb = ex[:, None, None].astype(float)
    eyb = ey[:, None, None].astype(float)
    tx = (w0 * exb + w1 * (exb + 1) + w2 * exb + w3 * (exb + 1)) / 4.0
    ty = (w0 * eyb + w1 * eyb + w2 * (eyb + 1) + w3 * (eyb + 1)) / 4.0
    X = (tx / nelx) * Lx + p["xdims"][0]
    Y = (ty / nely) * Ly + p["ydims"][0]
    warp = float(p.get("mesh_warp", 0.0))
    if warp != 0.0:
        # general (curved, non-affine) quadrilaterals: the nodes of the brick are moved by a smooth displacement that vanishes
        # in the wall-normal direction on the domain boundary (test aid; the reference reads such meshes from gmsh files)
        kx, ky = float(max(1, nelx // 2)), float(max(1, nely // 2))
        Xn, Yn = (X - p["xdims"][0]) / Lx, (Y - p["ydims"][0]) / Ly
        ddx = warp * (Lx / nelx) * np.sin(PI_TRIG * kx * Xn) * np.cos(PI_TRIG * ky * Yn)
        ddy = warp * (Ly / nely) * np.cos(PI_TRIG * kx * Xn) * np.sin(PI_TRIG * ky * Yn)
        X = X + ddx
        Y = Y + ddy
    x = X.reshape(-1)
    y = Y.reshape(-1)
    # face table (p4est.c:1590-1704): p4est faces f=0..3 (-x,+x,-y,+y) -> numa local faces 5,6,3,4
    transform = np.array([5, 6, 3, 4], dtype=np.int32)
    qq = np.arange(nelem)
    cols = []
    for f in range(4):
        nx_ = ex + (-1 if f == 0 else 1 if f == 1 else 0)
        ny_ = ey + (-1 if f == 2 else 1 if f == 3 else 0)
        wall = (nx_ < 0) | (nx_ >= nelx) | (ny_ < 0) | (ny_ >= nely)
        ng = np.where(wall, 0, nx_ + nelx * ny_)     # global number of the neighbour
        nrank = np.where(wall, rank, owner[ng])
        proc = ~wall & (nrank != rank)
        inter = ~wall & ~proc
        nq_ = np.where(inter, g2l[ng], -1)
        bc = (p["x_boundary"][0] if f == 0 else p["x_boundary"][1] if f == 1 else
              p["y_boundary"][0] if f == 2 else p["y_boundary"][1])
        er = np.where(wall, -bc, np.where(proc, 0, nq_ + 1)).astype(np.int32)
        ilocr = np.where(inter, transform[f ^ 1], 0).astype(np.int32)
        keep = wall | proc | (inter & (qq < nq_))
        cols.append((np.full(nelem, transform[f], dtype=np.int32), ilocr, (qq + 1).astype(np.int32), er, keep,
                     np.where(proc, nrank, -1), np.minimum(gl, ng), np.maximum(gl, ng)))
    ilocl = np.stack([c[0] for c in cols], axis=1).ravel()
    ilocr = np.stack([c[1] for c in cols], axis=1).ravel()
    el = np.stack([c[2] for c in cols], axis=1).ravel()
    er = np.stack([c[3] for c in cols], axis=1).ravel()
    keep = np.stack([c[4] for c in cols], axis=1).ravel()
    prank = np.stack([c[5] for c in cols], axis=1).ravel()[keep]
    glo = np.stack([c[6] for c in cols], axis=1).ravel()[keep]
    ghi = np.stack([c[7] for c in cols], axis=1).ravel()[keep]
    nface = int(keep.sum())
    face = np.zeros((nface, 8), dtype=np.int32)
    face[:, 4] = ilocl[keep]; face[:, 5] = ilocr[keep]; face[:, 6] = el[keep]; face[:, 7] = er[keep]
    # neighbour lists (p4est.c:1340-1420, mod_parallel.F90:141-171): per neighbour rank, the shared faces in an order both
    # sides agree on -- here by the pair of global element numbers of the face
    nbh_proc, num_send_recv, nbh_send_recv = [], [], []
    for r in sorted(set(prank[prank >= 0].tolist())):
        fi = np.nonzero(prank == r)[0]
        fi = fi[np.lexsort((ghi[fi], glo[fi]))]
        nbh_proc.append(r + 1); num_send_recv.append(len(fi)); nbh_send_recv += (fi + 1).tolist()   # 1-based
    # geometry: axis-aligned rectangles
    dx, dy = Lx / nelx, Ly / nely
    em = np.zeros((nelem, 5))
    em[:, 0] = 2.0 / dx; em[:, 3] = 2.0 / dy; em[:, 4] = dx * dy / 4.0
    ntab = np.zeros((7, 3))
    ntab[3] = (0.0, -1.0, dx / 2.0); ntab[4] = (0.0, 1.0, dx / 2.0); ntab[5] = (-1.0, 0.0, dy / 2.0); ntab[6] = (1.0, 0.0, dy / 2.0)
    fg = ntab[face[:, 4]]
    wg = B["wgl"]
    massinv = np.tile(1.0 / (wg[None, :] * wg[:, None] * (dx * dy / 4.0)), (nelem, 1, 1)).reshape(-1)
    general = {}
    if warp != 0.0:
        general = _point_geometry(B, X, Y, face)
        massinv = 1.0 / general["point_metrics"][:, 4]       # DG: mass = jac (create_mass.F90:24-31)
        em = None; fg = None
    # ---- initial conditions (initial_conditions.F90)
    g = GRAVITY
    pi = PI_TRIG
    xmin, xmax = p["xdims"]
    ymin, ymax = p["ydims"]
    zbot = np.zeros(npoin)
    zi = np.zeros((nl + 1, npoin))
    alpha = np.zeros(nl)
    tauw = np.zeros((npoin, 2))
    tc = p["test_case"]
    if tc == "bump":
        H = 40.0
        zbot[:] = -H
        for k in range(nl + 1):
            zi[k] = -(k) * H / float(nl)
        xm, yl = 0.5 * (xmax + xmin), 0.5 * (ymax + ymin)
        r = np.sqrt((x - xm) ** 2 + (y - yl) ** 2)
        m = r < 250.0
        zi[1][m] = zi[1][m] + 0.5 * 1.0 * (1.0 + np.cos(pi * r[m] / 250.0))
        alpha[0] = 0.9737e-3
        alpha[1:] = 0.9735e-3
    elif tc == "lakeAtrest":
        H = 40.0
        zbot[:] = -H
        xm, yl = 0.5 * (xmin + xmax), 0.5 * (ymin + ymax)
        r = np.sqrt((x - xm) ** 2 + (y - yl) ** 2)
        m = r < 250.0
        zbot[m] = zbot[m] + 3.0 * (1.0 + np.cos(pi * r[m] / 250.0))
        for k in range(nl + 1):
            if nl < 5:
                zi[k] = -(k) * H / float(nl)
            else:
                zi[k] = -(k) * 32 / float(nl - 1)
        if nl >= 5:
            zi[nl] = -H
        rho0 = 1027.01037
        alpha[0] = 1.0 / rho0
        for k in range(2, nl + 1):
            alpha[k - 1] = 1.0 / (rho0 + k * 0.2110 / float(nl))
    elif tc == "double-gyre":
        H = 9928.0
        zbot[:] = -H
        zi[1] = -1489.5
        zi[2] = -H
        alpha[0], alpha[1] = 9.7370e-04, 9.7350e-04
        tauw[:, 0] = -0.1 * np.cos(2.0 * pi * y / Ly)
    elif tc == "double-gyre-synth":
        z = p["synth_z"]
        zbot[:] = z[nl]
        for k in range(nl + 1):
            zi[k] = z[k]
        alpha[:] = p["synth_alpha"]
        tauw[:, 0] = -0.1 * np.cos(2.0 * pi * y / Ly)
    else:
        raise ValueError("unknown test case " + tc)
    zi = np.maximum(zi, zbot[None, :])
    pbprime = np.zeros(npoin)
    for k in range(nl):
        pbprime = pbprime + (g / alpha[k]) * (zi[k] - zi[k + 1])
    q = np.zeros((nl, npoin, 3))
    qprime = np.zeros((nl, npoin, 3))
    ope = np.zeros(npoin)
    for k in range(nl):
        q[k, :, 0] = (g / alpha[k]) * (zi[k] - zi[k + 1])
        ope = ope + q[k, :, 0] / pbprime
    for k in range(nl):
        qprime[k, :, 0] = q[k, :, 0] / ope
    if tc == "double-gyre-synth" and p.get("synth_perturb", 0.0) != 0.0:
        s = p["synth_perturb"] * np.sin(2.0 * pi * x / Lx) * np.sin(2.0 * pi * y / Ly)
        d = q[0, :, 0] * s
        q[0, :, 0] += d
        q[1, :, 0] -= d
        qprime[0, :, 0] = q[0, :, 0] / ope
        qprime[1, :, 0] = q[1, :, 0] / ope
    qb = np.zeros((npoin, 4))
    for k in range(nl):
        qb[:, 0] += q[k, :, 0]
        qb[:, 2] += q[k, :, 1]
        qb[:, 3] += q[k, :, 2]
    qb[:, 1] = qb[:, 0] - pbprime
    for k in range(nl):
        qprime[k, :, 1] = q[k, :, 1] / q[k, :, 0] - qb[:, 2] / qb[:, 0]
        qprime[k, :, 2] = q[k, :, 2] / q[k, :, 0] - qb[:, 3] / qb[:, 0]
    coriolis = p.get("f0", 0.0) + p.get("beta", 0.0) * (y - 0.5 * p["ydims"][1])  # mod_initial_mlswe.F90:308-315
    N_btp = int(math.ceil(p["dt"] / p["dt_btp"]))
    dt_btp = p["dt"] / float(N_btp)
    ks = p.get("kstages", 5)
    tab = {
        1: ([[1.0, 0.0, 0.0]], [1.0]),
        2: ([[1.0, 0.0, 0.0], [0.5, 0.5, 0.0]], [1.0, 0.5]),
        3: ([[1.0, 0.0, 0.0], [3.0 / 4.0, 1.0 / 4.0, 0.0], [1.0 / 3.0, 2.0 / 3.0, 0.0]], [1.0, 1.0 / 4.0, 2.0 / 3.0]),
        4: ([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [2.0 / 3.0, 1.0 / 3.0, 0.0], [0.0, 1.0, 0.0]], [0.5, 0.5, 1.0 / 6.0, 0.5]),
        5: ([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.355909775063326, 0.644090224936674, 0.0],
             [0.367933791638137, 0.632066208361863, 0.0], [0.0, 0.762406163401431, 0.237593836598569]],
            [0.377268915331368, 0.377268915331368, 0.242995220537396, 0.238458932846290, 0.287632146308408]),
    }[ks]
    ssprk_a = np.asfortranarray(np.array(tab[0]))  # (kstages,3) column-major
    deck = dict(
        nelem=nelem, ngl=ngl, nq=nq, nlayers=nl, nface=nface, kstages=ks, N_btp=N_btp, dt=float(p["dt"]), dt_btp=dt_btp,
        botfr=p.get("botfr", 0), method_visc=p.get("method_visc", 0), gravity=g, cd_mlswe=p.get("cd_mlswe", 0.0),
        visc_mlswe=p.get("visc_mlswe", 0.0), ad_mlswe=p.get("ad_mlswe", 0.0), max_shear_dz=p.get("max_shear_dz", 0.0),
        psiq=np.asfortranarray(B["psiq"]), dpsiq=np.asfortranarray(B["dpsiq"]), wnq=B["wnq"], wgl=B["wgl"],
        dpsi=np.asfortranarray(B["dpsi"]),
        face=face, elem_metrics=em, face_geom=fg,
        pbprime_df=pbprime, massinv=massinv, coriolis_df=coriolis, tau_wind_df=tauw, zbot_df=zbot, alpha_mlswe=alpha,
        ssprk_a=ssprk_a, ssprk_beta=np.array(tab[1]),
        rank=rank, nranks=nranks, nbh_proc=np.array(nbh_proc, dtype=np.int32),
        num_send_recv=np.array(num_send_recv, dtype=np.int32), nbh_send_recv=np.array(nbh_send_recv, dtype=np.int32),
        q_df=q, qb_df=qb, qprime_df=qprime, coord=np.stack([x, y], axis=1),
        elem_global=gl.copy(), npoin=npoin, npts=npts,
    )
    deck.update(general)
    return deck


def _point_geometry(B, X, Y, face):
    """Per-point geometry of general quadrilaterals from the node coordinates X, Y [nelem, m, n], as the reference derives it:
    nodal metrics (metrics.F90:40-127 with mod_gradient.F90:110-166), quadrature-point metrics (metrics_quad.F90:20-127 with
    compute_local_gradient_quad_v3, mod_gradient.F90:175-260), face normals and Jacobians at the face nodes and face quadrature
    points of the left element (create_normals.F90:17-215, create_normals_quad.F90:136-211).  Returns the arrays of the
    hnumo_desc_t members point_metrics_q (npoin_q,5), point_metrics (npoin,5), face_geom_q (nface,nq,3), face_geom_n (nface,ngl,3)."""
    ngl, nq = B["ngl"], B["nq"]
    dpsi, psiq, dpsiq, wgl, wnq = B["dpsi"], B["psiq"], B["dpsiq"], B["wgl"], B["wnq"]

    def metrics(xk, yk, xe, ye, w):
        xj = xk * ye - yk * xe
        return np.stack([ye / xj, -xe / xj, -yk / xj, xk / xj, (w[None, None, :] * w[None, :, None]) * np.abs(xj)], axis=-1)

    # nodal: x_ksi(i,j) = sum_n dpsi(n,i) x(n,j), x_eta(i,j) = sum_n x(i,n) dpsi(n,j); arrays [e, j, i]
    nk = [np.einsum("ni,ejn->eji", dpsi, A) for A in (X, Y)]
    ne = [np.einsum("eni,nj->eji", A, dpsi) for A in (X, Y)]
    # quadrature points: x_ksi(iq,jq) = sum_mn dpsiq(n,iq) psiq(m,jq) x(n,m)
    qk = [np.einsum("ni,mj,emn->eji", dpsiq, psiq, A) for A in (X, Y)]
    qe = [np.einsum("ni,mj,emn->eji", psiq, dpsiq, A) for A in (X, Y)]
    nelem = X.shape[0]
    out = dict(point_metrics=metrics(nk[0], nk[1], ne[0], ne[1], wgl).reshape(nelem * ngl * ngl, 5),
               point_metrics_q=metrics(qk[0], qk[1], qe[0], qe[1], wnq).reshape(nelem * nq * nq, 5))
    iel, iloc = face[:, 6] - 1, face[:, 4]

    def face_geom(n1, dk, de, w):
        l = np.arange(n1)[None, :]
        il = iloc[:, None]
        i = np.where(il == 3, l, np.where(il == 4, l, np.where(il == 5, 0, n1 - 1)))
        j = np.where(il == 3, 0, np.where(il == 4, n1 - 1, l))
        e = iel[:, None]
        xk, yk, xe, ye = dk[0][e, j, i], dk[1][e, j, i], de[0][e, j, i], de[1][e, j, i]
        nx = np.where(il == 3, yk, np.where(il == 4, -yk, np.where(il == 5, -ye, ye)))
        ny = np.where(il == 3, -xk, np.where(il == 4, xk, np.where(il == 5, xe, -xe)))
        nlen = np.sqrt(nx * nx + ny * ny)
        return np.stack([nx / nlen, ny / nlen, w[None, :] * nlen], axis=-1)

    out["face_geom_n"] = face_geom(ngl, nk, ne, wgl)
    out["face_geom_q"] = face_geom(nq, qk, qe, wnq)
    return out


def diagnostics(deck, q_df):
    """h, u, v, interface elevation per layer and layer masses (diagnostics.F90:24-45, compute_conserved.F90)."""
    nl, npoin = deck["nlayers"], deck["npoin"]
    g, alpha = deck["gravity"], deck["alpha_mlswe"]
    q = np.asarray(q_df).reshape(nl, npoin, 3)
    h = np.stack([(alpha[k] / g) * q[k, :, 0] for k in range(nl)])
    u = q[:, :, 1] / q[:, :, 0]
    v = q[:, :, 2] / q[:, :, 0]
    elev = np.zeros((nl + 1, npoin))
    elev[nl] = deck["zbot_df"]
    for k in range(nl - 1, -1, -1):
        elev[k] = elev[k + 1] + h[k]
    wj = 1.0 / deck["massinv"]
    mass = (h * wj[None, :]).sum(axis=1)
    return dict(h=h, u=u, v=v, ssh=elev[:nl], mass=mass)
