"""CPU tests that pin the oracle: the reference's own golden, documented invariants, conditioning of the path."""
import numpy as np
import pytest

import oracle_lib
from golden_util import ci_check, compare_extrema, read_fin
from hnumo_loader import hnumo_b200 as hn


def _diag(o):
    q, m = o.diagnostics()  # (nl, npoin, 5)
    return dict(h=q[:, :, 0], u=q[:, :, 1], v=q[:, :, 2], ssh=q[:, :, 4], mass=m)


@pytest.fixture(scope="module")
def bump108():
    o = oracle_lib.Oracle(hn.decks.SHIPPED["bump"])
    d0 = _diag(o)
    assert o.step(108) == 0
    return o, d0, _diag(o)


def test_oracle_reproduces_reference_golden(bump108):
    """CI/bump/ref_mlswe_FIN.txt: max/min of h,u,v,ssh per layer after 108 steps (60480 barotropic stages).
    The file holds 12 significant digits; agreement is limited by summation-order round-off amplified through the
    pressure-gradient cancellation (see test_conditioning), observed ~4e-9 of the velocity scale."""
    o, d0, d = bump108
    ref = read_fin()
    got = {k + 1: {f: (float(d[f][k].max()), float(d[f][k].min())) for f in ("h", "u", "v", "ssh")} for k in range(2)}
    assert compare_extrema(got, ref) < 2e-8
    # thickness extrema agree to 10 significant digits
    for layer in (1, 2):
        for a, b in zip(got[layer]["h"], ref[layer]["h"]):
            assert abs(a - b) / abs(b) < 1e-10


def test_oracle_mass_conservation(bump108):
    """check.F90:58-62: relative mass loss per layer must stay below 1e-12."""
    o, d0, d = bump108
    for k in range(2):
        assert abs(d["mass"][k] - d0["mass"][k]) / d0["mass"][k] < 1e-12


def test_oracle_bump_symmetry(bump108):
    """the bump is centred: max u = -min u = max v = -min v up to round-off (golden shows the same)."""
    o, d0, d = bump108
    for k in range(2):
        vals = [d["u"][k].max(), -d["u"][k].min(), d["v"][k].max(), -d["v"][k].min()]
        assert (max(vals) - min(vals)) / max(vals) < 1e-6


def test_oracle_lake_at_rest():
    """docs/source/test.rst:14-43: a lake at rest over a bump in the bathymetry stays at rest."""
    o = oracle_lib.Oracle(hn.decks.SHIPPED["lake"])
    d0 = _diag(o)
    assert o.step(5) == 0
    d = _diag(o)
    # "machine precision" here means round-off of the O(4e5 Pa) pressure terms: |u| ~ 1e-11 m/s per step
    assert np.abs(d["u"]).max() < 1e-9 and np.abs(d["v"]).max() < 1e-9
    assert np.abs(d["ssh"][0]).max() < 1e-10
    assert np.abs(d["h"] - d0["h"]).max() < 1e-7


def test_oracle_lgl_and_interpolation():
    o = oracle_lib.Oracle(dict(hn.decks.SHIPPED["bump"], nelx=2, nely=2))
    xgl, wgl, xnq, wnq = o.get("xgl"), o.get("wgl"), o.get("xnq"), o.get("wnq")
    assert abs(wgl.sum() - 2.0) < 1e-14 and abs(wnq.sum() - 2.0) < 1e-14
    # closed form for 5 LGL points: 0, +-sqrt(3/7), +-1 ; weights 32/45, 49/90, 1/10
    assert np.allclose(xgl, [-1, -np.sqrt(3 / 7), 0, np.sqrt(3 / 7), 1], atol=1e-15)
    assert np.allclose(wgl, [0.1, 49 / 90, 32 / 45, 49 / 90, 0.1], atol=1e-15)
    psiq = o.get("psiq").reshape(o.nq, o.ngl).T  # (ngl, nq)
    dpsiq = o.get("dpsiq").reshape(o.nq, o.ngl).T
    for deg in range(o.ngl):
        assert np.allclose(psiq.T @ xgl ** deg, xnq ** deg, atol=1e-13)      # exact interpolation up to degree N
        d = deg * xnq ** (deg - 1) if deg > 0 else 0 * xnq
        assert np.allclose(dpsiq.T @ xgl ** deg, d, atol=1e-12)
    assert abs(o.get("wjac").sum() - 2e3 * 2e3) / 4e6 < 1e-13          # sum of quadrature weights = area


def test_conditioning_noise_floor():
    """Why momentum-like fields cannot agree to 1e-11 in the plain relative sense between two correct FP64
    implementations: the momentum tendency is the small difference of O(H_bcl) volume and face pressure terms.
    Perturbing the geometry inputs at the 1e-14 level (numerically differentiated metrics as in metrics.F90 vs their
    per-element constants) changes the barotropic momentum after one step by ~1e-7 relative, while mass-like fields
    move by < 1e-11.  GPU parity tolerances in tests/test_gpu_parity.py are set from this floor."""
    p = hn.decks.SHIPPED["bump"]
    a = oracle_lib.Oracle(dict(p, affine_metrics=False))
    b = oracle_lib.Oracle(dict(p, affine_metrics=True))
    assert np.abs(a.get("massinv") / b.get("massinv") - 1).max() < 1e-12   # inputs differ by round-off only
    a.step(2); b.step(2)
    qa, qb = a.get("qb_df").reshape(-1, 4), b.get("qb_df").reshape(-1, 4)
    rel = lambda x, y: np.linalg.norm(x - y) / np.linalg.norm(y)
    assert rel(qa[:, 0], qb[:, 0]) < 1e-12
    mom = rel(qa[:, 2], qb[:, 2])
    assert 1e-10 < mom < 1e-5, mom
    c = np.sqrt(9.806 * 40.0)
    assert np.linalg.norm(qa[:, 2] - qb[:, 2]) / (c * np.linalg.norm(qb[:, 0])) < 1e-11


def test_oracle_rhs_invariants():
    """SURVEY 8(c) self-checks of the restated operators that need no Fortran run:
    (1) lake at rest: the barotropic RHS of the resting state vanishes to round-off of its O(g p_b^2) pressure terms
        (well-balancedness of volume + face terms, mod_rhs_btp.F90:102-370);
    (2) discrete mass conservation: sum_I wjac(I) rhs_pbpert(I) = boundary flux = 0 with walls on all sides, on a
        developed double-gyre state (check.F90:58 is the time-integrated form of this)."""
    o = oracle_lib.Oracle(hn.decks.SHIPPED["lake"])
    o.btp_bcl_coeffs()
    r = o.rhs_btp()
    pb = o.get("qb_df").reshape(-1, 4)[:, 0]
    g = 9.806
    # momentum tendencies are differences of terms of size H_bcl * massinv-weighted derivative ~ g*pb^2/alpha-scale
    Hn = np.abs(o.get("H_bcl")).max()
    h_min = np.sqrt((1.0 / o.get("massinv")).min())
    assert np.abs(r[:, 0]).max() <= 1e-12 * np.abs(pb).max()
    assert np.abs(r[:, 1:]).max() <= 1e-12 * Hn / h_min
    o2 = oracle_lib.Oracle(dict(hn.decks.SHIPPED["double_gyre"], nelx=8, nely=8))
    assert o2.step(2) == 0
    o2.btp_bcl_coeffs()
    r2 = o2.rhs_btp()
    wj = 1.0 / o2.get("massinv")
    total = float((wj * r2[:, 0]).sum())
    scale = float((wj * np.abs(r2[:, 0])).sum())
    assert scale > 0.0 and abs(total) <= 1e-12 * scale


def test_fin_file_passes_the_reference_ci_check(bump108, tmp_path):
    """The reference's CI (CI/bump/run_check.sh: numo3d, then check.F90 on mlswe_FIN.txt) applied to an mlswe_FIN.txt written by
    the package from the oracle's diagnostics of the same run: mass loss <= 1e-12 per layer, and the reported relative errors of
    the extrema of u, v, ssh against CI/bump/ref_mlswe_FIN.txt.  The file format is checked against the golden itself."""
    o, d0, d = bump108
    diag = dict(mass=d["mass"])
    for f in ("h", "u", "v", "ssh"):
        diag[f] = np.stack([d[f].max(axis=1), d[f].min(axis=1)], axis=1)
    path = tmp_path / "mlswe_FIN.txt"
    hn.write_fin(path, diag, d0["mass"])
    got = open(path).read().split("\n")
    ref = open(hn.__file__.replace("h-numo_b200/__init__.py", "tests/golden/ci_bump_ref_mlswe_FIN.txt")).read().split("\n")
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        # same columns: identical text up to the digits of the numbers
        assert len(a.rstrip()) == len(b.rstrip()) and a.split("=")[0] == b.split("=")[0], (a, b)
        if a.startswith("Fields"):
            assert a.split()[3] == b.split()[3] and [len(t) for t in a.split()] == [len(t) for t in b.split()], (a, b)
    errs = ci_check(path)
    # u, v: 1e-7 of the velocity scale is the conditioning noise floor of this run (test_conditioning_noise_floor); ssh of layer 1
    # is a 1e-5 m signal on a 20 m thickness
    for (layer, field), (emax, emin) in errs.items():
        tol = 1e-4 if field == "ssh" and layer == 1 else 1e-5
        assert emax < tol and emin < tol, (layer, field, emax, emin)


@pytest.mark.parametrize("name", ["bump_4x4", "double_gyre_4x4", "synth3_nop4_4x3", "synth_nop8_5layers_2x2", "synth3_curved_nop4_4x3"])
def test_oracle_matches_committed_fields(name):
    """Field-level fixtures generated by tests/make_golden.py pin the oracle against accidental change: mass-like fields to
    1e-12 relative L2, momentum-like fields to 1e-11 of their natural scale c*|dp| (round-off may differ with the number of
    OpenMP threads; anything larger is a change of the restated algorithm)."""
    import make_golden
    import os
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_fields_%s.npz" % name))
    got = make_golden.run(name)
    params, _ = make_golden.CASES[name]
    deck = hn.decks.build_deck(params)
    nl, npoin = deck["nlayers"], deck["npoin"]
    c = np.sqrt(deck["gravity"] * float(np.max(-deck["zbot_df"])))
    q, q0 = got["q_df"].reshape(nl, npoin, 3), ref["q_df"].reshape(nl, npoin, 3)
    qb, qb0 = got["qb_df"].reshape(npoin, 4), ref["qb_df"].reshape(npoin, 4)
    qp, qp0 = got["qprime_df"].reshape(nl, npoin, 3), ref["qprime_df"].reshape(nl, npoin, 3)
    for k in range(nl):
        dn = np.linalg.norm(q0[k, :, 0])
        assert np.linalg.norm(q[k, :, 0] - q0[k, :, 0]) <= 1e-12 * dn and np.linalg.norm(qp[k, :, 0] - qp0[k, :, 0]) <= 1e-12 * dn
        for v in (1, 2):
            assert np.linalg.norm(q[k, :, v] - q0[k, :, v]) <= 1e-11 * c * dn
            assert np.linalg.norm(qp[k, :, v] - qp0[k, :, v]) <= 1e-11 * c * np.sqrt(npoin)
    pn = np.linalg.norm(qb0[:, 0])
    assert np.linalg.norm(qb[:, :2] - qb0[:, :2]) <= 1e-12 * pn
    assert np.linalg.norm(qb[:, 2:] - qb0[:, 2:]) <= 1e-11 * c * pn


def test_oracle_shear_stress_properties():
    """rhs_layer_shear_stress (mod_create_rhs_mlswe.F90:146-279) as restated: the interface stresses telescope, so the stress
    tendencies of the layers sum to zero at every node (no stress through the surface, and -- parity hazard 2 -- none through the
    bottom); with one layer there is no interface at all; and ad_mlswe = 0 leaves the step untouched."""
    p = dict(hn.decks.synthetic_double_gyre(3, 3, nop=4, nlayers=3), ad_mlswe=1.0e6, max_shear_dz=2.0)
    O = oracle_lib.Oracle(p)
    assert O.step(2) == 0
    rs = O.shear_stress()
    assert np.abs(rs).max() > 0.0
    assert np.abs(rs.sum(axis=0)).max() <= 1e-12 * np.abs(rs).max()
    p0 = dict(p, ad_mlswe=0.0)
    A, B = oracle_lib.Oracle(p0), oracle_lib.Oracle(dict(p0, max_shear_dz=0.0))
    A.step(1); B.step(1)
    assert np.array_equal(A.get("q_df"), B.get("q_df"))
    assert not np.array_equal(A.get("q_df"), O.get("q_df"))
