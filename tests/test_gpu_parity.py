"""GPU parity tests (run on the B200 box): CUDA library through its C-ABI vs the CPU oracle on the same inputs.

Tolerances.  Mass-like fields (dp, dp', pb) must agree to 1e-10 relative L2.  Momentum-like fields are the result
of an O(1e8..1e13)-fold cancellation between volume and face pressure terms (tests/test_oracle_golden.py::
test_conditioning_noise_floor), so two correct FP64 evaluations with different summation order differ by ~1e-9..1e-7
in the plain relative sense; they are required to agree to 1e-11 of their natural scale c*|dp| (c = sqrt(g H), the
north-star tolerance applied to the scale of the field) and to 1e-6 plain relative L2.
"""
import threading

import numpy as np
import pytest

from parity_util import hn, make_pair, rel_l2, state_errors, sync_state_from_oracle
import oracle_lib

pytestmark = pytest.mark.gpu

def _nop8_small_step():
    base = hn.decks.synthetic_double_gyre(3, 4, nop=8, nlayers=2)
    return hn.decks.synthetic_double_gyre(3, 4, nop=8, nlayers=2, dt_btp=base["dt_btp"] / 2.5, dt=base["dt_btp"] * 4)


DECKS = {
    "bump": lambda: dict(hn.decks.SHIPPED["bump"]),
    "lake": lambda: dict(hn.decks.SHIPPED["lake"]),
    "double_gyre": lambda: dict(hn.decks.SHIPPED["double_gyre"], nelx=8, nely=8),
    "synth3": lambda: hn.decks.synthetic_double_gyre(8, 8, nop=4, nlayers=3),
    "noslip_rk3": lambda: dict(hn.decks.SHIPPED["double_gyre"], nelx=6, nely=5, x_boundary=(2, 2), kstages=3, botfr=2, cd_mlswe=1e-3),
    "nop3_5layers": lambda: dict(hn.decks.synthetic_double_gyre(5, 5, nop=3, nlayers=5)),
    # BASELINE config 5 in small: nop 8 runs the element-record kernel in its block-per-element form (variant 0)
    "nop8": lambda: dict(hn.decks.synthetic_double_gyre(4, 3, nop=8, nlayers=3)),
    # SSPRK(3,3) needs a smaller barotropic step than the shipped SSP(5,3) at the same resolution
    "nop8_noslip_drag2": lambda: dict(_nop8_small_step(), x_boundary=(2, 2), botfr=2, cd_mlswe=1e-3, kstages=3),
    "nop8_inviscid": lambda: dict(hn.decks.synthetic_double_gyre(3, 3, nop=8, nlayers=2), visc_mlswe=0.0, botfr=0),
    # the orders in between run the same block-per-element kernel (6x11, 7x13, 8x15 operators)
    "nop5": lambda: dict(hn.decks.synthetic_double_gyre(4, 3, nop=5, nlayers=3)),
    "nop6": lambda: dict(hn.decks.synthetic_double_gyre(3, 4, nop=6, nlayers=2), visc_mlswe=0.0),
    "nop7": lambda: dict(hn.decks.synthetic_double_gyre(3, 3, nop=7, nlayers=3), x_boundary=(2, 2)),
}
VARIANTS = [0, 1]  # 0: element-record kernel (default), 1: simple reference-form kernel (bisecting aid)


def natural_errors(S, O, deck):
    """errors of every prognostic field relative to its natural scale"""
    q, qb, qp = S.download_state()
    nl, npoin = deck["nlayers"], deck["npoin"]
    qo = O.get("q_df").reshape(nl, npoin, 3); qbo = O.get("qb_df").reshape(npoin, 4); qpo = O.get("qprime_df").reshape(nl, npoin, 3)
    c = np.sqrt(deck["gravity"] * float(np.max(-deck["zbot_df"])))
    out = {"mass": 0.0, "mom": 0.0}
    for k in range(nl):
        dn = np.linalg.norm(qo[k, :, 0])
        out["mass"] = max(out["mass"], np.linalg.norm(q[k, :, 0] - qo[k, :, 0]) / dn, np.linalg.norm(qp[k, :, 0] - qpo[k, :, 0]) / dn)
        for v in (1, 2):
            out["mom"] = max(out["mom"], np.linalg.norm(q[k, :, v] - qo[k, :, v]) / (c * dn),
                             np.linalg.norm(qp[k, :, v] - qpo[k, :, v]) / (c * np.sqrt(npoin)))
    pn = np.linalg.norm(qbo[:, 0])
    out["mass"] = max(out["mass"], np.linalg.norm(qb[:, 0] - qbo[:, 0]) / pn, np.linalg.norm(qb[:, 1] - qbo[:, 1]) / pn)
    for v in (2, 3):
        out["mom"] = max(out["mom"], np.linalg.norm(qb[:, v] - qbo[:, v]) / (c * pn))
    return out


@pytest.mark.parametrize("name", ["lake", "double_gyre"])
def test_static_derivations(name):
    """quantities the reference set-up derives (mod_initial_mlswe.F90) and the library re-derives on the device"""
    deck, S, O = make_pair(DECKS[name]())
    for nm in ["coriolis_quad", "tau_wind", "pbprime", "one_over_pbprime", "coeff_pbpert_L", "coeff_pbpert_R", "coeff_pbub_LR",
               "coeff_mass_pbpert_LR", "one_over_pbprime_edge", "pbprime_face", "zbot_face", "a_bcl", "b_bcl"]:
        assert rel_l2(S.get_array(nm), O.get(nm), floor=1e-30) < 1e-14, nm
    ref = O.get("grad_zbot_quad")
    assert np.linalg.norm(S.get_array("grad_zbot_quad") - ref) < 1e-12 * max(np.linalg.norm(ref), 1.0)
    S.close()


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("name", ["bump", "double_gyre", "synth3", "noslip_rk3", "nop3_5layers"])
def test_phase_parity(name, variant):
    """btp_bcl_coeffs_qdf, create_rhs_btp and ti_barotropic_ssprk_mlswe phase by phase on a developed state"""
    deck, S, O = make_pair(DECKS[name](), variant=variant)
    check_barotropic_phases(deck, S, O, name)
    S.close()


def check_barotropic_phases(deck, S, O, name):
    O.step(1)
    sync_state_from_oracle(S, O)
    O.btp_bcl_coeffs(); S.btp_bcl_coeffs()
    for nm in ["Q_uu_dp", "Q_uv_dp", "Q_vv_dp", "H_bcl", "Q_uu_dp_edge", "Q_uv_dp_edge", "Q_vv_dp_edge", "H_bcl_edge"]:
        ref = O.get(nm)
        assert rel_l2(S.get_array(nm), ref, floor=1e-12 * np.linalg.norm(O.get("Q_uu_dp" if nm.startswith("Q") and "edge" not in nm else nm)) + 1e-300) < 1e-12, nm
    if deck["visc_mlswe"] != 0.0:
        assert rel_l2(S.get_array("pbprime_visc"), O.get("pbprime_visc")) < 1e-14
        ref = O.get("btp_dpp_graduv")
        assert np.linalg.norm(S.get_array("btp_dpp_graduv") - ref) < 1e-6 * np.linalg.norm(ref) + 1e-14
    r_o, r_g = O.rhs_btp(), S.rhs_btp()       # evaluated by the stage kernel the solver runs (variant 0: element-record kernel)
    assert rel_l2(r_g[:, 0], r_o[:, 0], floor=1e-30) < 1e-11
    # momentum tendency = massinv x (volume - face) pressure terms of size H_bcl that cancel to ~1e-9 of themselves: the error of
    # any correct evaluation is round-off of those terms, eps * H_bcl * sqrt(massinv_max) (measured 2-3e-15 of that scale for both
    # stage kernels on every deck, profiles/probe_rhs.py); 2e-14 allows a factor ten
    Hn = np.linalg.norm(O.get("H_bcl")) / np.sqrt(O.npoin_q)
    scale = Hn * np.sqrt(deck["massinv"].max())
    for v in (1, 2):
        assert np.abs(r_g[:, v] - r_o[:, v]).max() <= 2e-14 * scale, (v, name, np.abs(r_g[:, v] - r_o[:, v]).max() / scale)
    O.btp_substeps(); S.btp_substeps()
    for nm in ["ope_ave", "H_ave", "ope2_ave", "ope2_ave_df", "H_face_ave", "ope_face_ave", "ope2_face_ave", "one_plus_eta_edge_2_ave"]:
        assert rel_l2(S.get_array(nm), O.get(nm)) < 1e-12, nm
    c = np.sqrt(deck["gravity"] * float(np.max(-deck["zbot_df"])))
    pbn = np.linalg.norm(O.get("qb_df").reshape(-1, 4)[:, 0]) / np.sqrt(O.npoin)
    for nm, scale in [("Qu_ave", c * c * pbn), ("Qv_ave", c * c * pbn), ("Quv_ave", c * c * pbn), ("btp_mass_flux_ave", c * pbn),
                      ("uvb_ave", c), ("uvb_ave_df", c), ("btp_mass_flux_face_ave", c * pbn), ("Qu_face_ave", c * c * pbn),
                      ("Qv_face_ave", c * c * pbn), ("uvb_face_ave", c), ("tau_bot_ave", c * pbn)]:
        a, b = S.get_array(nm), O.get(nm)
        assert np.linalg.norm(a - b) / np.sqrt(b.size) < 1e-11 * scale, nm
        assert rel_l2(a, b, floor=1e-30) < 1e-5 or np.linalg.norm(b) / np.sqrt(b.size) < 1e-9 * scale, nm
    if deck["visc_mlswe"] != 0.0:
        # graduvb_ave: time average of the LDG auxiliary variable grad(ub, vb) (mod_laplacian_quad.F90:56), compared directly; its
        # scale is c / dx with 1/dx ~ sqrt(massinv) (the default kernel derives it after the loop from the nodal velocity sums)
        a, b = S.get_array("graduvb_ave"), O.get("graduvb_ave")
        assert np.linalg.norm(a - b) / np.sqrt(b.size) < 1e-11 * c * np.sqrt(deck["massinv"].max()), "graduvb_ave"
        assert rel_l2(a, b, floor=1e-30) < 1e-5 or np.linalg.norm(b) / np.sqrt(b.size) < 1e-9 * c * np.sqrt(deck["massinv"].max()), "graduvb_ave"
    e = natural_errors(S, O, deck)
    assert e["mass"] < 1e-10 and e["mom"] < 1e-11, e


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("name", ["nop8", "nop8_noslip_drag2", "nop8_inviscid"])
def test_phase_parity_high_order(name, variant):
    test_phase_parity(name, variant)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("name,nsteps", [("nop8", 2), ("nop8_noslip_drag2", 2), ("nop8_inviscid", 2)])
def test_step_parity_high_order(name, nsteps, variant):
    test_step_parity(name, nsteps, variant)


@pytest.mark.parametrize("name", ["nop5", "nop6", "nop7"])
def test_intermediate_orders(name):
    """nop 5, 6, 7: element-record kernel in block-per-element form, phase by phase and over whole steps"""
    test_phase_parity(name, 0)
    test_step_parity(name, 2, 0)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("name,nsteps", [("bump", 5), ("lake", 5), ("double_gyre", 5), ("synth3", 5), ("noslip_rk3", 4), ("nop3_5layers", 4)])
def test_step_parity(name, nsteps, variant):
    """whole ti_rk_bcl steps from the initial conditions"""
    deck, S, O = make_pair(DECKS[name](), variant=variant)
    assert S.step(nsteps) == 0
    assert O.step(nsteps) == 0
    e = natural_errors(S, O, deck)
    assert e["mass"] < 1e-10 and e["mom"] < 1e-11, e
    plain = state_errors(S, O, deck)
    assert max(plain.values()) < 1e-6, plain
    S.close()


SHEAR_DECKS = {
    # ad_mlswe is taken large so that the term is visible next to the momentum (as written, the implicit system solves for
    # u = udp/dp against a diagonal of size dp, so the stress is ~1e-7 of what a dimensionally consistent system would give)
    "double_gyre": lambda: dict(hn.decks.SHIPPED["double_gyre"], nelx=6, nely=6, ad_mlswe=1.0e6, max_shear_dz=2.0),
    "synth3": lambda: dict(hn.decks.synthetic_double_gyre(6, 5, nop=4, nlayers=3), ad_mlswe=1.0e6, max_shear_dz=5.0),
    "bump": lambda: dict(hn.decks.SHIPPED["bump"], nelx=6, nely=6, ad_mlswe=1.0e4, max_shear_dz=0.5),     # f = 0: coeff = ad / (alpha dz)
    "nop8_5layers": lambda: dict(hn.decks.synthetic_double_gyre(3, 3, nop=8, nlayers=5), ad_mlswe=1.0e6, max_shear_dz=2.0),
}


@pytest.mark.parametrize("name", list(SHEAR_DECKS))
def test_vertical_shear_stress(name):
    """ad_mlswe > 0 (rhs_layer_shear_stress, mod_create_rhs_mlswe.F90:146-279, call sites
    mod_splitting.F90:139-164,247-271) against the oracle over whole steps, and the term really acts: the run differs from
    the run without it by far more than the tolerance."""
    p = SHEAR_DECKS[name]()
    deck, S, O = make_pair(p)
    n = 4
    assert S.step(n) == 0 and O.step(n) == 0
    e = natural_errors(S, O, deck)
    assert e["mass"] < 1e-10 and e["mom"] < 1e-11, e
    # the term by itself, on the same developed state: rhs_layer_shear_stress(q_df)
    sync_state_from_oracle(S, O)
    rs_o, rs_g = O.shear_stress(), S.layer_shear_stress()
    assert np.abs(rs_o).max() > 0.0
    assert rel_l2(rs_g, rs_o) < 1e-11, rel_l2(rs_g, rs_o)
    O0 = oracle_lib.Oracle(dict(p, ad_mlswe=0.0, affine_metrics=True))
    assert O0.step(n) == 0
    c = np.sqrt(deck["gravity"] * float(np.max(-deck["zbot_df"])))
    qa = O.get("q_df").reshape(deck["nlayers"], -1, 3); q0 = O0.get("q_df").reshape(deck["nlayers"], -1, 3)
    effect = max(np.linalg.norm(qa[k, :, v] - q0[k, :, v]) / np.linalg.norm(qa[k, :, v]) for k in range(deck["nlayers"]) for v in (1, 2))
    assert effect > 1e-8, effect      # relative change of the layer momentum by the stress (small: see the note at SHEAR_DECKS)
    S.close()


VISCQ_DECKS = {
    "double_gyre": lambda: dict(hn.decks.SHIPPED["double_gyre"], nelx=6, nely=6, method_visc=1),
    "synth3": lambda: dict(hn.decks.synthetic_double_gyre(6, 5, nop=4, nlayers=3), method_visc=1),
    "noslip_rk3": lambda: dict(hn.decks.SHIPPED["double_gyre"], nelx=5, nely=4, x_boundary=(2, 2), kstages=3, botfr=2, cd_mlswe=1e-3, method_visc=1),
    "nop3_5layers": lambda: dict(hn.decks.synthetic_double_gyre(4, 4, nop=3, nlayers=5), method_visc=1),
    "nop8": lambda: dict(hn.decks.synthetic_double_gyre(3, 3, nop=8, nlayers=3), method_visc=1),
    # a strong viscosity, so that an error in the term could not hide below the tolerance
    "strong": lambda: dict(hn.decks.synthetic_double_gyre(5, 5, nop=4, nlayers=2), method_visc=1, visc_mlswe=5.0e4),
}


@pytest.mark.parametrize("name", list(VISCQ_DECKS))
def test_quadrature_point_viscosity(name):
    """method_visc == 1 (btp_create_laplacian_v2 / bcl_create_laplacian_v2, mod_laplacian_quad.F90:125-223,
    252-355) against the oracle: the barotropic RHS and the layer momentum RHS on a developed state, then whole steps; and the
    result differs from the nodal form (method_visc == 3) by far more than the tolerance, i.e. the other code path really ran."""
    p = VISCQ_DECKS[name]()
    deck, S, O = make_pair(p)
    O.step(1)
    sync_state_from_oracle(S, O)
    O.btp_bcl_coeffs(); S.btp_bcl_coeffs()
    r_o, r_g = O.rhs_btp(), S.rhs_btp()
    assert rel_l2(r_g[:, 0], r_o[:, 0], floor=1e-30) < 1e-11
    Hn = np.linalg.norm(O.get("H_bcl")) / np.sqrt(O.npoin_q)
    scale = Hn * np.sqrt(deck["massinv"].max())
    for v in (1, 2):
        assert np.abs(r_g[:, v] - r_o[:, v]).max() <= 2e-14 * scale, (v, name, np.abs(r_g[:, v] - r_o[:, v]).max() / scale)
    O.btp_substeps(); S.btp_substeps()
    sync_state_from_oracle(S, O)          # same barotropic state on both sides for the layer phase (as in test_layer_phase_parity)
    m_o, m_g = O.layer_momentum_rhs(), S.layer_momentum_rhs()
    assert np.abs(m_g - m_o).max() <= 2e-14 * scale, (name, np.abs(m_g - m_o).max() / scale)
    n = 3
    assert S.step(n) == 0 and O.step(n) == 0
    e = natural_errors(S, O, deck)
    assert e["mass"] < 1e-10 and e["mom"] < 1e-11, e
    O3 = oracle_lib.Oracle(dict(p, method_visc=3, affine_metrics=True))
    assert O3.step(n + 1) == 0
    qa = O.get("q_df").reshape(deck["nlayers"], -1, 3); q3 = O3.get("q_df").reshape(deck["nlayers"], -1, 3)
    c = np.sqrt(deck["gravity"] * float(np.max(-deck["zbot_df"])))
    differ = max(np.linalg.norm(qa[k, :, v] - q3[k, :, v]) / (c * np.linalg.norm(qa[k, :, 0])) for k in range(deck["nlayers"]) for v in (1, 2))
    assert differ > 100 * e["mom"], (differ, e)
    S.close()


@pytest.mark.parametrize("nranks,partition", [(2, "rows"), (4, "morton")])
def test_quadrature_point_viscosity_partitioned(nranks, partition):
    """method_visc == 1 on a partition: the face values of the flux variable travel in one message per neighbour and stage
    (width nq; create_communicator_quad, src/create_rhs_communicator.F90:82-134).  Like the nodal form, the as-written LDG face
    flux takes the evaluating rank's element as the left one on a processor face, so the result depends on the partition at
    O(visc) -- the same tolerance as test_partitioned_equals_single."""
    params = dict(hn.decks.SHIPPED["double_gyre"], nelx=6, nely=8, method_visc=1)
    if partition != "rows":
        params["partition"] = partition
    single = hn.decks.build_deck(params)
    S = hn.Solver(single)
    S.upload_state(single["q_df"], single["qb_df"], single["qprime_df"])
    assert S.step(3) == 0
    q1, qb1, qp1 = S.download_state()
    S.close()
    decks, outs = _run_partitioned(params, nranks, 3, gid=700 + nranks)
    npts = single["npts"]
    c = np.sqrt(single["gravity"] * 9928.0)
    tol = 1e-10
    for d, (q, qb, qp) in zip(decks, outs):
        idx = (d["elem_global"][:, None] * npts + np.arange(npts)[None, :]).ravel()
        assert rel_l2(qb[:, 0], qb1[idx, 0]) < tol
        assert rel_l2(q[:, :, 0], q1[:, idx, 0]) < 10 * tol
        for v in (2, 3):
            assert np.linalg.norm(qb[:, v] - qb1[idx, v]) / (c * np.linalg.norm(qb1[idx, 0])) < tol
        for v in (1, 2):
            assert np.linalg.norm(q[:, :, v] - q1[:, idx, v]) / (c * np.linalg.norm(q1[:, idx, 0])) < tol


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("kstages", [1, 2, 4])
def test_ssprk_stage_counts(kstages, variant):
    """the SSPRK tables of mod_initial_mlswe.F90:652-678 other than the shipped (5,3) and the (3,3) of the no-slip deck: kstages = 1
    (forward Euler), 2 and 4 -- the load/store pattern of the SSPRK work states in the stage kernel differs for each.  The low-order
    schemes need a smaller barotropic step for stability."""
    base = hn.decks.synthetic_double_gyre(6, 6, nop=4, nlayers=3)
    cut = {1: 8.0, 2: 4.0, 4: 2.0}[kstages]
    p = dict(hn.decks.synthetic_double_gyre(6, 6, nop=4, nlayers=3, dt_btp=base["dt_btp"] / cut, dt=base["dt_btp"] / cut * 10), kstages=kstages)
    deck, S, O = make_pair(p, variant=variant)
    assert deck["kstages"] == kstages
    assert S.step(3) == 0 and O.step(3) == 0
    e = natural_errors(S, O, deck)
    assert e["mass"] < 1e-10 and e["mom"] < 1e-11, e
    S.close()


def test_high_order_many_layers():
    """BASELINE config 5 in small: nop=8 (ngl=9, nq=17), 10 layers -- orders without a compile-time instantiation run the
    run-time-size kernels; nl > 3 uses the intent semantics of SURVEY 8 hazard 1 in both the oracle and the library"""
    deck, S, O = make_pair(hn.decks.synthetic_double_gyre(3, 3, nop=8, nlayers=10))
    assert S.step(2) == 0 and O.step(2) == 0
    e = natural_errors(S, O, deck)
    assert e["mass"] < 1e-10 and e["mom"] < 1e-11, e
    S.close()


def test_100_steps_bump():
    """north-star check: 100 baroclinic steps (56000 barotropic stages) of the bump deck"""
    deck, S, O = make_pair(DECKS["bump"]())
    assert S.step(100) == 0 and O.step(100) == 0
    e = natural_errors(S, O, deck)
    assert e["mass"] < 1e-10 and e["mom"] < 1e-11, e
    plain = state_errors(S, O, deck)
    assert max(plain.values()) < 1e-6, plain
    S.close()


def test_drop_in_signature_matches_resident_stepping():
    """hnumo_ti_rk_bcl(q_df, qb_df, qprime_df) on host buffers == upload + step + download"""
    deck, S, O = make_pair(DECKS["double_gyre"]())
    q, qb, qp = deck["q_df"].copy(), deck["qb_df"].copy(), deck["qprime_df"].copy()
    for _ in range(2):
        assert S.ti_rk_bcl(q, qb, qp) == 0
    S.upload_state(deck["q_df"], deck["qb_df"], deck["qprime_df"])
    S.step(2)
    q2, qb2, qp2 = S.download_state()
    assert np.array_equal(q, q2) and np.array_equal(qb, qb2) and np.array_equal(qp, qp2)
    S.close()


def test_variants_agree_bitwise_on_mass():
    """the optimised stage kernel and the simple one integrate the same scheme"""
    p = DECKS["double_gyre"]()
    deck = hn.decks.build_deck(p)
    outs = []
    for variant in VARIANTS:
        S = hn.Solver(deck, variant=variant)
        S.upload_state(deck["q_df"], deck["qb_df"], deck["qprime_df"])
        S.step(3)
        outs.append(S.download_state())
        S.close()
    assert rel_l2(outs[0][1][:, 0], outs[1][1][:, 0]) < 1e-13
    assert rel_l2(outs[0][0][:, :, 0], outs[1][0][:, :, 0]) < 1e-12


def _run_partitioned(params, nranks, nsteps, gid, variant=0, options=()):
    decks = [hn.decks.build_deck(params, r, nranks) for r in range(nranks)]
    solvers = [hn.Solver(d, variant=variant) for d in decks]
    for S, d in zip(solvers, decks):
        S.comm_init(hn.local_group_id(gid))
        for k, v in options:
            S.set_option(k, v)
        S.upload_state(d["q_df"], d["qb_df"], d["qprime_df"])
    rcs = [None] * nranks

    def work(i):
        rcs[i] = solvers[i].step(nsteps)

    th = [threading.Thread(target=work, args=(i,)) for i in range(nranks)]
    [t.start() for t in th]
    [t.join(timeout=600) for t in th]
    assert all(rc == 0 for rc in rcs), rcs
    outs = [S.download_state() for S in solvers]
    [S.close() for S in solvers]
    return decks, outs


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("visc", [0.0, 50.0])
@pytest.mark.parametrize("nranks", [2, 4])
def test_partitioned_equals_single(nranks, visc, variant):
    """k-way element partition with face-halo exchange == 1-way; in-process back end on one GPU.

    With visc_mlswe == 0 every face term is antisymmetric under (L<->R, n->-n) and the two runs agree to round-off.
    With viscosity the reference's as-written LDG face flux (mod_laplacian_quad.F90:485-486) is NOT symmetric: each
    rank evaluates a processor face with itself as the left element, so the reference itself depends on the
    partition at O(visc).  The library reproduces the reference in both configurations; the tolerance is wider."""
    params = dict(hn.decks.SHIPPED["double_gyre"], nelx=6, nely=8, visc_mlswe=visc)
    tol = 1e-13 if visc == 0.0 else 1e-10
    single = hn.decks.build_deck(params)
    S = hn.Solver(single, variant=variant)
    S.upload_state(single["q_df"], single["qb_df"], single["qprime_df"])
    assert S.step(3) == 0
    q1, qb1, qp1 = S.download_state()
    S.close()
    decks, outs = _run_partitioned(params, nranks, 3, gid=100 + nranks + (10 if visc else 0) + 20 * variant, variant=variant)
    npts = single["npts"]
    c = np.sqrt(single["gravity"] * 9928.0)
    for d, (q, qb, qp) in zip(decks, outs):
        idx = (d["elem_global"][:, None] * npts + np.arange(npts)[None, :]).ravel()
        assert rel_l2(qb[:, 0], qb1[idx, 0]) < tol
        assert rel_l2(q[:, :, 0], q1[:, idx, 0]) < 10 * tol
        for v in (2, 3):
            assert np.linalg.norm(qb[:, v] - qb1[idx, v]) / (c * np.linalg.norm(qb1[idx, 0])) < tol
        for v in (1, 2):
            assert np.linalg.norm(q[:, :, v] - q1[:, idx, v]) / (c * np.linalg.norm(q1[:, idx, 0])) < tol


@pytest.mark.parametrize("visc", [0.0, 50.0])
@pytest.mark.parametrize("partition,nranks", [("blocks:2x2", 4), ("morton", 4), ("morton", 3), ("blocks:3x2", 6)])
def test_general_partitions_equal_single(partition, nranks, visc):
    """Partitions the way p4est produces them: 2-D blocks and chunks of the Morton curve -- ranks with up to
    four neighbours, non-contiguous boundary elements, several messages per exchange -- give the 1-rank result
    (to round-off without viscosity; see test_partitioned_equals_single for the O(visc) partition dependence of the
    reference's as-written LDG face flux)."""
    params = dict(hn.decks.SHIPPED["double_gyre"], nelx=6, nely=8, visc_mlswe=visc, partition=partition)
    tol = 1e-13 if visc == 0.0 else 1e-10
    single = hn.decks.build_deck(params)
    S = hn.Solver(single)
    S.upload_state(single["q_df"], single["qb_df"], single["qprime_df"])
    assert S.step(3) == 0
    q1, qb1, qp1 = S.download_state()
    S.close()
    gid = 500 + 10 * nranks + (1 if visc else 0) + (5 if partition == "morton" else 0)
    decks, outs = _run_partitioned(params, nranks, 3, gid=gid)
    npts = single["npts"]
    c = np.sqrt(single["gravity"] * 9928.0)
    assert sorted(np.concatenate([d["elem_global"] for d in decks]).tolist()) == list(range(48))
    for d, (q, qb, qp) in zip(decks, outs):
        idx = (d["elem_global"][:, None] * npts + np.arange(npts)[None, :]).ravel()
        assert rel_l2(qb[:, 0], qb1[idx, 0]) < tol
        assert rel_l2(q[:, :, 0], q1[:, idx, 0]) < 10 * tol
        for v in (2, 3):
            assert np.linalg.norm(qb[:, v] - qb1[idx, v]) / (c * np.linalg.norm(qb1[idx, 0])) < tol
        for v in (1, 2):
            assert np.linalg.norm(q[:, :, v] - q1[:, idx, v]) / (c * np.linalg.norm(q1[:, idx, 0])) < tol


@pytest.mark.parametrize("nranks", [2, 4])
def test_partitioned_high_order(nranks):
    """nop 8 on 2 and 4 partitions (block-per-element stage kernel, overlapped exchange) == 1 partition"""
    params = dict(hn.decks.synthetic_double_gyre(3, 4, nop=8, nlayers=2), visc_mlswe=0.0)
    single = hn.decks.build_deck(params)
    S = hn.Solver(single)
    S.upload_state(single["q_df"], single["qb_df"], single["qprime_df"])
    assert S.step(2) == 0
    q1, qb1, qp1 = S.download_state()
    S.close()
    decks, outs = _run_partitioned(params, nranks, 2, gid=400 + nranks)
    npts = single["npts"]
    c = np.sqrt(single["gravity"] * 9928.0)
    for d, (q, qb, qp) in zip(decks, outs):
        idx = (d["elem_global"][:, None] * npts + np.arange(npts)[None, :]).ravel()
        assert rel_l2(qb[:, 0], qb1[idx, 0]) < 1e-13
        for v in (2, 3):
            assert np.linalg.norm(qb[:, v] - qb1[idx, v]) / (c * np.linalg.norm(qb1[idx, 0])) < 1e-13


@pytest.mark.parametrize("nranks", [2, 4])
def test_overlapped_exchange_is_bitwise_the_serial_one(nranks):
    """the boundary elements advance on the exchange stream while the interior advances on the compute
    stream.  The per-element arithmetic does not change, so the result must be bit-identical to exchange-after-stage."""
    params = dict(hn.decks.synthetic_double_gyre(9, 8, nop=4, nlayers=3))
    _, on = _run_partitioned(params, nranks, 3, gid=300 + nranks, options=(("overlap", 1),))
    _, off = _run_partitioned(params, nranks, 3, gid=310 + nranks, options=(("overlap", 0),))
    for (q, qb, qp), (q0, qb0, qp0) in zip(on, off):
        assert np.array_equal(q, q0) and np.array_equal(qb, qb0) and np.array_equal(qp, qp0)


@pytest.mark.parametrize("lw", [47, 63, 0])
@pytest.mark.parametrize("name", ["bump", "double_gyre", "synth3", "noslip_rk3"])
def test_layer_phase_parity(name, lw):
    """layer_mass_rhs and layer_momentum_rhs (mod_create_rhs_mlswe.F90:28-78) through their own C-ABI entries, on a developed
    state with the time averages of one barotropic solve: a compensating error between Apply_layers_fluxes and the update
    would show here.  lw: warp-per-element layer kernels (47 default, 63 all of them) / block-per-element kernels (0)."""
    deck, S, O = make_pair(DECKS[name]())
    S.set_option("layer_warp", lw)
    check_layer_phases(deck, S, O)
    S.close()


def check_layer_phases(deck, S, O):
    O.step(1)
    sync_state_from_oracle(S, O)
    O.btp_bcl_coeffs(); S.btp_bcl_coeffs()
    O.btp_substeps(); S.btp_substeps()
    sync_state_from_oracle(S, O)          # same barotropic state on both sides for the layer phases; the averages stay each side's own
    c = np.sqrt(deck["gravity"] * float(np.max(-deck["zbot_df"])))
    qo = O.get("q_df").reshape(deck["nlayers"], deck["npoin"], 3)
    a, b = S.layer_mass_rhs(), O.layer_mass_rhs()
    # dp_advec = -div(u dp): scale c * dp / dx, with 1/dx ~ sqrt(massinv) * sqrt(mean weight)
    hscale = np.sqrt(deck["massinv"].max())
    for k in range(deck["nlayers"]):
        scale = c * np.abs(qo[k, :, 0]).max() * hscale
        assert np.abs(a[k] - b[k]).max() <= 1e-11 * scale, (k, np.abs(a[k] - b[k]).max(), scale)
        assert rel_l2(a[k], b[k], floor=1e-8 * scale * np.sqrt(deck["npoin"])) < 1e-6
    a, b = S.layer_momentum_rhs(), O.layer_momentum_rhs()
    Hn = np.linalg.norm(O.get("H_bcl")) / np.sqrt(O.npoin_q)          # size of the cancelling pressure terms
    for k in range(deck["nlayers"]):
        for v in (0, 1):
            assert np.abs(a[k, :, v] - b[k, :, v]).max() <= 2e-14 * Hn * hscale, (k, v, np.abs(a[k, :, v] - b[k, :, v]).max(), Hn * hscale)


@pytest.mark.parametrize("name", ["synth3", "double_gyre"])
def test_batched_momentum_volume_equals_the_layerwise_one(name):
    """k_mom_volume_b (all layers of an element in one sweep) against k_mom_volume (layer by layer) on the same developed state: the
    same per-field arithmetic up to the compiler's choice of FMA contractions -- the layer momentum RHS agrees to round-off of the
    cancelling pressure terms (ten times tighter than the bound against the oracle in test_layer_phase_parity), bitwise at 2 layers"""
    deck = hn.decks.build_deck(DECKS[name]())
    S = hn.Solver(deck)
    S.upload_state(deck["q_df"], deck["qb_df"], deck["qprime_df"])
    assert S.step(1) == 0
    S.btp_bcl_coeffs(); S.btp_substeps()
    Hq = S.get_array("H_bcl")
    scale = np.linalg.norm(Hq) / np.sqrt(Hq.size) * np.sqrt(deck["massinv"].max())
    rhs = []
    for batched in (1, 0):
        S.set_option("mom_volume_batched", batched)
        rhs.append(S.layer_momentum_rhs().copy())
    assert np.abs(rhs[0]).max() > 0.0
    assert np.abs(rhs[0] - rhs[1]).max() <= 2e-15 * scale, (np.abs(rhs[0] - rhs[1]).max(), scale)
    if deck["nlayers"] == 2:
        assert np.array_equal(rhs[0], rhs[1])
    S.close()


@pytest.mark.parametrize("partition,nranks", [("rows", 2), ("morton", 4), ("blocks:3x2", 6)])
def test_halo_exchange_pattern(partition, nranks):
    """hnumo_halo_exchange (the device replacement of create_nbhs_face_df / send_bound_dg_general_df): every rank sends a nodal
    field that encodes (global element, node); what arrives for a processor face must be the code of the neighbour element's
    node on the other side of that face -- checked against the single-partition face table."""
    params = dict(hn.decks.SHIPPED["double_gyre"], nelx=6, nely=6, partition=partition)
    single = hn.decks.build_deck(params)
    npts, ngl = single["npts"], single["ngl"]
    decks = [hn.decks.build_deck(params, r, nranks) for r in range(nranks)]
    solvers = [hn.Solver(d) for d in decks]
    gid = 900 + nranks
    for S in solvers:
        S.comm_init(hn.local_group_id(gid))
    fields = []
    for d in decks:
        code = (d["elem_global"][:, None] * 1000.0 + np.arange(npts)[None, :]).ravel()
        fields.append(np.stack([code, -code], axis=1))
    outs = [None] * nranks

    def work(i):
        outs[i] = solvers[i].halo_exchange(fields[i])

    th = [threading.Thread(target=work, args=(i,)) for i in range(nranks)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    # neighbour across every interior face of the single-partition mesh: (left elem, left face nodes) <-> (right elem, right face nodes)
    face = single["face"]

    def fnodes(iloc):   # local face 3..6 -> nodal indices along the face (mod_grid / create_normals.F90:264-296)
        n = np.arange(ngl)
        return {3: n, 4: (ngl - 1) * ngl + n, 5: n * ngl, 6: n * ngl + ngl - 1}[int(iloc)]

    expect = {}
    for f in range(single["nface"]):
        il, ir, el, er = face[f, 4], face[f, 5], face[f, 6] - 1, face[f, 7] - 1
        if face[f, 7] > 0:
            expect[(el, int(il))] = er * 1000.0 + fnodes(ir)
            expect[(er, int(ir))] = el * 1000.0 + fnodes(il)
    nchecked = 0
    for d, out in zip(decks, outs):
        assert out is not None and out.shape[0] == len(d["nbh_send_recv"])
        for hidx, lf in enumerate(d["nbh_send_recv"]):
            F = d["face"][lf - 1]
            eg = int(d["elem_global"][F[6] - 1])
            want = expect[(eg, int(F[4]))]
            assert np.array_equal(out[hidx, :, 0], want) and np.array_equal(out[hidx, :, 1], -want), (eg, F, out[hidx, :, 0], want)
            nchecked += 1
    assert nchecked > 0
    [S.close() for S in solvers]
