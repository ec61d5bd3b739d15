"""GPU parity on general (curved, non-affine) quadrilaterals -- SURVEY 8(f) rank 4, "non-brick quads with per-point metrics".

The reference's metric, normal and operator-table code (metrics.F90, metrics_quad.F90, create_normals(_quad).F90,
Tensor_product.F90) is written for isoparametric quadrilaterals; its shipped cases are bricks.  The decks here are bricks whose
nodes are displaced smoothly (`mesh_warp`, h-numo_b200/decks.py and oracle Config::mesh_warp): every element is curved, the
metric terms vary by tens of percent inside an element.  The library is handed the per-point geometry (hnumo_desc_t
point_metrics_q, point_metrics, face_geom_q, face_geom_n, coord) and runs its run-time-size kernels; the oracle evaluates the
reference's dense-table form on the same mesh.  Tolerances as in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from parity_util import hn, make_pair, rel_l2, state_errors, sync_state_from_oracle
from test_gpu_parity import natural_errors, _run_partitioned, check_barotropic_phases, check_layer_phases

pytestmark = pytest.mark.gpu

W = 0.15
DECKS = {
    "synth3": lambda: dict(hn.decks.synthetic_double_gyre(6, 5, nop=4, nlayers=3), mesh_warp=W),
    "double_gyre": lambda: dict(hn.decks.SHIPPED["double_gyre"], nelx=6, nely=6, mesh_warp=W),
    "bump": lambda: dict(hn.decks.SHIPPED["bump"], nelx=6, nely=6, mesh_warp=W),                      # no viscosity, f = 0
    "noslip_rk3": lambda: dict(hn.decks.SHIPPED["double_gyre"], nelx=5, nely=6, x_boundary=(2, 2), kstages=3, botfr=2, cd_mlswe=1e-3, mesh_warp=W),
    "nop3_5layers": lambda: dict(hn.decks.synthetic_double_gyre(5, 4, nop=3, nlayers=5), mesh_warp=-0.1),
    "nop6": lambda: dict(hn.decks.synthetic_double_gyre(3, 4, nop=6, nlayers=2), mesh_warp=W),
}


@pytest.mark.parametrize("name,nsteps", [("synth3", 4), ("double_gyre", 4), ("bump", 5), ("noslip_rk3", 4), ("nop3_5layers", 3), ("nop6", 2)])
def test_step_parity_general_quads(name, nsteps):
    """whole ti_rk_bcl steps on curved elements; the mesh matters (the result differs from the brick result by far more than the tolerance)"""
    p = DECKS[name]()
    deck, S, O = make_pair(p)
    assert deck["point_metrics_q"] is not None and deck["elem_metrics"] is None
    assert S.step(nsteps) == 0 and O.step(nsteps) == 0
    e = natural_errors(S, O, deck)
    assert e["mass"] < 1e-10 and e["mom"] < 1e-11, e
    assert max(state_errors(S, O, deck).values()) < 1e-6
    d = S.diagnostics()
    assert np.isfinite(d["cfl"]) and d["min_dx"] > 0 and d["min_dy"] > 0
    qg = S.download_state()[0]
    S.close()
    deck0, S0, O0 = make_pair(dict(p, mesh_warp=0.0))
    assert S0.step(nsteps) == 0
    q0 = S0.download_state()[0]
    S0.close()
    assert rel_l2(qg[:, :, 0] - deck["q_df"][:, :, 0], q0[:, :, 0] - deck0["q_df"][:, :, 0]) > 1e-3 or name == "bump"


@pytest.mark.parametrize("name", ["synth3", "noslip_rk3", "bump", "nop3_5layers"])
def test_phase_parity_general_quads(name):
    """btp_bcl_coeffs_qdf, create_rhs_btp, the time averages of one barotropic solve, layer_mass_rhs and layer_momentum_rhs on curved
    elements, phase by phase (the checks of test_phase_parity and test_layer_phase_parity)"""
    deck, S, O = make_pair(DECKS[name]())
    check_barotropic_phases(deck, S, O, name)
    S.close()
    deck, S, O = make_pair(DECKS[name]())
    check_layer_phases(deck, S, O)
    S.close()


@pytest.mark.parametrize("visc", [0.0, 50.0])
@pytest.mark.parametrize("partition,nranks", [("rows", 2), ("morton", 4)])
def test_partitioned_general_quads(partition, nranks, visc):
    """k-way partition of a curved mesh == 1-way (the per-point face geometry of a processor face is each rank's own copy)"""
    params = dict(hn.decks.SHIPPED["double_gyre"], nelx=6, nely=8, visc_mlswe=visc, partition=partition, mesh_warp=W)
    tol = 1e-13 if visc == 0.0 else 1e-10
    single = hn.decks.build_deck(params)
    S = hn.Solver(single)
    S.upload_state(single["q_df"], single["qb_df"], single["qprime_df"])
    assert S.step(3) == 0
    q1, qb1, qp1 = S.download_state()
    S.close()
    decks, outs = _run_partitioned(params, nranks, 3, gid=900 + 10 * nranks + (1 if visc else 0))
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


def test_general_quads_keep_mass_and_reject_the_affine_kernels():
    """mass conservation to round-off on curved elements (check.F90:58); the element-record and warp-per-element kernels carry one
    Jacobian per element and are refused; a descriptor with only part of the per-point geometry is refused"""
    p = DECKS["synth3"]()
    deck = hn.decks.build_deck(p)
    S = hn.Solver(deck)
    S.upload_state(deck["q_df"], deck["qb_df"], deck["qprime_df"])
    m0 = S.diagnostics()["mass"].copy()
    assert S.step(5) == 0
    assert np.abs(S.diagnostics()["mass"] / m0 - 1).max() < 1e-13
    for key, val in (("stage_kernel_variant", 0), ("layer_warp", 47)):
        with pytest.raises(hn.HnumoError):
            S.set_option(key, val)
    S.close()
    bad = dict(deck); bad["face_geom_n"] = None
    with pytest.raises(hn.HnumoError):
        hn.Solver(bad)
