"""GPU parity at the edges of the parameter space the C-ABI accepts: lowest orders (nop 1, 2: run-time-size kernels), inexact
integration (nq = 2 nop - 1, mod_basis.F90:76-79), the maximum layer count of the reference's lake case (20,
initial_conditions.F90:130-169), one-element and one-element-wide meshes (every side a wall), one element per rank."""
import numpy as np
import pytest

from parity_util import hn, make_pair, rel_l2, state_errors
from test_gpu_parity import natural_errors, _run_partitioned

pytestmark = pytest.mark.gpu


def _small_dt(p, cut):
    return dict(p, dt_btp=p["dt_btp"] / cut, dt=p["dt"] / cut)


CASES = {
    "nop1": lambda: hn.decks.synthetic_double_gyre(6, 5, nop=1, nlayers=3),
    "nop2_noslip": lambda: dict(hn.decks.synthetic_double_gyre(5, 6, nop=2, nlayers=2), y_boundary=(2, 2)),
    "inexact_nop4": lambda: dict(_small_dt(hn.decks.synthetic_double_gyre(5, 5, nop=4, nlayers=3), 2.0), dg_integ_exact=False),
    "inexact_nop3_novisc": lambda: dict(_small_dt(hn.decks.synthetic_double_gyre(4, 5, nop=3, nlayers=2), 2.0), dg_integ_exact=False, visc_mlswe=0.0),
    "lake_20_layers": lambda: dict(hn.decks.SHIPPED["lake"], nelx=5, nely=5, nlayers=20),
    "bump_one_element": lambda: dict(hn.decks.SHIPPED["bump"], nelx=1, nely=1, xdims=(0.0, 200.0), ydims=(0.0, 200.0)),
    "gyre_strip_1x6": lambda: hn.decks.synthetic_double_gyre(1, 6, nop=4, nlayers=3),
    # largest per-element working set the run-time-size kernels see: nop 8, 20 layers, vertical shear stress, quadrature-point viscosity
    "nop8_20_layers_shear_viscq": lambda: dict(hn.decks.synthetic_double_gyre(2, 2, nop=8, nlayers=20), ad_mlswe=1.0e6, max_shear_dz=2.0, method_visc=1),
    "nop8_20_layers": lambda: hn.decks.synthetic_double_gyre(2, 3, nop=8, nlayers=20),
    "gyre_strip_5x1_noslip": lambda: dict(hn.decks.synthetic_double_gyre(5, 1, nop=4, nlayers=2), x_boundary=(2, 2), y_boundary=(2, 2)),
}


@pytest.mark.parametrize("name", list(CASES))
def test_edge_case_step_parity(name):
    deck, S, O = make_pair(CASES[name]())
    n = 2 if name.startswith("nop8") else 3
    assert S.step(n) == 0 and O.step(n) == 0
    e = natural_errors(S, O, deck)
    assert e["mass"] < 1e-10 and e["mom"] < 1e-11, e
    assert max(state_errors(S, O, deck).values()) < 1e-6
    if name == "lake_20_layers":      # at rest over the bump of the bottom: stays at rest (docs/source/test.rst:14-43)
        d = S.diagnostics()
        assert np.abs(d["u"]).max() < 1e-9 and np.abs(d["v"]).max() < 1e-9
    S.close()


def test_one_element_per_rank():
    """2 x 2 elements on 4 ranks: every element is a boundary element of its rank, two processor faces and two walls each"""
    params = dict(hn.decks.synthetic_double_gyre(2, 2, nop=4, nlayers=3), partition="morton")
    single = hn.decks.build_deck(params)
    S = hn.Solver(single)
    S.upload_state(single["q_df"], single["qb_df"], single["qprime_df"])
    assert S.step(2) == 0
    q1, qb1, qp1 = S.download_state()
    S.close()
    decks, outs = _run_partitioned(params, 4, 2, gid=1300)
    assert [d["nelem"] for d in decks] == [1, 1, 1, 1]
    npts = single["npts"]
    c = np.sqrt(single["gravity"] * 9928.0)
    for d, (q, qb, qp) in zip(decks, outs):
        idx = (d["elem_global"][:, None] * npts + np.arange(npts)[None, :]).ravel()
        assert rel_l2(qb[:, 0], qb1[idx, 0]) < 1e-10
        for v in (2, 3):
            assert np.linalg.norm(qb[:, v] - qb1[idx, v]) / (c * np.linalg.norm(qb1[idx, 0])) < 1e-10
