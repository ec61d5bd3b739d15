"""Regenerates tests/golden/oracle_fields_*.npz: field-level vectors of the CPU oracle on small decks after a few
baroclinic steps.  They pin the oracle against accidental change (tests/test_oracle_golden.py::test_oracle_matches_committed_fields).

Usage: python tests/make_golden.py        (run from the repository root; needs only the CPU oracle)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib  # noqa: E402
from hnumo_loader import hnumo_b200 as hn  # noqa: E402

CASES = {
    # name: (deck parameters, baroclinic steps)
    "bump_4x4": (dict(hn.decks.SHIPPED["bump"], nelx=4, nely=4), 3),
    "double_gyre_4x4": (dict(hn.decks.SHIPPED["double_gyre"], nelx=4, nely=4), 3),
    "synth3_nop4_4x3": (hn.decks.synthetic_double_gyre(4, 3, nop=4, nlayers=3), 2),
    "synth_nop8_5layers_2x2": (hn.decks.synthetic_double_gyre(2, 2, nop=8, nlayers=5), 1),
    # curved (non-affine) elements: per-point metric terms, face normals and Jacobians (Config::mesh_warp)
    "synth3_curved_nop4_4x3": (dict(hn.decks.synthetic_double_gyre(4, 3, nop=4, nlayers=3), mesh_warp=0.15), 2),
}


def run(name):
    params, nsteps = CASES[name]
    o = oracle_lib.Oracle(params)
    assert o.step(nsteps) == 0
    return dict(q_df=o.get("q_df").copy(), qb_df=o.get("qb_df").copy(), qprime_df=o.get("qprime_df").copy(), nsteps=np.int64(nsteps))


if __name__ == "__main__":
    for name in CASES:
        out = os.path.join(HERE, "golden", "oracle_fields_%s.npz" % name)
        np.savez_compressed(out, **run(name))
        print("wrote", out, os.path.getsize(out), "bytes")
