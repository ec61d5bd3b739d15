"""The plain-C caller of the C-ABI (tests/c_abi_harness.c, built with gcc, no Python in the call path) gives bitwise the
result of the ctypes path: the boundary does not depend on who binds it."""
import os
import subprocess

import numpy as np
import pytest

import harness_util
from parity_util import hn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,params", [("double_gyre", dict(nelx=6, nely=5)), ("bump", dict())])
def test_c_harness_equals_ctypes_path(name, params, tmp_path):
    deck = hn.decks.build_deck(dict(hn.decks.SHIPPED[name], **params))
    exe = harness_util.build_harness()
    deck_file, out_file = str(tmp_path / "deck.bin"), str(tmp_path / "out.bin")
    harness_util.write_deck(deck_file, deck)
    r = subprocess.run([exe, deck_file, out_file, "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    q, qb, qp, diag = harness_util.read_out(out_file, deck)
    S = hn.Solver(deck)
    a, b, c = deck["q_df"].copy(), deck["qb_df"].copy(), deck["qprime_df"].copy()
    for _ in range(3):
        assert S.ti_rk_bcl(a, b, c) == 0
    S.close()
    assert np.array_equal(q, a) and np.array_equal(qb, b) and np.array_equal(qp, c)
    assert np.isfinite(diag).all() and diag[0] > 0.0
