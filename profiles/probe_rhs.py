"""GPU probe (not a test): measured errors of hnumo_rhs_btp vs the oracle per deck and stage-kernel variant, and the
effect of CUDA-graph replay on the launch-bound decks.  python profiles/probe_rhs.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import hn, make_pair, sync_state_from_oracle
from test_gpu_parity import DECKS

for name in ["bump", "lake", "double_gyre", "synth3", "noslip_rk3", "nop3_5layers", "nop8"]:
    for variant in (0, 1):
        deck, S, O = make_pair(DECKS[name](), variant=variant)
        O.step(1); sync_state_from_oracle(S, O)
        O.btp_bcl_coeffs(); S.btp_bcl_coeffs()
        r_o, r_g = O.rhs_btp(), S.rhs_btp()
        Hn = np.linalg.norm(O.get("H_bcl")) / np.sqrt(O.npoin_q)
        mi = deck["massinv"]; h = np.sqrt(1.0 / mi.max())
        out = []
        for v in range(3):
            out.append("%.2e/%.2e" % (np.linalg.norm(r_g[:, v] - r_o[:, v]) / max(np.linalg.norm(r_o[:, v]), 1e-300), np.abs(r_g[:, v] - r_o[:, v]).max()))
        # scale of the cancelling terms of the momentum tendency: massinv * (face weight) * H_bcl ~ H_bcl / (w_min * dx)
        print(name, "variant", variant, "rhs rel/maxabs:", out, "Hn=%.3e max|rhs_mom|=%.3e  Hn*sqrt(massinv_max)=%.3e" % (Hn, np.abs(r_o[:, 1:]).max(), Hn * np.sqrt(mi.max())))
        S.close()

for name in ["bump", "lake", "double_gyre"]:
    deck = hn.decks.build_deck(dict(hn.decks.SHIPPED[name]))
    for ug in (0, 1):
        S = hn.Solver(deck)
        S.set_option("use_graph", ug)
        S.upload_state(deck["q_df"], deck["qb_df"], deck["qprime_df"])
        S.step(2)
        t0 = time.perf_counter(); S.step(10); t1 = time.perf_counter()
        q, qb, qp = S.download_state()
        print(name, "use_graph", ug, "ms/step %.3f" % (100 * (t1 - t0)), "checksum %.17g" % float(np.sum(qb[:, 1])))
        S.close()
