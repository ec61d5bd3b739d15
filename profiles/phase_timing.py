"""Phase-level timeline of the element-record stage kernel from an instrumented build (-DHN_PAIR_TIMING).

Usage (GPU box): HNUMO_LIB_PATH=h-numo_b200/libhnumo_b200_timing.so python profiles/phase_timing.py [nel] [NE] [W]
Prints the median cycles a sampled warp spends between the time stamps of h-numo_b200/csrc/stage_pair.cuh."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hnumo_loader import hnumo_b200 as hn

nel = int(sys.argv[1]) if len(sys.argv) > 1 else 250
nop = int(os.environ.get("NOP", "4"))   # NOP=8: block-per-element form of the kernel
if nop == 4:
    p = hn.decks.synthetic_double_gyre(nel, nel, nop=4, nlayers=3, dt=12.0 * 1000.0 / nel, dt_btp=0.6 * (1 + 1e-9) * 1000.0 / nel)
else:
    p = hn.decks.synthetic_double_gyre(nel, nel, nop=nop, nlayers=int(os.environ.get("LAYERS", "3")))
deck = hn.decks.build_deck(p)
S = hn.Solver(deck, variant=0)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    S.set_option(k, float(v))
S.upload_state(deck["q_df"], deck["qb_df"], deck["qprime_df"])
S.step(1)
t = S.get_array("pair_timing").reshape(-1, 16)
t = t[t[:, 0] > 0]
names = ["P1 hdr+nodal", "P2 pass1", "P3 pass2+ldg", "P4 quad", "P5 scatter1", "P6 scat2+7a traces", "P7b face interp", "P7c face flux",
         "P7d proj", "P8 update", "P9 grad", "P9 trace out"]
idx = list(range(12)) + [15]
tot = np.median(t[:, 15] - t[:, 0])
print("sampled warps", len(t), "median cycles per unit", tot)
for k in range(12):
    d = t[:, idx[k + 1]] - t[:, idx[k]]
    print("%-22s median %8.0f  p10 %8.0f  p90 %8.0f  share %5.1f%%" % (names[k], np.median(d), np.percentile(d, 10), np.percentile(d, 90), 100 * np.median(d) / tot))
d = t[:, 12] - t[:, 1]
print("  (P2: header registers + L2 prefetch of traces/face sums: median %.0f)" % np.median(d))
S.close()
