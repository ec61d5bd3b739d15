"""Stage-kernel timing of one library build (GPU box): HNUMO_LIB_PATH=... python profiles/tools/stage_time.py [nel] [key=value ...]
Prints ms per barotropic stage (CUDA events around the stage loop inside the library), roofline fraction and the step time.
The physics result is not checked here (experimental builds may be timing-only)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hnumo_loader import hnumo_b200 as hn
nel = int(sys.argv[1]) if len(sys.argv) > 1 else 500
nop = int(os.environ.get("NOP", "4")); nl = int(os.environ.get("LAYERS", "3"))
if nop == 4:
    p = hn.decks.synthetic_double_gyre(nel, nel, nop=4, nlayers=nl, dt=12.0 * 1000.0 / nel, dt_btp=0.6 * (1 + 1e-9) * 1000.0 / nel)
else:
    p = hn.decks.synthetic_double_gyre(nel, nel, nop=nop, nlayers=nl)
deck = hn.decks.build_deck(p)
S = hn.Solver(deck, variant=0)
for kv in sys.argv[2:]:
    k, v = kv.split("="); S.set_option(k, float(v))
S.upload_state(deck["q_df"], deck["qb_df"], deck["qprime_df"])
rc = S.step(1); S.timing(reset=True)
rcs = [S.step(1) for _ in range(int(os.environ.get("STEPS", "2")))]
t = S.timing()
ngl = nop + 1; nq = 2 * nop + 1
W = 38.2 * ngl * ngl + 16 * nq * nq + 52 * nq + 52 * ngl
bpn = W * 8 / (ngl * ngl)
ms = t["ms_btp"] / t["stages"]
print(json.dumps(dict(lib=os.path.basename(hn.LIB_PATH), nel=nel, nop=nop, rc=[rc] + rcs, stage_ms=round(ms, 4), frac=round(bpn * deck["npoin"] / (ms * 1e-3) / 6456.2e9, 4),
                      step_ms=round(t["ms_step"] / t["steps"], 2), share=round(t["ms_btp"] / t["ms_step"], 4), opts=sys.argv[2:])))
S.close()
