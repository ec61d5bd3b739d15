import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from parity_util import *
p = dict(hn.decks.synthetic_double_gyre(int(os.environ.get("NX", "3")), int(os.environ.get("NY", "3")), nop=8, nlayers=2), visc_mlswe=float(os.environ.get("VISC", "0")), botfr=int(os.environ.get("BOTFR", "0")))
p["kstages"] = int(os.environ.get("KST", "1")); p["dt"] = p["dt_btp"] * int(os.environ.get("NBTP", "1"))
deck, S, O = make_pair(p, variant=int(os.environ.get("VARIANT", "0")))
O.btp_bcl_coeffs(); S.btp_bcl_coeffs()
O.btp_substeps(); S.btp_substeps()
q, qb, qp = S.download_state()
qbo = O.get("qb_df").reshape(-1, 4)
npts = deck["npts"]
for v in range(4):
    bad = np.isnan(qb[:, v])
    print("var", v, "nan", int(bad.sum()), "rel err", rel_l2(qb[~bad, v], qbo[~bad, v], floor=1e-30))
    if bad.any():
        idx = np.nonzero(bad)[0]
        print("   elements", sorted(set((idx // npts).tolist())), "nodes of first elem", (idx[idx // npts == idx[0] // npts] % npts).tolist()[:40])
    else:
        d = np.abs(qb[:, v] - qbo[:, v]).reshape(-1, npts)
        print("   worst element", int(d.max(axis=1).argmax()), "node", int(d.max(axis=0).argmax()), "abs", d.max(), "scale", np.abs(qbo[:, v]).max())
S.close()
