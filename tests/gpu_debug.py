"""Manual GPU debugging aid: prints phase-by-phase discrepancies between the CUDA library and the CPU oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from parity_util import *

def report(tag, a, b, floor=0.0):
    print("  %-28s rel_l2 = %.3e   (|ref| = %.3e)" % (tag, rel_l2(a, b, floor), np.linalg.norm(b)))

def run(name, params, nsteps, variant=1):
    print("=== deck", name, "variant", variant)
    deck, S, O = make_pair(params, variant=variant)
    # statics
    for nm in ["coriolis_quad", "tau_wind", "grad_zbot_quad", "pbprime", "one_over_pbprime", "coeff_pbpert_L", "coeff_pbpert_R",
               "coeff_pbub_LR", "coeff_mass_pbpert_LR", "one_over_pbprime_edge", "pbprime_face", "zbot_face", "a_bcl", "b_bcl"]:
        report(nm, S.get_array(nm), O.get(nm), floor=1e-30)
    # advance the oracle one step to get a non-trivial state, then sync
    O.step(1)
    sync_state_from_oracle(S, O)
    O.btp_bcl_coeffs(); S.btp_bcl_coeffs()
    for nm in ["Q_uu_dp", "Q_uv_dp", "Q_vv_dp", "H_bcl", "Q_uu_dp_edge", "Q_uv_dp_edge", "Q_vv_dp_edge", "H_bcl_edge", "btp_dpp_graduv", "pbprime_visc"]:
        report(nm, S.get_array(nm), O.get(nm), floor=1e-30)
    r_o = O.rhs_btp()  # note: updates oracle accumulators only
    r_g = S.rhs_btp()
    for v in range(3):
        report("rhs_btp[%d]" % v, r_g[:, v], r_o[:, v], floor=1e-30)
    O.btp_substeps(); S.btp_substeps()
    for nm in ["ope_ave", "H_ave", "Qu_ave", "Qv_ave", "Quv_ave", "ope2_ave", "btp_mass_flux_ave", "uvb_ave", "tau_bot_ave", "ope2_ave_df",
               "uvb_ave_df", "graduvb_ave", "btp_mass_flux_face_ave", "H_face_ave", "Qu_face_ave", "Qv_face_ave", "ope_face_ave",
               "ope2_face_ave", "one_plus_eta_edge_2_ave", "uvb_face_ave"]:
        report(nm, S.get_array(nm), O.get(nm), floor=1e-30)
    e = state_errors(S, O, deck)
    print("  after btp_substeps:", {k: "%.2e" % v for k, v in e.items() if k.startswith("pb")})
    # full steps from the synced state
    sync_state_from_oracle(S, O)
    for i in range(nsteps):
        O.step(1); rc = S.step(1)
        e = state_errors(S, O, deck)
        print("  step %d rc=%d max err %.3e  worst=%s" % (i + 1, rc, max(e.values()), max(e, key=e.get)))
    print("  ", {k: "%.2e" % v for k, v in e.items()})
    print("  timing", S.timing())
    S.close()

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    variant = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    run("bump", dict(hn.decks.SHIPPED["bump"]), n, variant)
    run("lake", dict(hn.decks.SHIPPED["lake"]), n, variant)
    p = dict(hn.decks.SHIPPED["double_gyre"]); p.update(nelx=8, nely=8)
    run("double_gyre 8x8", p, n, variant)
    p = hn.decks.synthetic_double_gyre(8, 8, nop=4, nlayers=3); 
    run("synth 3 layers", p, n, variant)
