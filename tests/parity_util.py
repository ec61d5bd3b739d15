"""Helpers shared by the GPU parity tests: run the CPU oracle and the CUDA library on the same deck."""
import numpy as np

import oracle_lib
from hnumo_loader import hnumo_b200 as hn


def rel_l2(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    den = max(np.linalg.norm(b), floor)
    if den == 0.0:
        return float(np.linalg.norm(a - b))
    return float(np.linalg.norm(a - b) / den)


def make_pair(params, variant=0, oracle_metrics=True, affine=True):
    """Same inputs for both sides.  The oracle differentiates the node coordinates numerically like the reference
    (metrics.F90), which leaves ~1e-14 relative round-off in its Jacobians; with oracle_metrics the library is
    given those numbers (first point of each element / face), as the Fortran shim would do."""
    deck = hn.decks.build_deck(params)
    params = dict(params)
    general = params.get("mesh_warp", 0.0) != 0.0
    params["affine_metrics"] = affine and not general
    O = oracle_lib.Oracle(params)
    if general and oracle_metrics:
        # general quadrilaterals: the per-point geometry of the oracle (= of the reference's metrics / normals code), as the Fortran
        # shim would pass mod_metrics and mod_face
        nq, ngl, nf = O.nq, O.ngl, O.nface
        deck["point_metrics_q"] = np.stack([O.get(k) for k in ("ksiq_x", "ksiq_y", "etaq_x", "etaq_y", "jacq")], axis=1)
        deck["point_metrics"] = np.stack([O.get(k) for k in ("ksi_x", "ksi_y", "eta_x", "eta_y", "jac")], axis=1)
        deck["face_geom_q"] = np.concatenate([O.get("normal_vector_q").reshape(nf, nq, 2), O.get("jac_faceq").reshape(nf, nq, 1)], axis=2)
        deck["face_geom_n"] = np.concatenate([O.get("normal_vector").reshape(nf, ngl, 2), O.get("jac_face").reshape(nf, ngl, 1)], axis=2)
        deck["massinv"] = O.get("massinv")
    elif oracle_metrics:
        nq2, nq = O.nq * O.nq, O.nq
        em = deck["elem_metrics"]
        em[:, 0] = O.get("ksiq_x")[::nq2]; em[:, 1] = O.get("ksiq_y")[::nq2]
        em[:, 2] = O.get("etaq_x")[::nq2]; em[:, 3] = O.get("etaq_y")[::nq2]
        wnq = O.get("wnq")
        em[:, 4] = O.get("jacq")[::nq2] / (wnq[0] * wnq[0])
        fg = deck["face_geom"]
        nv = O.get("normal_vector_q").reshape(O.nface, nq, 2)
        fg[:, 0] = nv[:, 0, 0]; fg[:, 1] = nv[:, 0, 1]
        fg[:, 2] = O.get("jac_faceq").reshape(O.nface, nq)[:, 0] / wnq[0]
        deck["massinv"] = O.get("massinv")
    S = hn.Solver(deck, variant=variant)
    S.upload_state(deck["q_df"], deck["qb_df"], deck["qprime_df"])
    return deck, S, O


def sync_state_from_oracle(S, O):
    S.upload_state(O.get("q_df"), O.get("qb_df"), O.get("qprime_df"))


def state_errors(S, O, deck):
    """per-field relative L2 errors of (q_df, qb_df, qprime_df); momentum-like fields are scaled by the
    natural magnitude c*dp so that fields at rest do not divide by zero."""
    q, qb, qp = S.download_state()
    nl, npoin = deck["nlayers"], deck["npoin"]
    qo = O.get("q_df").reshape(nl, npoin, 3)
    qbo = O.get("qb_df").reshape(npoin, 4)
    qpo = O.get("qprime_df").reshape(nl, npoin, 3)
    g = deck["gravity"]
    H = float(np.max(-deck["zbot_df"]))
    c = np.sqrt(g * H)
    errs = {}
    for k in range(nl):
        dpn = np.linalg.norm(qo[k, :, 0])
        errs["dp[%d]" % k] = rel_l2(q[k, :, 0], qo[k, :, 0])
        errs["udp[%d]" % k] = rel_l2(q[k, :, 1], qo[k, :, 1], floor=1e-6 * c * dpn)
        errs["vdp[%d]" % k] = rel_l2(q[k, :, 2], qo[k, :, 2], floor=1e-6 * c * dpn)
        errs["dpprime[%d]" % k] = rel_l2(qp[k, :, 0], qpo[k, :, 0])
        errs["uprime[%d]" % k] = rel_l2(qp[k, :, 1], qpo[k, :, 1], floor=1e-6 * c * np.sqrt(npoin))
        errs["vprime[%d]" % k] = rel_l2(qp[k, :, 2], qpo[k, :, 2], floor=1e-6 * c * np.sqrt(npoin))
    pbn = np.linalg.norm(qbo[:, 0])
    errs["pb"] = rel_l2(qb[:, 0], qbo[:, 0])
    errs["pbpert"] = rel_l2(qb[:, 1], qbo[:, 1], floor=1e-6 * pbn)
    errs["pbub"] = rel_l2(qb[:, 2], qbo[:, 2], floor=1e-6 * c * pbn)
    errs["pbvb"] = rel_l2(qb[:, 3], qbo[:, 3], floor=1e-6 * c * pbn)
    return errs
