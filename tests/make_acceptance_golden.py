"""Regenerates tests/golden/acceptance_*.npz: the CPU oracle's fields after 100 baroclinic steps of every BASELINE
configuration that fits a test (north star: "relative L2 error of at most 1e-11 per field after 100 steps"), and next to
them the ROUND-OFF FLOOR of every field: how far two equally valid FP64 evaluations of the same algorithm drift apart
over the same 100 steps,
  floor_metrics = oracle with the reference's numerically differentiated metrics (metrics.F90: ~1e-14 relative round-off that
                  differs from point to point) vs oracle with per-element constant metrics (what the CUDA library is given),
  floor_fma     = oracle compiled with fused multiply-adds (liboracle_fma.so) vs the oracle proper,
both measured as the plain relative L2 difference per field.  tests/test_gpu_acceptance.py runs the CUDA library from the
same initial state and compares with these fields; running the oracle for 100 steps takes minutes to a quarter of an hour
per case, which is why the fields are committed (for the 64x64 deck: every 4th element in x and y, 1/16 of the nodes).

Usage: OMP_NUM_THREADS=8 python tests/make_acceptance_golden.py [case ...]      (CPU only)"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib  # noqa: E402
from hnumo_loader import hnumo_b200 as hn  # noqa: E402

NSTEPS = 100


def cases():
    return {
        # name: (deck parameters, spin-up steps before the 100 compared steps, element stride of the stored subset)
        "bump": (dict(hn.decks.SHIPPED["bump"]), 0, 1),
        "lake": (dict(hn.decks.SHIPPED["lake"]), 0, 1),
        "double_gyre": (dict(hn.decks.SHIPPED["double_gyre"]), 0, 1),
        "double_gyre_spunup": (dict(hn.decks.SHIPPED["double_gyre"]), 200, 1),
        "synth3_64x64": (hn.decks.synthetic_double_gyre(64, 64, nop=4, nlayers=3), 0, 4),
        # general (curved, non-affine) quadrilaterals: the per-point geometry path of the library (the "metrics" floor is zero here:
        # there is only one way to evaluate the metric terms of a curved element)
        "synth3_curved_32x32": (dict(hn.decks.synthetic_double_gyre(32, 32, nop=4, nlayers=3), mesh_warp=0.15), 0, 2),
    }


def subset_nodes(params, stride, npts):
    nelx, nely = params["nelx"], params["nely"]
    el = np.array([e for e in range(nelx * nely) if (e % nelx) % stride == 0 and (e // nelx) % stride == 0], dtype=np.int64)
    return el, (el[:, None] * npts + np.arange(npts)[None, :]).ravel()


def fields(o):
    """the prognostic arrays of ti_rk_bcl in the reference layouts: q_df(3,npoin,nl), qb_df(4,npoin), qprime_df(3,npoin,nl)"""
    nl, n = o.nl, o.npoin
    return o.get("q_df").reshape(nl, n, 3).copy(), o.get("qb_df").reshape(n, 4).copy(), o.get("qprime_df").reshape(nl, n, 3).copy()


FIELD_NAMES = ["q_df.dp", "q_df.udp", "q_df.vdp", "qprime_df.dpp", "qprime_df.up", "qprime_df.vp", "qb_df.pb", "qb_df.pbpert", "qb_df.pbub", "qb_df.pbvb"]


def field_list(q, qb, qp):
    """name -> array: the layer fields keep their layer axis (the relative L2 norm is taken over all layers of a field)"""
    return {"q_df.dp": q[:, :, 0], "q_df.udp": q[:, :, 1], "q_df.vdp": q[:, :, 2], "qprime_df.dpp": qp[:, :, 0], "qprime_df.up": qp[:, :, 1],
            "qprime_df.vp": qp[:, :, 2], "qb_df.pb": qb[:, 0], "qb_df.pbpert": qb[:, 1], "qb_df.pbub": qb[:, 2], "qb_df.pbvb": qb[:, 3]}


def rel_l2(a, b):
    """plain relative L2 error; when the reference field is identically zero (a state at rest) the absolute L2 norm per point"""
    d = np.linalg.norm(np.asarray(a, dtype=np.float64).ravel() - np.asarray(b, dtype=np.float64).ravel())
    n = np.linalg.norm(np.asarray(b, dtype=np.float64).ravel())
    return float(d / n) if n > 0.0 else float(d / np.sqrt(np.asarray(b).size))


def run(name):
    params, spin, stride = cases()[name]
    t0 = time.time()
    base = oracle_lib.Oracle(dict(params, affine_metrics=True))
    if spin:
        assert base.step(spin) == 0
    start = fields(base)
    variants = {"metrics": oracle_lib.Oracle(dict(params, affine_metrics=False)), "fma": oracle_lib.Oracle(dict(params, affine_metrics=True), variant="fma")}
    for o in variants.values():   # every evaluation starts the compared 100 steps from the same state
        o.set("q_df", start[0]); o.set("qb_df", start[1]); o.set("qprime_df", start[2])
    assert base.step(NSTEPS) == 0
    ref = field_list(*fields(base))
    out = {}
    for vn, o in variants.items():
        assert o.step(NSTEPS) == 0
        got = field_list(*fields(o))
        out["floor_" + vn] = np.array([rel_l2(got[f], ref[f]) for f in FIELD_NAMES])
    el, idx = subset_nodes(params, stride, base.npts if hasattr(base, "npts") else base.ngl ** 2)
    q, qb, qp = fields(base)
    out.update(q_df=q[:, idx, :], qb_df=qb[idx, :], qprime_df=qp[:, idx, :], elements=el, stride=np.int64(stride), nsteps=np.int64(NSTEPS),
               spinup=np.int64(spin), field_names=np.array(FIELD_NAMES))
    if spin:
        out.update(start_q_df=start[0], start_qb_df=start[1], start_qprime_df=start[2])
    # invariants of the reference's own documentation, as measured values of the oracle (lake at rest, mass conservation)
    diag, mass = base.diagnostics()
    out["oracle_max_abs_u"] = np.float64(np.abs(diag[:, :, 1:3]).max())
    out["oracle_mass"] = mass
    print("%s: %d(+%d) steps in %.0f s; floors metrics %.1e..%.1e fma %.1e..%.1e" % (name, NSTEPS, spin, time.time() - t0, out["floor_metrics"].min(),
          out["floor_metrics"].max(), out["floor_fma"].min(), out["floor_fma"].max()), flush=True)
    return out


if __name__ == "__main__":
    names = sys.argv[1:] or list(cases())
    for name in names:
        res = run(name)
        path = os.path.join(HERE, "golden", "acceptance_%s.npz" % name)
        np.savez_compressed(path, **res)
        print("wrote", path, os.path.getsize(path), "bytes", flush=True)
