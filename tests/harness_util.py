"""Deck files for tests/c_abi_harness.c (the plain-C caller of the C-ABI) and its build."""
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HARNESS = os.path.join(HERE, "c_abi_harness")


def build_harness(force=False):
    """gcc, linked against the in-tree library (rpath), no CUDA headers"""
    src = os.path.join(HERE, "c_abi_harness.c")
    libdir = os.path.join(ROOT, "h-numo_b200")
    deps = [src, os.path.join(ROOT, "include", "hnumo_b200.h")]
    if force or not os.path.exists(HARNESS) or any(os.path.getmtime(d) > os.path.getmtime(HARNESS) for d in deps):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-Wall", "-o", HARNESS, src, "-L" + libdir, "-lhnumo_b200", "-Wl,-rpath," + libdir])
    return HARNESS


def write_deck(path, deck):
    f64 = lambda a, order="C": np.ascontiguousarray(np.asarray(a, dtype=np.float64)).tobytes() if order == "C" else np.asfortranarray(np.asarray(a, dtype=np.float64)).tobytes(order="F")
    i32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.int32)).tobytes()
    nsr = int(np.sum(deck["num_send_recv"])) if len(deck["nbh_proc"]) else 0
    hdr = [0x484e554d, deck["nelem"], deck["ngl"], deck["nq"], deck["nlayers"], deck["nface"], deck["kstages"], deck["N_btp"], deck["botfr"],
           deck["method_visc"], deck["rank"], deck["nranks"], len(deck["nbh_proc"]), nsr, 0, 0]
    with open(path, "wb") as f:
        f.write(struct.pack("16i", *[int(x) for x in hdr]))
        f.write(struct.pack("7d", deck["dt"], deck["dt_btp"], deck["gravity"], deck["cd_mlswe"], deck["visc_mlswe"], deck["ad_mlswe"], deck.get("max_shear_dz", 0.0)))
        for k in ("psiq", "dpsiq"):
            f.write(f64(deck[k], "F"))          # (ngl,nq) column-major
        f.write(f64(deck["wnq"])); f.write(f64(deck["wgl"])); f.write(f64(deck["dpsi"], "F"))
        f.write(i32(deck["face"])); f.write(f64(deck["elem_metrics"])); f.write(f64(deck["face_geom"]))
        for k in ("pbprime_df", "massinv", "coriolis_df", "tau_wind_df", "zbot_df", "alpha_mlswe"):
            f.write(f64(deck[k]))
        f.write(f64(deck["ssprk_a"], "F")); f.write(f64(deck["ssprk_beta"]))
        if len(deck["nbh_proc"]):
            f.write(i32(deck["nbh_proc"])); f.write(i32(deck["num_send_recv"])); f.write(i32(deck["nbh_send_recv"]))
        f.write(f64(deck["q_df"])); f.write(f64(deck["qb_df"])); f.write(f64(deck["qprime_df"]))


def read_out(path, deck):
    nl, n = deck["nlayers"], deck["npoin"]
    a = np.fromfile(path, dtype=np.float64)
    o = 0
    q = a[o:o + 3 * n * nl].reshape(nl, n, 3); o += 3 * n * nl
    qb = a[o:o + 4 * n].reshape(n, 4); o += 4 * n
    qp = a[o:o + 3 * n * nl].reshape(nl, n, 3); o += 3 * n * nl
    return q, qb, qp, a[o:]
