"""CPU tests of the host side: brick decks vs the oracle's independent set-up, C-ABI surface, partition logic."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib
from hnumo_loader import hnumo_b200 as hn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["bump", "lake", "double_gyre"])
def test_deck_matches_oracle_setup(name):
    p = dict(hn.decks.SHIPPED[name])
    if name == "double_gyre":
        p.update(nelx=6, nely=5)
    d = hn.decks.build_deck(p)
    o = oracle_lib.Oracle(p)
    assert d["nface"] == o.nface and d["nelem"] == o.nelem and d["N_btp"] == o.N_btp
    assert abs(d["dt_btp"] - o.dt_btp) == 0.0
    assert np.array_equal(d["face"][:, 4:], o.face()[:, 4:])
    for k in ("pbprime_df", "zbot_df", "coriolis_df", "alpha_mlswe", "ssprk_beta"):
        assert np.array_equal(d[k], o.get(k)), k
    assert np.array_equal(d["tau_wind_df"].ravel(), o.get("tau_wind_df"))
    assert np.array_equal(d["coord"].ravel(), o.get("coord"))
    for k in ("psiq", "dpsiq", "dpsi", "ssprk_a"):
        assert np.array_equal(np.asarray(d[k]).ravel(order="F"), o.get(k)), k
    assert np.array_equal(d["wnq"], o.get("wnq")) and np.array_equal(d["wgl"], o.get("wgl"))
    for k in ("q_df", "qb_df", "qprime_df"):
        assert np.array_equal(d[k].ravel(), o.get(k)), k
    # analytic brick geometry vs the numerically differentiated metrics of the reference (round-off only)
    assert np.abs(d["massinv"] / o.get("massinv") - 1).max() < 1e-12
    nq2 = o.nq ** 2
    assert np.abs(d["elem_metrics"][:, 0] / o.get("ksiq_x")[::nq2] - 1).max() < 1e-12
    assert np.abs(d["elem_metrics"][:, 3] / o.get("etaq_y")[::nq2] - 1).max() < 1e-12
    assert np.abs(o.get("ksiq_y")).max() < 1e-14 * np.abs(o.get("ksiq_x")).max() * 1e3
    nv = o.get("normal_vector_q").reshape(o.nface, o.nq, 2)
    assert np.abs(nv[:, 0, :] - d["face_geom"][:, :2]).max() < 1e-12


def test_synthetic_deck_matches_oracle():
    p = hn.decks.synthetic_double_gyre(6, 6, nop=4, nlayers=3)
    d = hn.decks.build_deck(p)
    o = oracle_lib.Oracle(p)
    for k in ("q_df", "qb_df", "qprime_df"):
        assert np.array_equal(d[k].ravel(), o.get(k)), k
    assert np.array_equal(d["alpha_mlswe"], o.get("alpha_mlswe"))
    # CFL-rescaled time step keeps the shipped barotropic Courant number
    c = np.sqrt(9.806 * 9928.0)
    xg = hn.decks.lgl(5)[0]
    dmin = (xg[1] - xg[0]) / 2 * 2e6 / 6
    assert abs(c * d["dt_btp"] / dmin - c * 25.0 / ((xg[1] - xg[0]) / 2 * 2e6 / 25)) < 1e-9


def test_library_exports_every_declared_symbol(hn_lib):
    hdr = open(os.path.join(ROOT, "include", "hnumo_b200.h")).read()
    declared = set(re.findall(r"\b(hnumo_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no prototypes found"
    for sym in declared:
        assert hasattr(hn_lib, sym), sym
    assert declared == set(hn.EXPORTS)


def test_fortran_interface_matches_header(hn_lib):
    """h-numo_b200/fortran/hnumo_b200_iface.F90 (the reference-side ISO_C_BINDING layer) lists the
    fields of hnumo_desc_t in the order of the C struct / ctypes mirror, and binds only exported symbols."""
    src = open(os.path.join(ROOT, "h-numo_b200", "fortran", "hnumo_b200_iface.F90")).read()
    body = src.split("type, bind(C) :: hnumo_desc_t")[1].split("end type")[0]
    fields = []
    for line in body.splitlines():
        line = line.split("!")[0]
        if "::" in line:
            fields += [f.strip() for f in line.split("::")[1].split(",") if f.strip()]
    assert fields == [f[0] for f in hn.Desc._fields_]
    hdr = open(os.path.join(ROOT, "include", "hnumo_b200.h")).read()
    struct = hdr.split("typedef struct hnumo_desc {")[1].split("} hnumo_desc_t;")[0]
    struct = re.sub(r"/\*.*?\*/", "", struct, flags=re.S)
    cfields = []
    for stmt in struct.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        names = stmt.replace("const", "").replace("*", " ").split()[1:]
        cfields += [n.strip(",") for n in " ".join(names).split(",")]
    cfields = [c.strip() for c in cfields if c.strip()]
    assert cfields == fields
    for name in re.findall(r'bind\(C, name="(hnumo_[a-z_0-9]+)"\)', src):
        assert hasattr(hn_lib, name), name
    shim = open(os.path.join(ROOT, "h-numo_b200", "fortran", "ti_rk_bcl_b200.F90")).read()
    assert re.search(r"subroutine ti_rk_bcl\(q_df, qb_df, qprime_df\)", shim)


def test_no_cpu_fallback_without_gpu(hn_lib):
    """On a box without a CUDA device hnumo_init must fail loudly (no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    d = hn.decks.build_deck(dict(hn.decks.SHIPPED["bump"], nelx=2, nely=2))
    with pytest.raises(hn.HnumoError):
        hn.Solver(d)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "h-numo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".F90")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in txt and "liboracle" not in txt and "hnumo_oracle" not in txt, f


@pytest.mark.parametrize("nranks", [2, 4])
def test_partition_is_consistent(nranks):
    """row-block partition: element sets tile the brick, processor faces pair up in the same order on both sides"""
    p = dict(hn.decks.SHIPPED["double_gyre"], nelx=5, nely=8)
    decks = [hn.decks.build_deck(p, r, nranks) for r in range(nranks)]
    glob = hn.decks.build_deck(p)
    allg = np.concatenate([d["elem_global"] for d in decks])
    assert sorted(allg.tolist()) == list(range(glob["nelem"]))
    assert sum(d["nface"] for d in decks) == glob["nface"] + (nranks - 1) * p["nelx"]
    for r, d in enumerate(decks):
        off = 0
        for nb, cnt in zip(d["nbh_proc"], d["num_send_recv"]):
            other = decks[nb - 1]
            j = list(other["nbh_proc"]).index(r + 1)
            ooff = int(np.sum(other["num_send_recv"][:j]))
            assert other["num_send_recv"][j] == cnt
            for i in range(cnt):
                f = d["face"][d["nbh_send_recv"][off + i] - 1]
                g = other["face"][other["nbh_send_recv"][ooff + i] - 1]
                assert f[7] == 0 and g[7] == 0
                eg = d["elem_global"][f[6] - 1]
                og = other["elem_global"][g[6] - 1]
                assert abs(int(eg) - int(og)) == p["nelx"]          # vertical neighbours in the global brick
                assert {int(f[4]), int(g[4])} == {3, 4}
            off += cnt
        # state restricted to the partition equals the global state
        npts = d["npts"]
        idx = (d["elem_global"][:, None] * npts + np.arange(npts)[None, :]).ravel()
        assert np.array_equal(d["qb_df"], glob["qb_df"][idx])


def test_gloo_halo_ordering_world2():
    """world_size-2 gloo run of the host-side exchange logic: each rank sends the face-node coordinates of its
    processor faces in nbh_send_recv order; what arrives must be the coordinates of the same physical nodes."""
    script = os.path.join(ROOT, "tests", "gloo_halo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", script]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("HALO_OK") == 2, r.stdout + r.stderr
