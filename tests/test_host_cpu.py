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


@pytest.mark.parametrize("nop", [4, 3])
def test_general_quadrilateral_deck_matches_oracle(nop):
    """mesh_warp: curved, non-affine elements.  The per-point geometry decks.py hands to the library (nodal and quadrature-point
    metric terms, face normals and Jacobians at the face nodes and face quadrature points) is the one the oracle derives with the
    reference's metrics / normals code (metrics.F90, metrics_quad.F90, create_normals(_quad).F90); the elements really are
    non-affine (the metric terms vary by tens of percent inside an element)."""
    p = dict(hn.decks.synthetic_double_gyre(6, 4, nop=nop, nlayers=3), mesh_warp=0.15)
    d = hn.decks.build_deck(p)
    o = oracle_lib.Oracle(p)
    assert d["elem_metrics"] is None and d["face_geom"] is None
    c = o.get("coord").reshape(-1, 2)
    assert np.abs(d["coord"] - c).max() <= 1e-9 * 2.0e6
    assert np.abs(d["coord"] - hn.decks.build_deck(dict(p, mesh_warp=0.0))["coord"]).max() > 0.1 * 2.0e6 / 6 * 0.15
    ref_q = np.stack([o.get(k) for k in ("ksiq_x", "ksiq_y", "etaq_x", "etaq_y", "jacq")], axis=1)
    ref_n = np.stack([o.get(k) for k in ("ksi_x", "ksi_y", "eta_x", "eta_y", "jac")], axis=1)
    for a, b in ((d["point_metrics_q"], ref_q), (d["point_metrics"], ref_n)):
        assert a.shape == b.shape
        assert np.abs(a - b).max(axis=0).max() <= 1e-11 * np.abs(b).max(axis=0).min() or np.all(np.abs(a - b).max(axis=0) <= 1e-11 * np.abs(b).max(axis=0))
    nq2 = o.nq ** 2
    kx = ref_q[:, 0].reshape(-1, nq2)
    assert (kx.max(axis=1) / kx.min(axis=1)).max() > 1.2          # non-affine: the metric terms vary inside an element
    fq = np.concatenate([o.get("normal_vector_q").reshape(o.nface, o.nq, 2), o.get("jac_faceq").reshape(o.nface, o.nq, 1)], axis=2)
    fn = np.concatenate([o.get("normal_vector").reshape(o.nface, o.ngl, 2), o.get("jac_face").reshape(o.nface, o.ngl, 1)], axis=2)
    assert np.abs(d["face_geom_q"][..., :2] - fq[..., :2]).max() < 1e-11 and np.abs(d["face_geom_q"][..., 2] / fq[..., 2] - 1).max() < 1e-11
    assert np.abs(d["face_geom_n"][..., :2] - fn[..., :2]).max() < 1e-11 and np.abs(d["face_geom_n"][..., 2] / fn[..., 2] - 1).max() < 1e-11
    assert np.abs(d["massinv"] / o.get("massinv") - 1).max() < 1e-11
    for k in ("q_df", "qb_df", "qprime_df"):
        assert np.allclose(d[k].ravel(), o.get(k), rtol=1e-12, atol=0.0), k
    # the oracle itself on the warped mesh: discrete mass conservation of the barotropic RHS and of a whole step
    assert o.step(2) == 0
    q = o.get("q_df").reshape(3, -1, 3)
    mass0 = (d["q_df"][:, :, 0] / o.get("massinv")[None, :]).sum(axis=1)
    mass1 = (q[:, :, 0] / o.get("massinv")[None, :]).sum(axis=1)
    assert np.abs(mass1 / mass0 - 1).max() < 1e-13


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


def test_gloo_halo_ordering_world4_morton():
    """the same on 4 ranks of a Morton-curve partition (ranks with two and three neighbours, several messages per exchange)"""
    script = os.path.join(ROOT, "tests", "gloo_halo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29543", HALO_NELX="6", HALO_NELY="8", HALO_PARTITION="morton")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
           "--master-port", "29543", script]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("HALO_OK") == 4, r.stdout + r.stderr


def test_general_partitions_cover_the_brick_and_pair_up():
    """element_owner: every element has exactly one owner; neighbour lists are symmetric with equal face counts"""
    p = dict(hn.decks.SHIPPED["double_gyre"], nelx=6, nely=8)
    for kind, n in (("rows", 4), ("blocks:2x2", 4), ("blocks:3x2", 6), ("morton", 3), ("morton", 4), ("morton", 8)):
        pp = dict(p, partition=kind)
        ds = [hn.decks.build_deck(pp, r, n) for r in range(n)]
        assert sorted(np.concatenate([d["elem_global"] for d in ds]).tolist()) == list(range(48)), kind
        for r, d in enumerate(ds):
            assert d["num_send_recv"].sum() == (d["face"][:, 7] == 0).sum(), kind
            for nb, cnt in zip(d["nbh_proc"], d["num_send_recv"]):
                o = ds[nb - 1]
                j = o["nbh_proc"].tolist().index(r + 1)
                assert o["num_send_recv"][j] == cnt, (kind, r, nb)
            # left element of an interior face has the lower local number (p4est.c:1693)
            inter = d["face"][:, 7] > 0
            assert np.all(d["face"][inter, 6] < d["face"][inter, 7]), kind


def test_snapshot_roundtrip_and_restart(tmp_path):
    """text snapshots in the reference's format (diagnostics.F90:73-91: one d23.16 value per line) and the
    restart conversion of mod_restart.F90:15-66.  Host-only entry points of the C-ABI library (no GPU)."""
    import numpy as np

    p = hn.decks.synthetic_double_gyre(3, 2, nop=4, nlayers=3)
    deck = hn.decks.build_deck(p)
    rng = np.random.default_rng(7)
    q = deck["q_df"].copy(); qb = deck["qb_df"].copy()
    # a state in motion, consistent like the model keeps it: qb momentum = sum of the layer momenta, pb = sum of dp
    for k in range(3):
        q[k, :, 1] = q[k, :, 0] * rng.normal(0.0, 0.3, deck["npoin"])
        q[k, :, 2] = q[k, :, 0] * rng.normal(0.0, 0.3, deck["npoin"])
    qb[:, 2] = q[:, :, 1].sum(axis=0); qb[:, 3] = q[:, :, 2].sum(axis=0)
    path = tmp_path / "mlswe0007"
    hn.snapshot_write(path, deck, q, qb)
    lines = open(path).read().split("\n")
    npn, nl = deck["npoin"], 3
    assert lines[0] == "%4d" % nl and lines[1] == "%10d" % npn
    # Fortran d23.16: 23 characters, "0.dddddddddddddddd" mantissa with 16 digits, D exponent
    import re
    assert all(len(x) == 23 and re.fullmatch(r" *-?0\.\d{16}D[+-]\d\d", x) for x in lines[2:200])
    assert len([x for x in lines if x]) == 4 + 2 * npn + 3 * npn + 4 * nl * npn + npn
    # (16 significant digits: the reference's format does not round-trip the last bit of a double)
    assert abs(float(lines[2].replace("D", "E")) - deck["dt"]) <= 1e-15 * deck["dt"]
    info = hn.snapshot_info(path)
    assert info["nlayers"] == nl and info["npoin"] == npn
    assert abs(info["dt"] - deck["dt"]) <= 1e-15 * deck["dt"] and abs(info["dt_btp"] - deck["dt_btp"]) <= 1e-15 * deck["dt_btp"]
    # layer thickness h of layer 1 at the first point, as diagnostics.F90 defines it
    h0 = float(lines[4 + 2 * npn + 3 * npn].replace("D", "E"))
    assert abs(h0 - deck["alpha_mlswe"][0] / deck["gravity"] * q[0, 0, 0]) <= 1e-15 * abs(h0)
    q2, qb2, qp2, coord = hn.snapshot_read_restart(path, deck)
    assert np.allclose(coord, deck["coord"], rtol=1e-15, atol=1e-9)
    assert np.allclose(qb2[:, [0, 2, 3]], qb[:, [0, 2, 3]], rtol=2e-15, atol=1e-300)
    assert np.allclose(qb2[:, 1], qb2[:, 0] - deck["pbprime_df"], rtol=0, atol=0)
    assert np.allclose(q2, q, rtol=4e-15, atol=1e-12)
    # primes as restart_mlswe rebuilds them
    ope = q2[:, :, 0].sum(axis=0) / deck["pbprime_df"]
    for k in range(nl):
        assert np.allclose(qp2[k, :, 0], q2[k, :, 0] / ope, rtol=1e-15)
        assert np.allclose(qp2[k, :, 1], q2[k, :, 1] / q2[k, :, 0] - qb2[:, 2] / qb2[:, 0], rtol=1e-13, atol=1e-15)
    # error behaviour: size mismatch and missing file
    other = hn.decks.build_deck(hn.decks.synthetic_double_gyre(2, 2, nop=4, nlayers=3))
    with pytest.raises(hn.HnumoError, match="-8"):
        hn.snapshot_read_restart(path, other)
    with pytest.raises(hn.HnumoError, match="-6"):
        hn.snapshot_info(tmp_path / "missing")


def test_netcdf_snapshot_layout_and_restart(tmp_path):
    """SURVEY 8(f) rank 3, NetCDF layout: hnumo_snapshot_write_nc writes the file of src/diagnostics_nc.F90:98-165 (classic format with
    64-bit offsets, no NetCDF library); an independent reader (scipy.io.netcdf_file) must see the reference's dimensions, variable
    names, attributes and values, and hnumo_snapshot_read_nc_restart must rebuild the state like restart_mlswe does."""
    import numpy as np
    from scipy.io import netcdf_file

    p = hn.decks.synthetic_double_gyre(3, 2, nop=4, nlayers=3)
    deck = hn.decks.build_deck(p)
    rng = np.random.default_rng(11)
    q = deck["q_df"].copy(); qb = deck["qb_df"].copy()
    for k in range(3):
        q[k, :, 0] *= 1.0 + 1e-3 * rng.normal(0.0, 1.0, deck["npoin"])
        q[k, :, 1] = q[k, :, 0] * rng.normal(0.0, 0.3, deck["npoin"])
        q[k, :, 2] = q[k, :, 0] * rng.normal(0.0, 0.3, deck["npoin"])
    qb[:, 0] = q[:, :, 0].sum(axis=0); qb[:, 1] = qb[:, 0] - deck["pbprime_df"]
    qb[:, 2] = q[:, :, 1].sum(axis=0); qb[:, 3] = q[:, :, 2].sum(axis=0)
    path = tmp_path / "mlswe0003.nc"
    hn.snapshot_write_nc(path, deck, q, qb)
    raw = open(path, "rb").read()
    assert raw[:4] == b"CDF\x02"                                   # nf90_64bit_offset
    npn, nl, g, al = deck["npoin"], 3, deck["gravity"], deck["alpha_mlswe"]
    f = netcdf_file(str(path), "r", mmap=False)
    try:
        assert list(f.dimensions.keys()) == ["time", "npoin", "nlayers", "zi"]
        assert f.dimensions["time"] is None and f.dimensions["npoin"] == npn and f.dimensions["nlayers"] == nl and f.dimensions["zi"] == nl + 1
        assert f.filename if hasattr(f, "filename") else True
        assert f._attributes["filename"] == b"mlswe0003.nc" and f._attributes["zi"] == b"Number of interfaces"
        assert list(f.variables.keys()) == ["dt", "dt_btp", "x", "y", "pb", "pbub", "pbvb", "h", "u", "v", "eta"]
        v = f.variables
        assert v["dt"].dimensions == ("time",) and v["h"].dimensions == ("nlayers", "npoin") and v["eta"].dimensions == ("zi", "npoin")
        assert v["dt"].units == b"seconds" and v["h"].name == b"Layer thickness" and v["x"].axis == b"X"
        assert v["pb"].units == "N/m\u00b2".encode() and v["pbub"].units == "kg\u00b7m/s".encode()
        assert v["dt"][:].shape == (1,) and v["dt"][0] == deck["dt"] and v["dt_btp"][0] == deck["dt_btp"]
        assert np.array_equal(v["x"][:], deck["coord"][:, 0]) and np.array_equal(v["y"][:], deck["coord"][:, 1])
        assert np.array_equal(v["pb"][:], qb[:, 0]) and np.array_equal(v["pbub"][:], qb[:, 2]) and np.array_equal(v["pbvb"][:], qb[:, 3])
        h = np.stack([(al[k] / g) * q[k, :, 0] for k in range(nl)])
        assert np.array_equal(v["h"][:], h)
        assert np.array_equal(v["u"][:], q[:, :, 1] / q[:, :, 0]) and np.array_equal(v["v"][:], q[:, :, 2] / q[:, :, 0])
        eta = v["eta"][:]
        assert np.array_equal(eta[nl], deck["zbot_df"])
        z = deck["zbot_df"].copy()
        for k in range(nl - 1, 0, -1):                               # mslwe_elevation, diagnostics_nc.F90:59-63
            z = z + h[k]
            assert np.array_equal(eta[k], z)
        assert np.allclose(eta[0], qb[:, 0] / deck["pbprime_df"] - 1.0, rtol=0, atol=4e-16)
    finally:
        f.close()
    q2, qb2, qp2, coord = hn.snapshot_read_nc_restart(path, deck)
    assert np.array_equal(coord, deck["coord"]) and np.array_equal(qb2[:, [0, 2, 3]], qb[:, [0, 2, 3]])
    assert np.array_equal(qb2[:, 1], qb2[:, 0] - deck["pbprime_df"])
    assert np.allclose(q2, q, rtol=4e-16, atol=0)                    # binary doubles: only the h <-> dp scaling rounds
    ope = q2[:, :, 0].sum(axis=0) / deck["pbprime_df"]
    for k in range(nl):
        assert np.allclose(qp2[k, :, 0], q2[k, :, 0] / ope, rtol=1e-15)
        assert np.allclose(qp2[k, :, 2], q2[k, :, 2] / q2[k, :, 0] - qb2[:, 3] / qb2[:, 0], rtol=1e-13, atol=1e-15)
    # the text and the NetCDF snapshots hold the same fields
    tpath = tmp_path / "mlswe0003"
    hn.snapshot_write(tpath, deck, q, qb)
    q3, qb3, _, _ = hn.snapshot_read_restart(tpath, deck)
    assert np.allclose(q3, q2, rtol=4e-15, atol=1e-12) and np.allclose(qb3, qb2, rtol=4e-15, atol=1e-6)
    # error behaviour
    other = hn.decks.build_deck(hn.decks.synthetic_double_gyre(2, 2, nop=4, nlayers=3))
    with pytest.raises(hn.HnumoError, match="-8"):
        hn.snapshot_read_nc_restart(path, other)
    with pytest.raises(hn.HnumoError, match="-7"):
        hn.snapshot_read_nc_restart(tpath, deck)
    with pytest.raises(hn.HnumoError, match="-6"):
        hn.snapshot_read_nc_restart(tmp_path / "missing.nc", deck)


def test_fortran_d23_16_formatting(tmp_path):
    """corner cases of the d23.16 writer / reader: zero, negative, tiny, huge (three-digit exponents drop the letter)"""
    import numpy as np

    deck = dict(hn.decks.build_deck(hn.decks.synthetic_double_gyre(1, 1, nop=2, nlayers=2)))
    vals = [0.0, -1.5, 1.0, 9.999999999999999e22, 1e-5, -2.2250738585072014e-308, 1.6e308, 123456.789e100, 0.1]
    npn = deck["npoin"]
    coord = np.zeros((npn, 2)); coord.flat[:len(vals)] = vals[:2 * npn]
    deck["coord"] = coord
    path = tmp_path / "mlswe0000"
    hn.snapshot_write(path, deck, deck["q_df"], deck["qb_df"])
    lines = open(path).read().split("\n")
    got = lines[4:4 + 2 * npn]
    assert got[0] == " 0.0000000000000000D+00" and got[1] == "-0.1500000000000000D+01" and got[2] == " 0.1000000000000000D+01"
    assert got[3] == " 0.9999999999999999D+23" and got[4] == " 0.1000000000000000D-04"
    assert got[5] == "-0.2225073858507201-307" and got[6] == " 0.1600000000000000+309"
    _, _, _, c2 = hn.snapshot_read_restart(path, deck)
    for a, b in zip(c2.flat[:2 * npn], coord.flat[:2 * npn]):
        assert a == b or abs(a - b) <= 1e-15 * abs(b)


@pytest.mark.parametrize("change,code,text", [
    (dict(ad_mlswe=1.0e-3), -2, "max_shear_dz"),                   # vertical shear stress needs its length scale (division by max_shear_dz)
    (dict(nlayers=0), -2, "unsupported sizes"),
    (dict(kstages=6), -2, "unsupported sizes"),
    (dict(ngl=12), -2, "unsupported sizes"),
    (dict(elem_metrics=None), -2, "elem_metrics"),                 # affine mesh without its per-element geometry
    (dict(point_metrics_q=np.zeros((4 * 81, 5))), -2, "general quadrilaterals need"),   # per-point geometry given only in part
])
def test_init_rejects_unsupported_configurations(hn_lib, change, code, text):
    """argument validation happens before any device work: same answers with and without a GPU"""
    d = dict(hn.decks.build_deck(dict(hn.decks.SHIPPED["bump"], nelx=2, nely=2)))
    d.update(change)
    with pytest.raises(hn.HnumoError) as e:
        hn.Solver(d)
    assert "(%d)" % code in str(e.value) and text in str(e.value), str(e.value)


def test_null_handle_and_bad_buffers_are_errors(hn_lib):
    L = hn.load_library()
    assert L.hnumo_step(None, 1) == -2 and L.hnumo_finalize(None) == -2
    assert L.hnumo_diagnostics(None, None, 0) == -2
    assert L.hnumo_snapshot_write(None, 1, 1, 0.0, 0.0, None, None, None, None, None, 9.806) == -2


def test_snapshot_is_readable_by_the_reference_loader(tmp_path):
    """The reference ships a reader of its snapshot format: Examples/bump/load_data_numo.m (flat list of numbers: nk, npoin, dt,
    dt_btp, coord(2,npoin), pb, u*pb, v*pb, dp(npoin,nk), u, v, z(npoin,nk+1)).  Restated in numpy and applied to a file
    written by hnumo_snapshot_write, it must return the state that was written."""
    deck = hn.decks.build_deck(dict(hn.decks.SHIPPED["bump"], nelx=3, nely=2))
    q = deck["q_df"].copy(); qb = deck["qb_df"].copy()
    q[:, :, 1] = 0.01 * q[:, :, 0]; q[:, :, 2] = -0.02 * q[:, :, 0]
    qb[:, 2] = q[:, :, 1].sum(axis=0); qb[:, 3] = q[:, :, 2].sum(axis=0)
    path = tmp_path / "mlswe0001"
    hn.snapshot_write(path, deck, q, qb)
    temp = np.array([float(t.replace("D", "E")) for t in open(path).read().split()])   # MATLAB: load(file, '-ascii')
    count = 0
    nk = int(temp[count]); count += 1
    npoin = int(temp[count]); count += 1
    dt = temp[count]; count += 1
    dt_btp = temp[count]; count += 1
    coord = temp[count:count + 2 * npoin].reshape(npoin, 2); count += 2 * npoin          # reshape([2,npoin]) column-major
    pb = temp[count:count + npoin]; count += npoin
    upb = temp[count:count + npoin]; count += npoin
    vpb = temp[count:count + npoin]; count += npoin
    dp = temp[count:count + npoin * nk].reshape(nk, npoin); count += npoin * nk           # reshape([npoin,nk]) column-major
    u = temp[count:count + npoin * nk].reshape(nk, npoin); count += npoin * nk
    v = temp[count:count + npoin * nk].reshape(nk, npoin); count += npoin * nk
    z = temp[count:count + npoin * (nk + 1)].reshape(nk + 1, npoin); count += npoin * (nk + 1)
    assert count == temp.size and nk == deck["nlayers"] and npoin == deck["npoin"]
    assert abs(dt - deck["dt"]) <= 1e-15 * dt and abs(dt_btp - deck["dt_btp"]) <= 1e-15 * dt_btp
    assert np.allclose(coord, deck["coord"], rtol=1e-15, atol=1e-12)
    assert np.allclose(pb, qb[:, 0], rtol=2e-16 * 8) and np.allclose(upb, qb[:, 2], rtol=2e-15, atol=1e-300) and np.allclose(vpb, qb[:, 3], rtol=2e-15, atol=1e-300)
    g, al = deck["gravity"], deck["alpha_mlswe"]
    for k in range(nk):
        assert np.allclose(dp[k], al[k] / g * q[k, :, 0], rtol=2e-15)      # "dp" of the loader is the thickness h in metres
        assert np.allclose(u[k], 0.01, rtol=1e-14) and np.allclose(v[k], -0.02, rtol=1e-14)
    assert np.allclose(z[nk], deck["zbot_df"], rtol=1e-15)
    assert np.allclose(z[0], deck["zbot_df"] + sum(al[k] / g * q[k, :, 0] for k in range(nk)), rtol=1e-14, atol=1e-12)


def test_c_harness_builds_and_fails_loudly_without_a_gpu(hn, hn_lib, tmp_path):
    """tests/c_abi_harness.c is a plain-C caller of the C-ABI (no Python, no CUDA headers).  It links against the library and,
    on a box without a GPU, stops at hnumo_init with the library's own message: there is no CPU fallback behind the boundary."""
    import subprocess
    import torch
    import harness_util
    exe = harness_util.build_harness()
    deck = hn.decks.build_deck(dict(hn.decks.SHIPPED["bump"], nelx=3, nely=3))
    deck_file, out_file = str(tmp_path / "deck.bin"), str(tmp_path / "out.bin")
    harness_util.write_deck(deck_file, deck)
    r = subprocess.run([exe, deck_file, out_file, "1"], capture_output=True, text=True, timeout=300)
    if torch.cuda.is_available():
        assert r.returncode == 0, (r.stdout, r.stderr)
    else:
        assert r.returncode == 3 and "no CUDA device" in r.stderr, (r.returncode, r.stderr)
