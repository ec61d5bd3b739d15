"""hnumo_b200 -- Python host side of libhnumo_b200.so (ctypes over the C-ABI of include/hnumo_b200.h).

The product is the CUDA library; this package only (1) builds it in-tree for sm_100a, (2) binds the C entry points
the Fortran driver would bind through ISO_C_BINDING and (3) provides brick decks (decks.py) so
tests and bench.py can drive the hot path without the Fortran program.  There is no CPU fallback: creating a Solver
without the CUDA library or without a GPU raises.

The directory name contains a hyphen, so import it through ``hnumo_loader`` at the repo root
(``from hnumo_loader import hnumo_b200``) or ``__graft_entry__``.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

from . import decks  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("HNUMO_LIB_PATH", os.path.join(_HERE, "libhnumo_b200.so"))  # override: instrumented debug builds
_LIBS = {}

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared", "-diag-suppress", "177,550"]


EXACT_LIB_PATH = os.path.join(_HERE, "libhnumo_b200_exact.so")


def build_library(force=False, verbose=False, exact=False):
    """nvcc cross-compile of the single translation unit into h-numo_b200/libhnumo_b200.so (in-tree).

    exact=True builds the measurement aid libhnumo_b200_exact.so instead: the same sources with the two deliberate arithmetic
    deviations of the stage kernel switched off (-DHN_EXACT_DIV: IEEE division instead of rcp.approx + 2 Newton steps;
    -DHN_NO_GZ_FLUSH: grad(z_bot) kept as computed).  tests/test_gpu_acceptance.py loads it next to the default build."""
    src = os.path.join(_HERE, "csrc", "hnumo_b200.cu")
    out = EXACT_LIB_PATH if exact else LIB_PATH
    deps = [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc"))]
    deps.append(os.path.join(_ROOT, "include", "hnumo_b200.h"))
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-DHN_EXACT_DIV", "-DHN_NO_GZ_FLUSH"] if exact else []) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, src, "-lpthread", "-ldl"]
    env = dict(os.environ)
    env.pop("CXX", None); env.pop("CC", None)
    r = subprocess.run(cmd, env=env, stderr=None if verbose else subprocess.PIPE, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stderr or "")
        raise subprocess.CalledProcessError(r.returncode, cmd)
    return out


class Desc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("nelem", C.c_int32), ("ngl", C.c_int32), ("nq", C.c_int32), ("nlayers", C.c_int32), ("nface", C.c_int32),
        ("kstages", C.c_int32), ("N_btp", C.c_int32),
        ("dt", C.c_double), ("dt_btp", C.c_double),
        ("botfr", C.c_int32), ("method_visc", C.c_int32),
        ("gravity", C.c_double), ("cd_mlswe", C.c_double), ("visc_mlswe", C.c_double), ("ad_mlswe", C.c_double),
        ("psiq", C.c_void_p), ("dpsiq", C.c_void_p), ("wnq", C.c_void_p), ("wgl", C.c_void_p), ("dpsi", C.c_void_p),
        ("face", C.c_void_p), ("elem_metrics", C.c_void_p), ("face_geom", C.c_void_p),
        ("pbprime_df", C.c_void_p), ("massinv", C.c_void_p), ("coriolis_df", C.c_void_p), ("tau_wind_df", C.c_void_p),
        ("zbot_df", C.c_void_p), ("alpha_mlswe", C.c_void_p), ("ssprk_a", C.c_void_p), ("ssprk_beta", C.c_void_p),
        ("rank", C.c_int32), ("nranks", C.c_int32), ("num_nbh", C.c_int32),
        ("nbh_proc", C.c_void_p), ("num_send_recv", C.c_void_p), ("nbh_send_recv", C.c_void_p),
        ("device", C.c_int32), ("stage_kernel_variant", C.c_int32),
        ("max_shear_dz", C.c_double),
        ("point_metrics_q", C.c_void_p), ("point_metrics", C.c_void_p), ("face_geom_q", C.c_void_p), ("face_geom_n", C.c_void_p),
        ("coord", C.c_void_p),
    ]


EXPORTS = ["hnumo_init", "hnumo_device_count", "hnumo_finalize", "hnumo_last_error", "hnumo_upload_state", "hnumo_download_state",
           "hnumo_step", "hnumo_ti_rk_bcl", "hnumo_btp_bcl_coeffs", "hnumo_btp_substeps", "hnumo_rhs_btp", "hnumo_layer_mass_rhs", "hnumo_layer_momentum_rhs", "hnumo_layer_shear_stress", "hnumo_halo_exchange",
           "hnumo_get_array", "hnumo_diagnostics", "hnumo_snapshot_write", "hnumo_snapshot_info",
           "hnumo_snapshot_read_restart", "hnumo_snapshot_write_nc", "hnumo_snapshot_read_nc_restart", "hnumo_comm_get_unique_id", "hnumo_comm_init", "hnumo_timing", "hnumo_set_option"]


def load_library(path=None):
    """ctypes handle of the library (default: the in-tree build; `path`: another build of the same sources, e.g. the
    exact-division build profiles/build_exact.sh makes for the parity report)"""
    path = path or LIB_PATH
    if path not in _LIBS:
        if not os.path.exists(path):
            raise RuntimeError("%s is not built (run __graft_entry__.build()); there is no CPU fallback" % os.path.basename(path))
        L = C.CDLL(path)
        L.hnumo_last_error.restype = C.c_char_p
        L.hnumo_init.argtypes = [C.POINTER(Desc), C.POINTER(C.c_void_p)]
        L.hnumo_finalize.argtypes = [C.c_void_p]
        L.hnumo_upload_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hnumo_download_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hnumo_step.argtypes = [C.c_void_p, C.c_int32]
        L.hnumo_ti_rk_bcl.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hnumo_btp_bcl_coeffs.argtypes = [C.c_void_p]
        L.hnumo_btp_substeps.argtypes = [C.c_void_p]
        L.hnumo_rhs_btp.argtypes = [C.c_void_p, C.c_void_p]
        L.hnumo_layer_mass_rhs.argtypes = [C.c_void_p, C.c_void_p]
        L.hnumo_layer_momentum_rhs.argtypes = [C.c_void_p, C.c_void_p]
        L.hnumo_layer_shear_stress.argtypes = [C.c_void_p, C.c_void_p]
        L.hnumo_halo_exchange.restype = C.c_int64
        L.hnumo_halo_exchange.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.hnumo_get_array.restype = C.c_int64
        L.hnumo_get_array.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
        L.hnumo_snapshot_write.argtypes = [C.c_char_p, C.c_int32, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_double]
        L.hnumo_snapshot_info.argtypes = [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.hnumo_snapshot_read_restart.argtypes = [C.c_char_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                                  C.c_void_p, C.c_void_p]
        L.hnumo_snapshot_write_nc.argtypes = [C.c_char_p, C.c_int32, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        L.hnumo_snapshot_read_nc_restart.argtypes = [C.c_char_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                                     C.c_void_p, C.c_void_p]
        L.hnumo_diagnostics.restype = C.c_int64
        L.hnumo_diagnostics.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.hnumo_comm_get_unique_id.argtypes = [C.c_void_p]
        L.hnumo_comm_init.argtypes = [C.c_void_p, C.c_void_p]
        L.hnumo_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.hnumo_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        _LIBS[path] = L
    return _LIBS[path]


class HnumoError(RuntimeError):
    pass


class Solver:
    """Device-resident hot path for one partition; mirrors the reference call sequence around ti_rk_bcl."""

    def __init__(self, deck, device=0, variant=0, lib_path=None):
        self.L = load_library(lib_path)
        self.deck = deck
        self._keep = []
        d = Desc()
        d.abi_version = 3
        for k in ("nelem", "ngl", "nq", "nlayers", "nface", "kstages", "N_btp", "botfr", "method_visc", "rank", "nranks"):
            setattr(d, k, int(deck[k]))
        for k in ("dt", "dt_btp", "gravity", "cd_mlswe", "visc_mlswe", "ad_mlswe"):
            setattr(d, k, float(deck[k]))

        def ptr(a, dtype):
            a = np.ascontiguousarray(a, dtype=dtype) if not (isinstance(a, np.ndarray) and a.flags.f_contiguous and a.dtype == dtype and a.ndim == 2) else a
            self._keep.append(a)
            return a.ctypes.data

        # 2-D operator tables are column-major (Fortran); np arrays given as F-ordered keep that layout
        for k in ("psiq", "dpsiq", "dpsi", "ssprk_a"):
            a = np.asfortranarray(deck[k], dtype=np.float64)
            self._keep.append(a)
            setattr(d, k, a.ctypes.data)
        for k in ("wnq", "wgl", "elem_metrics", "face_geom", "pbprime_df", "massinv", "coriolis_df", "tau_wind_df", "zbot_df",
                  "alpha_mlswe", "ssprk_beta"):
            setattr(d, k, ptr(deck[k], np.float64) if deck.get(k) is not None else None)   # elem_metrics / face_geom: None for general quadrilaterals
        d.face = ptr(deck["face"], np.int32)
        d.num_nbh = len(deck["nbh_proc"])
        d.nbh_proc = ptr(deck["nbh_proc"], np.int32) if d.num_nbh else None
        d.num_send_recv = ptr(deck["num_send_recv"], np.int32) if d.num_nbh else None
        d.nbh_send_recv = ptr(deck["nbh_send_recv"], np.int32) if d.num_nbh else None
        d.device = device
        d.stage_kernel_variant = variant
        d.max_shear_dz = float(deck.get("max_shear_dz", 0.0))
        if deck.get("point_metrics_q") is not None:   # general quadrilaterals: geometry per point ((npoin_q,5), (npoin,5), (nface,nq,3), (nface,ngl,3), (npoin,2))
            for k in ("point_metrics_q", "point_metrics", "face_geom_q", "face_geom_n", "coord"):
                setattr(d, k, ptr(np.ascontiguousarray(deck[k], dtype=np.float64).reshape(-1), np.float64) if deck.get(k) is not None else None)
        self.h = C.c_void_p()
        rc = self.L.hnumo_init(C.byref(d), C.byref(self.h))
        if rc != 0:
            raise HnumoError("hnumo_init failed (%d): %s" % (rc, self.L.hnumo_last_error().decode()))
        self.npoin, self.nl = int(deck["npoin"]), int(deck["nlayers"])

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.hnumo_finalize(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc < 0:
            raise HnumoError("%s failed (%d): %s" % (what, rc, self.L.hnumo_last_error().decode()))
        return rc

    def upload_state(self, q_df, qb_df, qprime_df):
        q = np.ascontiguousarray(q_df, dtype=np.float64)
        qb = np.ascontiguousarray(qb_df, dtype=np.float64)
        qp = np.ascontiguousarray(qprime_df, dtype=np.float64)
        return self._check(self.L.hnumo_upload_state(self.h, q.ctypes.data, qb.ctypes.data, qp.ctypes.data), "upload_state")

    def download_state(self, out=None):
        if out is None:
            out = (np.empty((self.nl, self.npoin, 3)), np.empty((self.npoin, 4)), np.empty((self.nl, self.npoin, 3)))
        q, qb, qp = out
        self._check(self.L.hnumo_download_state(self.h, q.ctypes.data, qb.ctypes.data, qp.ctypes.data), "download_state")
        return q, qb, qp

    def step(self, n=1):
        """n calls of ti_rk_bcl on the resident state; returns 1 if a layer thickness went negative."""
        return self._check(self.L.hnumo_step(self.h, int(n)), "step")

    def ti_rk_bcl(self, q_df, qb_df, qprime_df):
        """Drop-in with the reference signature: arrays are updated in place (host buffers)."""
        return self._check(self.L.hnumo_ti_rk_bcl(self.h, q_df.ctypes.data, qb_df.ctypes.data, qprime_df.ctypes.data), "ti_rk_bcl")

    def btp_bcl_coeffs(self):
        return self._check(self.L.hnumo_btp_bcl_coeffs(self.h), "btp_bcl_coeffs")

    def btp_substeps(self):
        return self._check(self.L.hnumo_btp_substeps(self.h), "btp_substeps")

    def rhs_btp(self):
        out = np.empty((self.npoin, 3))
        self._check(self.L.hnumo_rhs_btp(self.h, out.ctypes.data), "rhs_btp")
        return out

    def layer_mass_rhs(self):
        """dp_advec[k, I] of layer_mass_rhs on the resident state (per-phase entry)"""
        out = np.empty((self.nl, self.npoin))
        self._check(self.L.hnumo_layer_mass_rhs(self.h, out.ctypes.data), "layer_mass_rhs")
        return out

    def layer_momentum_rhs(self):
        """rhs_mom[k, I, 0:2] of layer_momentum_rhs on the resident state (per-phase entry)"""
        out = np.empty((self.nl, self.npoin, 2))
        self._check(self.L.hnumo_layer_momentum_rhs(self.h, out.ctypes.data), "layer_momentum_rhs")
        return out

    def layer_shear_stress(self):
        """rhs_stress[k, I, 0:2] of rhs_layer_shear_stress on the resident q_df (per-phase entry, ad_mlswe > 0)"""
        out = np.empty((self.nl, self.npoin, 2))
        self._check(self.L.hnumo_layer_shear_stress(self.h, out.ctypes.data), "layer_shear_stress")
        return out

    def halo_exchange(self, nodal):
        """nodal[I, v] -> halo[h, n, v]: the neighbour's values at the nodes of this rank's processor faces (collective)"""
        nodal = np.ascontiguousarray(nodal, dtype=np.float64).reshape(self.npoin, -1)
        nv = nodal.shape[1]
        nh = len(self.deck["nbh_send_recv"])
        out = np.zeros((max(nh, 1), self.deck["ngl"], nv))
        n = self.L.hnumo_halo_exchange(self.h, nodal.ctypes.data, nv, out.ctypes.data)
        if n < 0:
            raise HnumoError("halo_exchange failed (%d): %s" % (n, self.L.hnumo_last_error().decode()))
        return out[:n]

    def get_array(self, name):
        cap = 8 * max(self.deck["nelem"] * self.deck["nq"] ** 2, self.deck["nface"] * self.deck["nq"] * 4)
        out = np.empty(cap)
        n = self.L.hnumo_get_array(self.h, name.encode(), out.ctypes.data, cap)
        if n < 0:
            raise HnumoError("get_array(%s) failed (%d): %s" % (name, n, self.L.hnumo_last_error().decode()))
        return out[:n].copy()

    def comm_init(self, id128):
        buf = (C.c_char * 128).from_buffer_copy(bytes(id128).ljust(128, b"\0"))
        return self._check(self.L.hnumo_comm_init(self.h, buf), "comm_init")

    def timing(self, reset=False):
        out = (C.c_double * 8)()
        self.L.hnumo_timing(self.h, out, 1 if reset else 0)
        return dict(ms_btp=out[0], stages=int(out[1]), ms_step=out[2], steps=int(out[3]), launches=int(out[4]),
                    ms_btp_last=out[5], ms_step_last=out[6])

    def diagnostics(self):
        """Device-side diagnostics of the resident state (hnumo_diagnostics): dict of per-layer mass and (max, min) of
        h, u, v, dp, elevation; (max, min) of qb; cfl_b, cfl, min_dx, min_dy."""
        nl = self.deck["nlayers"]
        out = np.zeros(11 * nl + 12)
        n = self.L.hnumo_diagnostics(self.h, out.ctypes.data_as(C.c_void_p), C.c_int64(out.size))
        if n < 0:
            raise HnumoError("diagnostics failed (%d): %s" % (n, self.L.hnumo_last_error().decode()))
        per = out[:11 * nl].reshape(nl, 11)
        d = dict(mass=per[:, 0].copy())
        for i, f in enumerate(("h", "u", "v", "dp", "ssh")):
            d[f] = np.stack([per[:, 1 + i], per[:, 6 + i]], axis=1)   # [layer] -> (max, min)
        t = out[11 * nl:]
        d["qb"] = np.stack([t[0:4], t[4:8]], axis=1)
        d.update(cfl_b=float(t[8]), cfl=float(t[9]), min_dx=float(t[10]), min_dy=float(t[11]))
        return d

    def set_option(self, key, value):
        return self._check(self.L.hnumo_set_option(self.h, key.encode(), float(value)), "set_option")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def snapshot_write(path, deck, q_df, qb_df):
    """Reference-format text snapshot (`mlswe####`, diagnostics.F90:73-91) of a state in the reference layouts."""
    L = load_library()
    q = np.ascontiguousarray(q_df, dtype=np.float64); qb = np.ascontiguousarray(qb_df, dtype=np.float64)
    coord = np.ascontiguousarray(deck["coord"], dtype=np.float64)
    zb = np.ascontiguousarray(deck["zbot_df"], dtype=np.float64); al = np.ascontiguousarray(deck["alpha_mlswe"], dtype=np.float64)
    rc = L.hnumo_snapshot_write(str(path).encode(), deck["nlayers"], deck["npoin"], deck["dt"], deck["dt_btp"], _ptr(coord), _ptr(q), _ptr(qb),
                                _ptr(zb), _ptr(al), deck["gravity"])
    if rc != 0:
        raise HnumoError("snapshot_write failed (%d): %s" % (rc, L.hnumo_last_error().decode()))


def snapshot_write_nc(path, deck, q_df, qb_df):
    """NetCDF snapshot `mlswe####.nc` with the layout of the reference's diagnostics_nc (src/diagnostics_nc.F90:98-165)."""
    L = load_library()
    q = np.ascontiguousarray(q_df, dtype=np.float64); qb = np.ascontiguousarray(qb_df, dtype=np.float64)
    coord = np.ascontiguousarray(deck["coord"], dtype=np.float64)
    zb = np.ascontiguousarray(deck["zbot_df"], dtype=np.float64); al = np.ascontiguousarray(deck["alpha_mlswe"], dtype=np.float64)
    pbp = np.ascontiguousarray(deck["pbprime_df"], dtype=np.float64)
    rc = L.hnumo_snapshot_write_nc(str(path).encode(), deck["nlayers"], deck["npoin"], deck["dt"], deck["dt_btp"], _ptr(coord), _ptr(q), _ptr(qb),
                                   _ptr(zb), _ptr(al), _ptr(pbp), deck["gravity"])
    if rc != 0:
        raise HnumoError("snapshot_write_nc failed (%d): %s" % (rc, L.hnumo_last_error().decode()))


def snapshot_read_nc_restart(path, deck):
    """(q_df, qb_df, qprime_df, coord) rebuilt from a NetCDF snapshot, as restart_mlswe rebuilds them from the text one."""
    L = load_library()
    nl, npn = deck["nlayers"], deck["npoin"]
    q = np.zeros((nl, npn, 3)); qb = np.zeros((npn, 4)); qp = np.zeros((nl, npn, 3)); coord = np.zeros((npn, 2))
    pbp = np.ascontiguousarray(deck["pbprime_df"], dtype=np.float64); al = np.ascontiguousarray(deck["alpha_mlswe"], dtype=np.float64)
    rc = L.hnumo_snapshot_read_nc_restart(str(path).encode(), nl, npn, _ptr(pbp), _ptr(al), deck["gravity"], _ptr(q), _ptr(qb), _ptr(qp), _ptr(coord))
    if rc != 0:
        raise HnumoError("snapshot_read_nc_restart failed (%d): %s" % (rc, L.hnumo_last_error().decode()))
    return q, qb, qp, coord


def snapshot_info(path):
    L = load_library()
    nl, npn, dt, dtb = C.c_int32(), C.c_int64(), C.c_double(), C.c_double()
    rc = L.hnumo_snapshot_info(str(path).encode(), C.byref(nl), C.byref(npn), C.byref(dt), C.byref(dtb))
    if rc != 0:
        raise HnumoError("snapshot_info failed (%d): %s" % (rc, L.hnumo_last_error().decode()))
    return dict(nlayers=nl.value, npoin=npn.value, dt=dt.value, dt_btp=dtb.value)


def snapshot_read_restart(path, deck):
    """restart_mlswe (mod_restart.F90:15-66): (q_df, qb_df, qprime_df, coord) rebuilt from a text snapshot for upload_state."""
    L = load_library()
    nl, npn = deck["nlayers"], deck["npoin"]
    q = np.zeros((nl, npn, 3)); qb = np.zeros((npn, 4)); qp = np.zeros((nl, npn, 3)); coord = np.zeros((npn, 2))
    pbp = np.ascontiguousarray(deck["pbprime_df"], dtype=np.float64); al = np.ascontiguousarray(deck["alpha_mlswe"], dtype=np.float64)
    rc = L.hnumo_snapshot_read_restart(str(path).encode(), nl, npn, _ptr(pbp), _ptr(al), deck["gravity"], _ptr(q), _ptr(qb), _ptr(qp), _ptr(coord))
    if rc != 0:
        raise HnumoError("snapshot_read_restart failed (%d): %s" % (rc, L.hnumo_last_error().decode()))
    return q, qb, qp, coord


def _fortran_e(x, width, digits):
    """Fortran Ew.d edit descriptor: 0.ddddE+XX, right-justified"""
    x = float(x)
    if x == 0.0:
        body = "0." + "0" * digits + "E+00"
    else:
        m, e = ("%.*e" % (digits - 1, abs(x))).split("e")
        ex = int(e) + 1
        body = ("-" if x < 0 else "") + "0." + m.replace(".", "") + ("E%+03d" % ex if abs(ex) < 100 else "%+04d" % ex)   # three-digit exponents drop the letter
    return body.rjust(width)


def write_fin(path, diag, mass0):
    """`mlswe_FIN.txt` as the reference writes it at the end of a run (print_diagnostics.F90:163-184): per layer the relative
    mass loss |m - m0| / m0 and the max/min of h, u, v, ssh -- from the device-side diagnostics (Solver.diagnostics()) of the
    final state and the layer masses of the initial state.  The reference's CI compares exactly this file (CI/bump/check.F90)."""
    with open(path, "w") as f:
        for k in range(len(diag["mass"])):
            f.write("Layer = %8d\n" % (k + 1))
            f.write("Mass Loss  = %s \n" % _fortran_e(abs(diag["mass"][k] - mass0[k]) / mass0[k], 16, 8))
            for name in ("h", "u", "v", "ssh"):
                f.write("Fields:   Max/Min = %-3s %s %s \n" % (name, _fortran_e(diag[name][k][0], 24, 12), _fortran_e(diag[name][k][1], 24, 12)))


def nccl_unique_id():
    buf = (C.c_char * 128)()
    rc = load_library().hnumo_comm_get_unique_id(buf)
    if rc != 0:
        raise HnumoError("hnumo_comm_get_unique_id failed: %s" % load_library().hnumo_last_error().decode())
    return bytes(buf)


def local_group_id(gid):
    """128-byte id selecting the in-process back end (several partitions on one GPU; tests only)."""
    import struct
    return b"LOCAL\0\0\0" + struct.pack("i", gid)
