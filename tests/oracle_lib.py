"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIBS = {}

TEST_CASES = {"bump": 0, "lakeAtrest": 1, "double-gyre": 2, "double-gyre-synth": 3}


class Config(C.Structure):
    _fields_ = [
        ("nelx", C.c_int), ("nely", C.c_int), ("nop", C.c_int), ("nlayers", C.c_int),
        ("xdims", C.c_double * 2), ("ydims", C.c_double * 2),
        ("x_boundary", C.c_int * 2), ("y_boundary", C.c_int * 2),
        ("dt", C.c_double), ("dt_btp", C.c_double),
        ("kstages", C.c_int), ("botfr", C.c_int), ("cd_mlswe", C.c_double),
        ("method_visc", C.c_int), ("visc_mlswe", C.c_double),
        ("f0", C.c_double), ("beta", C.c_double),
        ("test_case", C.c_int), ("dg_integ_exact", C.c_int),
        ("synth_z", C.c_double * 33), ("synth_alpha", C.c_double * 32), ("synth_perturb", C.c_double),
        ("affine_metrics", C.c_int),
        ("ad_mlswe", C.c_double), ("max_shear_dz", C.c_double),
        ("mesh_warp", C.c_double),
    ]


def build(force=False, variant=""):
    """variant "" = the oracle proper; "fma" = the same sources compiled with FP contraction (oracle/Makefile)"""
    so = os.path.join(_ORACLE_DIR, "liboracle%s.so" % ("_" + variant if variant else ""))
    srcs = [os.path.join(_ORACLE_DIR, f) for f in os.listdir(_ORACLE_DIR) if f.endswith((".cpp", ".hpp"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "-B", os.path.basename(so)], stdout=subprocess.DEVNULL)
    return so


def set_threads(n):
    """OpenMP threads of the oracle from now on (torchrun exports OMP_NUM_THREADS=1; bench.py sets the count explicitly)"""
    lib().orc_set_threads(int(n))


def max_threads():
    return int(lib().orc_get_max_threads())


def lib(variant=""):
    if variant not in _LIBS:
        L = C.CDLL(build(variant=variant))
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_get_max_threads.restype = C.c_int
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(Config)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_info.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.orc_scalars.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.orc_array_size.restype = C.c_long
        L.orc_array_size.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_get_array.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.orc_set_array.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.orc_get_face.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_step.argtypes = [C.c_void_p, C.c_int]
        L.orc_btp_bcl_coeffs.argtypes = [C.c_void_p]
        L.orc_rhs_btp.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_btp_substeps.argtypes = [C.c_void_p]
        L.orc_layer_mass_rhs.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_layer_momentum_rhs.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_shear_stress.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_diagnostics.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _LIBS[variant] = L
    return _LIBS[variant]


def make_config(deck):
    """deck: dict with the numo3d.in keys the hot path reads (see h-numo_b200 decks)."""
    c = Config()
    c.nelx, c.nely, c.nop, c.nlayers = deck["nelx"], deck["nely"], deck["nop"], deck["nlayers"]
    c.xdims[0], c.xdims[1] = deck["xdims"]
    c.ydims[0], c.ydims[1] = deck["ydims"]
    c.x_boundary[0], c.x_boundary[1] = deck.get("x_boundary", (4, 4))
    c.y_boundary[0], c.y_boundary[1] = deck.get("y_boundary", (4, 4))
    c.dt, c.dt_btp = deck["dt"], deck["dt_btp"]
    c.kstages = deck.get("kstages", 5)
    c.botfr = deck.get("botfr", 0)
    c.cd_mlswe = deck.get("cd_mlswe", 0.0)
    c.method_visc = deck.get("method_visc", 0)
    c.visc_mlswe = deck.get("visc_mlswe", 0.0)
    c.f0, c.beta = deck.get("f0", 0.0), deck.get("beta", 0.0)
    c.test_case = TEST_CASES[deck["test_case"]]
    c.dg_integ_exact = 1 if deck.get("dg_integ_exact", True) else 0
    for i, z in enumerate(deck.get("synth_z", [])):
        c.synth_z[i] = z
    for i, a in enumerate(deck.get("synth_alpha", [])):
        c.synth_alpha[i] = a
    c.synth_perturb = deck.get("synth_perturb", 0.0)
    c.affine_metrics = 1 if deck.get("affine_metrics", False) else 0
    c.ad_mlswe, c.max_shear_dz = deck.get("ad_mlswe", 0.0), deck.get("max_shear_dz", 0.0)
    c.mesh_warp = deck.get("mesh_warp", 0.0)
    return c


class Oracle:
    def __init__(self, deck, variant=""):
        self.L = lib(variant)
        self.cfg = make_config(deck)
        self.h = self.L.orc_create(C.byref(self.cfg))
        info = (C.c_int * 10)()
        self.L.orc_info(self.h, info)
        (self.ngl, self.nq, self.nelem, self.npoin, self.npoin_q, self.nface, self.nl, self.N_btp,
         self.kstages, _) = list(info)
        sc = (C.c_double * 6)()
        self.L.orc_scalars(self.h, sc)
        self.dt, self.dt_btp, self.gravity = sc[0], sc[1], sc[2]

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def error_flag(self):
        info = (C.c_int * 10)()
        self.L.orc_info(self.h, info)
        return info[9]

    def timing(self):
        sc = (C.c_double * 6)()
        self.L.orc_scalars(self.h, sc)
        return sc[3], int(sc[4])  # seconds in the barotropic loop, barotropic stages executed

    def get(self, name):
        n = self.L.orc_array_size(self.h, name.encode())
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float64)
        self.L.orc_get_array(self.h, name.encode(), out.ctypes.data)
        return out

    def set(self, name, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64).ravel()
        n = self.L.orc_array_size(self.h, name.encode())
        assert a.size == n, (name, a.size, n)
        self.L.orc_set_array(self.h, name.encode(), a.ctypes.data)

    def face(self):
        out = np.empty(8 * self.nface, dtype=np.int32)
        self.L.orc_get_face(self.h, out.ctypes.data)
        return out.reshape(self.nface, 8)

    def step(self, n=1):
        return self.L.orc_step(self.h, n)

    def btp_bcl_coeffs(self):
        self.L.orc_btp_bcl_coeffs(self.h)

    def rhs_btp(self):
        out = np.empty(3 * self.npoin)
        self.L.orc_rhs_btp(self.h, out.ctypes.data)
        return out.reshape(self.npoin, 3)

    def btp_substeps(self):
        self.L.orc_btp_substeps(self.h)

    def layer_mass_rhs(self):
        out = np.empty(self.npoin * self.nl)
        self.L.orc_layer_mass_rhs(self.h, out.ctypes.data)
        return out.reshape(self.nl, self.npoin)

    def layer_momentum_rhs(self):
        out = np.empty(2 * self.npoin * self.nl)
        self.L.orc_layer_momentum_rhs(self.h, out.ctypes.data)
        return out.reshape(self.nl, self.npoin, 2)

    def shear_stress(self):
        out = np.empty(2 * self.npoin * self.nl)
        self.L.orc_shear_stress(self.h, out.ctypes.data)
        return out.reshape(self.nl, self.npoin, 2)

    def diagnostics(self):
        q = np.empty(5 * self.npoin * self.nl)
        m = np.empty(self.nl)
        self.L.orc_diagnostics(self.h, q.ctypes.data, m.ctypes.data)
        return q.reshape(self.nl, self.npoin, 5), m
