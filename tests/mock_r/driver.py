"""Builds r-package/TADpoleB200/src/r_shim.c against the toy R runtime in this directory and drives its .Call entry
points through ctypes, the way an R session would (TEST INFRASTRUCTURE: R itself is not installed here)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SHIM = os.path.join(ROOT, "r-package", "TADpoleB200", "src", "r_shim.c")
OUT = os.path.join(HERE, "r_shim_mock.so")
LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP, RAWSXP = 10, 13, 14, 16, 19, 24
NA_REAL_BITS = 0x7FF00000000007A2


class RError(RuntimeError):
    pass


class Sexp(int):
    """An opaque SEXP handle (external pointers, raw results)."""


class MockR:
    def __init__(self):
        libdir = os.path.join(ROOT, "tadpole_b200")
        srcs = [SHIM, os.path.join(HERE, "mock_r.c")]
        if not os.path.exists(OUT) or any(os.path.getmtime(s) > os.path.getmtime(OUT) for s in srcs):
            subprocess.check_call(["gcc", "-O1", "-std=c11", "-fPIC", "-shared", "-Wall", "-Wno-cast-function-type", "-Werror",
                                   "-I" + HERE, "-I" + os.path.join(ROOT, "include"), "-o", OUT, *srcs,
                                   "-L" + libdir, "-ltadpole_b200", "-Wl,-rpath," + libdir])
        self.lib = L = ctypes.CDLL(OUT)
        vp = ctypes.c_void_p
        L.mock_dot_call.restype = vp
        L.mock_dot_call.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_char_p)]
        for f in ("mock_vector", "mock_matrix", "mock_string", "mock_nil", "mock_data", "mock_elt"):
            getattr(L, f).restype = vp
        L.mock_vector.argtypes = [ctypes.c_int, ctypes.c_int, vp]
        L.mock_matrix.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
        L.mock_string.argtypes = [ctypes.c_char_p]
        for f in ("mock_type", "mock_len", "mock_nrow", "mock_ncol", "mock_data"):
            getattr(L, f).argtypes = [vp]
        L.mock_elt.argtypes = [vp, ctypes.c_int]
        L.mock_registered.argtypes = [ctypes.c_char_p]
        L.mock_init()

    # ---- R values in ----
    def to_r(self, x):
        L = self.lib
        if x is None:
            return L.mock_nil()
        if isinstance(x, Sexp):
            return int(x)                                         # already a SEXP (external pointer, ...)
        if isinstance(x, str):
            return L.mock_string(x.encode())
        if isinstance(x, list):                                   # an R list (VECSXP)
            elts = [self.to_r(e) for e in x]
            arr = (ctypes.c_void_p * max(len(elts), 1))(*elts)
            return L.mock_vector(VECSXP, len(elts), ctypes.cast(arr, ctypes.c_void_p))
        if isinstance(x, bool):
            return L.mock_vector(LGLSXP, 1, np.array([int(x)], np.int32).ctypes.data)
        if isinstance(x, int):
            return L.mock_vector(INTSXP, 1, np.array([x], np.int32).ctypes.data)
        if isinstance(x, float):
            return L.mock_vector(REALSXP, 1, np.array([x]).ctypes.data)
        a = np.asarray(x)
        if a.dtype == np.uint8:
            a = np.ascontiguousarray(a)
            return L.mock_vector(RAWSXP, a.size, a.ctypes.data)
        typ, a = (REALSXP, a.astype(np.float64)) if a.dtype.kind == "f" else (INTSXP, a.astype(np.int32))
        if a.ndim == 2:                                           # R matrices are column-major
            f = np.asfortranarray(a)
            return L.mock_matrix(typ, a.shape[0], a.shape[1], f.ctypes.data)
        a = np.ascontiguousarray(a)
        return L.mock_vector(typ, a.size, a.ctypes.data if a.size else None)

    # ---- R values out ----
    def from_r(self, s):
        L = self.lib
        t, n = L.mock_type(s), L.mock_len(s)
        if t == 0:
            return None
        if t == VECSXP:
            return [self.from_r(L.mock_elt(s, i)) for i in range(n)]
        if t == 22:
            return Sexp(s)
        if t == STRSXP:
            return [ctypes.string_at(L.mock_data(L.mock_elt(s, i))).decode() for i in range(n)]
        dt = {LGLSXP: np.int32, INTSXP: np.int32, REALSXP: np.float64, RAWSXP: np.uint8}[t]
        if n == 0:
            flat = np.zeros(0, dt)
        else:
            flat = np.ctypeslib.as_array(ctypes.cast(L.mock_data(s), ctypes.POINTER(np.ctypeslib.as_ctypes_type(dt))), (n,)).copy()
        nr, nc = L.mock_nrow(s), L.mock_ncol(s)
        if nc != 1 or nr != n:
            return flat.reshape((nc, nr)).T                       # column-major -> [nrow, ncol]
        return flat.astype(bool) if t == LGLSXP else flat

    def call(self, name, *args, raw=False):
        """.Call(name, ...): raises RError with the message R's error() carried."""
        sx = [self.to_r(a) for a in args]
        arr = (ctypes.c_void_p * max(len(sx), 1))(*sx)
        err = ctypes.c_char_p()
        out = self.lib.mock_dot_call(name.encode(), len(sx), arr, ctypes.byref(err))
        if err.value is not None:
            raise RError(err.value.decode())
        return Sexp(out) if raw else self.from_r(out)

    def registered(self, name):
        return self.lib.mock_registered(name.encode())

    def reset(self):
        """End of the R session: finalizers run, every object is released."""
        self.lib.mock_reset()
