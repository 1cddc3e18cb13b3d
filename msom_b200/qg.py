"""Mirror of the reference's SWIG module `qg` (msqg/qg.i, qg_bfn.i) over libqg.so.

Usage is the reference's own (msqg/qg_bfn.py:33-86):

    import msom_b200.qg as bas
    bas.read_params("params.in"); bas.init_grid(N); bas.set_vars(); bas.set_vars_bfn()
    bas.set_const(); bas.pyp2q(p, q); bas.pystep_bfn(var, F1, direction, flag_q); bas.pyq2p(p, q)
    bas.trash_vars(); bas.trash_vars_bfn()

Arrays are C-contiguous float64 (nl, N, N) = [layer][y][x]; `IN_ARRAY3`
arguments are read, `INPLACE_ARRAY3` arguments (tend / the converted field)
are written in place, exactly like the SWIG typemaps.  Functions raise
RuntimeError where the reference would exit(0).
"""
import ctypes as C
import os

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_QG = os.path.join(_HERE, "lib", "libqg.so")
_lib = None


def _L():
    global _lib
    if _lib is None:
        capi.lib()  # loads libmsqg_cuda.so (RTLD_GLOBAL) and points it at a LAPACK for eigmod
        if not os.path.exists(LIB_QG):
            raise ImportError("%s not built (no CPU fallback exists)" % LIB_QG)
        L = C.CDLL(LIB_QG)
        dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        L.read_params.argtypes = [C.c_char_p]
        L.init_grid.argtypes = [C.c_int]
        L.pystep_bfn.argtypes = [dp, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
        L.pyq2p.argtypes = [dp, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int, C.c_int]
        L.pyp2q.argtypes = [dp, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int, C.c_int]
        L.pystep_de.argtypes = [dp, C.c_int, C.c_int, C.c_int] * 7 + [C.c_int]
        L.qg_model.restype = C.c_void_p
        L.qg_params.restype = C.POINTER(capi.Params)
        L.qg_set_device.argtypes = [C.c_int]
        L.qg_set_mode_pv_invert.argtypes = [C.c_int]
        L.qg_set_stochastic.argtypes = [C.c_int]
        L.qg_set_verbose.argtypes = [C.c_int]
        L.qg_time.restype = C.c_double
        L.qg_outdir.restype = C.c_char_p
        L.qg_set_outdir.argtypes = [C.c_char_p]
        L.qg_run_iteration.argtypes = [C.c_int]
        L.qg_write_bas.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_double, dp]
        L.qg_read_bas.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_double, dp]
        _lib = L
    return _lib


def _ck(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, capi.lib().msqg_last_error().decode(errors="replace")))


def read_params(path2file):
    _ck(_L().read_params(str(path2file).encode()), "read_params")


def init_grid(n):
    _ck(_L().init_grid(int(n)), "init_grid")


def set_vars():
    _ck(_L().set_vars(), "set_vars")


def set_const():
    _ck(_L().set_const(), "set_const")


def create_outdir():
    _ck(_L().create_outdir(), "create_outdir")


def backup_config():
    _ck(_L().backup_config(), "backup_config")


def trash_vars():
    _ck(_L().trash_vars(), "trash_vars")


def set_vars_bfn():
    _ck(_L().set_vars_bfn(), "set_vars_bfn")


def trash_vars_bfn():
    _ck(_L().trash_vars_bfn(), "trash_vars_bfn")


def _in(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _inplace(a):
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.ndim == 3):
        raise TypeError("in-place argument must be a C-contiguous float64 array of shape (nl, N, N)")
    return a


def pystep_bfn(varin, tend, direction, vartype):
    v, t = _in(varin), _inplace(tend)
    _ck(_L().pystep_bfn(v, *v.shape, t, *t.shape, float(direction), int(vartype)), "pystep_bfn")


def pyq2p(po, qo):
    p, q = _inplace(po), _in(qo)
    _ck(_L().pyq2p(p, *p.shape, q, *q.shape), "pyq2p")


def pyp2q(po, qo):
    p, q = _in(po), _inplace(qo)
    _ck(_L().pyp2q(p, *p.shape, q, *q.shape), "pyp2q")


def set_vars_energy():
    _ck(_L().set_vars_energy(), "set_vars_energy")


def trash_vars_energy():
    _ck(_L().trash_vars_energy(), "trash_vars_energy")


def pystep_de(po, de_bf, de_vd, de_j1, de_j2, de_j3, de_ft, onlyKE=0):
    """qg_energy.i:30-39: energy tendencies of the state po (all de_* arrays are written in place)"""
    args = [_in(po)] + [_inplace(a) for a in (de_bf, de_vd, de_j1, de_j2, de_j3, de_ft)]
    flat = []
    for a in args:
        flat += [a, *a.shape]
    _ck(_L().pystep_de(*flat, int(onlyKE)), "pystep_de")


def run():
    _ck(_L().run(), "run")


# extras (not in the SWIG module): compile-time switches of the reference as runtime knobs
def set_device(d):
    _L().qg_set_device(int(d))


def set_mode_pv_invert(mode):
    _L().qg_set_mode_pv_invert(int(mode))


def set_stochastic(on):
    _L().qg_set_stochastic(int(on))


def set_verbose(v):
    _L().qg_set_verbose(int(v))


def params():
    return _L().qg_params().contents
