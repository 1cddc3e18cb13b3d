"""msom_b200 -- B200-native (sm_100a) multilayer quasi-geostrophic timestep.

Drop-in for the hot path of bderembl/msom's msqg (qg.c + qg.h + layer.h +
poisson_layer.h + eigmode.h).  Layers:
  csrc/   hand-written CUDA kernels + the handle-based C ABI (include/msqg.h)
  host/   plain-C host side mirroring the reference surface (read_params,
          set_vars, set_const, run, .bas I/O, pystep_bfn/pyq2p/pyp2q)
  capi.py ctypes binding of the C ABI;  qg.py mirror of the SWIG module `qg`
There is no CPU fallback: importing works anywhere, compute needs a B200.
"""
from . import capi  # noqa: F401
