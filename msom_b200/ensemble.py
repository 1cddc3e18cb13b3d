"""Ensemble of independent msqg members (BASELINE config 5: stochastic forcing, members batched
per GPU).  The reference has no ensemble driver: one process is one member, with its own libc
rand() stream.  Here each member is one handle (own CUDA stream, own noise stream seeded like
srand(seed)) and the members of a GPU are advanced concurrently by the library's own ensemble
driver (msqg_ensemble_* in include/msqg.h: one persistent host thread per member, C++); GPUs hold
disjoint member sets -- replicas only, no communication.  This file is the
ctypes mirror of that driver."""
import ctypes as C

import numpy as np

from . import capi as G


class Ensemble:
    def __init__(self, params, nmembers, device=0, seeds=None, threads=None, noise="libc", smoother="lex"):
        """noise: 'libc' replays the reference's rand() stream on the host (parity; ~80 ns of CPU per sample, which at
        512^2 x 3 is 60 ms per step and member -- the GPU idles), 'philox' draws the field on the device (production).
        smoother: 'lex' (reference order) or 'rb' (throughput mode, no cooperative launches: members overlap freely).
        `threads` is kept for source compatibility (the driver always runs one host thread per member)."""
        self.L = G.lib()
        self.seeds = list(seeds) if seeds is not None else [1000 + i for i in range(nmembers)]
        h = C.c_void_p()
        sd = (C.c_uint * nmembers)(*self.seeds)
        G.check(self.L.msqg_ensemble_create(C.byref(params), device, nmembers, sd, {"libc": 0, "philox": 1}[noise],
                                            {"lex": 0, "rb": 1}[smoother], C.byref(h)))
        self.h = h
        self.members = [G.Model(params, device, handle=self.L.msqg_ensemble_member(h, k)) for k in range(nmembers)]

    def __len__(self):
        return len(self.members)

    def each(self, fn):
        """fn(member_index, model) on every member (host-side work: field transfers); returns the list of results."""
        return [fn(i, m) for i, m in enumerate(self.members)]

    def set(self, fid, arrays):
        """arrays: one array for all members, or a list with one per member"""
        per = arrays if isinstance(arrays, (list, tuple)) else [arrays] * len(self.members)
        self.each(lambda i, m: m.set(fid, per[i]))

    def set_const(self):
        G.check(self.L.msqg_ensemble_set_const(self.h))

    def step(self, nsteps=1):
        """nsteps steps of every member, members concurrently; returns the per-member lists of time steps"""
        n = len(self.members)
        out = [[] for _ in range(n)]
        dts = np.zeros(n)
        if nsteps > 8:     # long runs in one call (the dt history is not kept)
            G.check(self.L.msqg_ensemble_step(self.h, nsteps, dts.ctypes.data_as(C.POINTER(C.c_double))))
            return [[float(d)] for d in dts]
        for _ in range(nsteps):
            G.check(self.L.msqg_ensemble_step(self.h, 1, dts.ctypes.data_as(C.POINTER(C.c_double))))
            for k in range(n):
                out[k].append(float(dts[k]))
        return out

    def get(self, fid):
        return self.each(lambda i, m: m.get(fid))

    def close(self):
        if getattr(self, "h", None):
            for m in self.members:
                m.close()          # non-owning wrappers
            self.L.msqg_ensemble_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
