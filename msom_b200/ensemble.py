"""Ensemble of independent msqg members (BASELINE config 5: stochastic forcing, members batched
per GPU).  The reference has no ensemble driver: one process is one member, with its own libc
rand() stream.  Here each member is one handle (own CUDA stream, own random_r state seeded like
srand(seed)), members of a GPU run concurrently from a thread pool (the C ABI releases the GIL),
and GPUs hold disjoint member sets -- replicas only, no communication."""
from concurrent.futures import ThreadPoolExecutor

from . import capi as G


class Ensemble:
    def __init__(self, params, nmembers, device=0, seeds=None, threads=None, noise="libc", smoother="lex"):
        """noise: 'libc' replays the reference's rand() stream on the host (parity; ~80 ns of CPU per sample, which at
        512^2 x 3 is 60 ms per step and member -- the GPU idles), 'philox' draws the field on the device (production).
        smoother: 'lex' (reference order) or 'rb' (throughput mode, no cooperative launches: members overlap freely)."""
        self.members = [G.Model(params, device) for _ in range(nmembers)]
        self.seeds = list(seeds) if seeds is not None else [1000 + i for i in range(nmembers)]
        for m, s in zip(self.members, self.seeds):
            m.L.msqg_seed_noise(m.h, s)
            G.check(m.L.msqg_set_noise_mode(m.h, {"libc": 0, "philox": 1}[noise]))
            m.set_smoother(smoother)
        self.pool = ThreadPoolExecutor(max_workers=threads or nmembers)

    def __len__(self):
        return len(self.members)

    def each(self, fn):
        """fn(member_index, model) on every member concurrently; returns the list of results."""
        return list(self.pool.map(lambda im: fn(*im), enumerate(self.members)))

    def set(self, fid, arrays):
        """arrays: one array for all members, or a list with one per member"""
        per = arrays if isinstance(arrays, (list, tuple)) else [arrays] * len(self.members)
        self.each(lambda i, m: m.set(fid, per[i]))

    def set_const(self):
        self.each(lambda i, m: m.set_const())

    def step(self, nsteps=1):
        def run(i, m):
            return [m.step() for _ in range(nsteps)]
        return self.each(run)

    def get(self, fid):
        return self.each(lambda i, m: m.get(fid))

    def close(self):
        self.pool.shutdown(wait=True)
        for m in self.members:
            m.close()
