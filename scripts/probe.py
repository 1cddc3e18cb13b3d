"""Quick timing probe: python scripts/probe.py N nl nsteps"""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from common import base_kw, synth_psi
from msom_b200 import capi as G
N = int(sys.argv[1]); nl = int(sys.argv[2]); nsteps = int(sys.argv[3])
m = G.Model(G.make_params(**base_kw(N, nl)))
m.set(G.PSI, synth_psi(N, nl)); m.set_const()
for i in range(3):
    m.step()
t0 = time.time()
c0 = m.total_cycles; l0 = m.launches
for i in range(nsteps):
    dt = m.step()
t1 = time.time()
s = m.mgstats()
print("N=%d nl=%d: %.3f ms/step, %.3f G cell-layer updates/s, cycles/step %.2f, launches/step %.1f, dt=%g nrelax=%d resa=%g" % (
    N, nl, (t1 - t0) / nsteps * 1e3, N * N * nl * nsteps / (t1 - t0) / 1e9, (m.total_cycles - c0) / nsteps, (m.launches - l0) / nsteps, dt, s.nrelax, s.resa))
for nr in (1, 2, 4, 8):
    print("  vcycle nrelax=%d: %.3f ms" % (nr, m.time_vcycle(nr, 5)))
