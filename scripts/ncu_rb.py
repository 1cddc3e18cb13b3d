"""One finest-level rb relax pass (4 sweeps) + a few steps, for ncu captures."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import base_kw, synth_psi
from msom_b200 import capi as G
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 4
m = G.Model(G.make_params(**base_kw(N, nl)), 0)
m.set_smoother(sys.argv[3] if len(sys.argv) > 3 else "rb")
m.set(G.PSI, synth_psi(N, nl))
m.set_const()
for _ in range(int(sys.argv[4]) if len(sys.argv) > 4 else 2):
    m.step()
print("ok", m.total_cycles, m.launches)
