import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import base_kw, synth_psi
from msom_b200 import capi as G
from msom_b200.dist import Group
N, nl, px, py, agg = 128, 2, 2, 1, 64
kw = base_kw(N, nl); psi = synth_psi(N, nl)
s = G.Model(G.make_params(**kw), 0); s.set_smoother("rb")
g = Group(G.make_params(**kw), px, py, agg, 0, smoother="rb")
s.set(G.PSI, psi); g.set_global(G.PSI, psi); s.set_const(); g.set_const()
def cmp(tag):
    for name, f in (("psi", G.PSI), ("q", G.Q), ("zeta", G.ZETA), ("tmp", G.TMP), ("qpred", G.QPRED)):
        a, b = g.get_global(f), s.get(f)
        d = np.argwhere(a != b)
        print(tag, name, "ndiff", len(d), "max", np.abs(a - b).max(), d[:6].tolist())
cmp("init")
for i in range(2):
    print("dt", g.step(), s.step(), g.total_cycles, s.total_cycles)
    cmp("step%d" % i)
