"""single GPU, local 2x1 / 2x2 group: cost of the exchange kernels themselves (no inter-GPU skew)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import base_kw, synth_psi
from msom_b200 import capi as G
from msom_b200.dist import Group
N, nl = 2048, 4
for px, py in ((2, 1), (2, 2)):
    g = Group(G.make_params(**base_kw(N, nl)), px, py, 512, 0, smoother="rb")
    g.set_global(G.PSI, synth_psi(N, nl)); g.set_const()
    for _ in range(3):
        g.step()
    g.profile(True)
    x0 = g.exchanges
    for _ in range(5):
        g.step()
    prof = g.profile_read(); g.profile(False)
    v = prof["exchange"]
    print("%dx%d local tiles %s: exchange %.3f ms/step over %.1f scopes/step = %.1f us each (exchanges/step %.1f)" % (
        px, py, g.transport, v["ms"] / 5, v["count"] / 5, 1e3 * v["ms"] / max(v["count"], 1), (g.exchanges - x0) / 5))
    g.close()
