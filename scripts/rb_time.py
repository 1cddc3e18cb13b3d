"""Quick timing of the two smoothers on one GPU (not the bench): ms/step, cycles, per-category kernel times."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import time
import numpy as np
from common import base_kw, synth_psi
from msom_b200 import capi as G

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 4
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["rb", "lex"]
over = {}
if len(sys.argv) > 5 and sys.argv[5] == "modal":
    over["mode_pv_invert"] = 1
for sm in modes:
    m = G.Model(G.make_params(**base_kw(N, nl, **over)), 0)
    m.set_smoother(sm)
    m.set(G.PSI, synth_psi(N, nl))
    m.set_const()
    for _ in range(3):
        m.step()
    m.profile(True)
    c0, l0 = m.total_cycles, m.launches
    t0 = time.perf_counter()
    for _ in range(steps):
        m.step()
    dt = (time.perf_counter() - t0) / steps
    prof = m.profile_read()
    m.profile(False)
    print("%s N=%d nl=%d: %.2f ms/step (with profiling events), %.2f G cell-layer/s, cycles/step %.1f, launches/step %.0f" % (
        sm, N, nl, dt * 1e3, N * N * nl / dt / 1e9, (m.total_cycles - c0) / steps, (m.launches - l0) / steps))
    for k, v in prof.items():
        print("   %-13s %8.3f ms/step  %5d launches  aux %d" % (k, v["ms"] / steps, v["count"] / steps, v["aux"] / steps))
    t0 = time.perf_counter()
    for _ in range(steps):
        m.step()
    dt = (time.perf_counter() - t0) / steps
    print("   without profiling: %.2f ms/step = %.2f G/s" % (dt * 1e3, N * N * nl / dt / 1e9))
    try:
        print("   vcycle (nrelax 4): %.3f ms" % m.time_vcycle(4, 5))
    except Exception as e:
        print("   vcycle:", e)
    m.close()
