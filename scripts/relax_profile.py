import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import ctypes as C
import numpy as np
from common import base_kw, synth_psi
from msom_b200 import capi as G
N = int(sys.argv[1]); nl = int(sys.argv[2]); nsw = int(sys.argv[3])
m = G.Model(G.make_params(**base_kw(N, nl)))
m.set(G.PSI, synth_psi(N, nl)); m.set_const()
L = G.lib()
L.msqg_test_relax_profile.argtypes = [C.c_void_p, C.c_int, C.c_int, np.ctypeslib.ndpointer(dtype=np.int64), C.c_int]
level = int(np.log2(N))
K = 4 if nsw <= 4 else 8; W = 32 // K
nw = (N + K - 1 + W - 1) // W
for rep in range(2):
    out = np.zeros((nw * 12,), dtype=np.int64)
    G.check(L.msqg_test_relax_profile(m.h, level, nsw, out, nw))
ext = out[nw * 4:].reshape(nw, 8); out = out[:nw * 4].reshape(nw, 4)
t0 = out[:, 0].min()
st = (out[:, 0] - t0) / 1e3; en = (out[:, 1] - t0) / 1e3
print("N=%d nl=%d nsweeps=%d workers=%d steps/worker=%d" % (N, nl, nsw, nw, N + W + K - 1))
for w in list(range(0, min(nw, 6))) + list(range(nw // 2, nw // 2 + 2)) + [nw - 2, nw - 1]:
    print(" w=%4d start %9.1f us end %9.1f us dur %9.1f us spins %d" % (w, st[w], en[w], en[w] - st[w], out[w, 2]))
print("end diffs: median %.2f us" % np.median(np.diff(en)))
print("total %.1f us; worker0 %.3f us/step; mean end-lag between workers %.2f us" % (en.max(), (en[0] - st[0]) / (N + W + K - 1), (en[-1] - en[0]) / (nw - 1)))
