"""One timestep of the bench workload (4096^2 x nl=4 by default) on cuda:0 -- the target of the per-kernel ncu captures
in profiles/ (each capture filters one kernel name; see profiles/README.md for the commands)."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from common import base_kw, synth_psi
from msom_b200 import capi as G
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 4
m = G.Model(G.make_params(**base_kw(N, nl)))
m.set(G.PSI, synth_psi(N, nl))
m.set_const()
dt = m.step()
s = m.mgstats()
print("one step: dt=%g cycles=%d nrelax=%d launches=%d" % (dt, s.i, s.nrelax, m.launches))
