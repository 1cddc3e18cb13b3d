"""Domain decomposition (reference -D_MPI=1 semantics) on ONE GPU with the local exchange back-end:
all tiles in one process, halo exchange by device copies.  Checked bit for bit against the oracle's
emulation of the MPI block decomposition (orc_set_decomp)."""
import numpy as np
import pytest

from common import base_kw, synth_psi

pytestmark = pytest.mark.gpu


def _pair(N, nl, px, py, agg_n, gpu):
    from oracle import oracle as O
    from msom_b200 import capi as G
    from msom_b200.dist import Group
    kw = base_kw(N, nl)
    mo = O.Model(O.make_params(**kw))
    mo.L.orc_set_decomp(mo.h, px, py, agg_n)
    g = Group(G.make_params(**kw), px, py, agg_n, gpu)
    psi = synth_psi(N, nl)
    mo.set(O.PSI, psi)
    g.set_global(G.PSI, psi)
    mo.set_const(); g.set_const()
    return mo, g, psi


@pytest.mark.parametrize("N,nl,px,py,agg_n", [(128, 2, 2, 1, 32), (128, 3, 2, 2, 32), (256, 2, 4, 2, 64), (128, 4, 1, 2, 64)])
def test_decomposed_invertq_and_steps(gpu, N, nl, px, py, agg_n):
    from oracle import oracle as O
    from msom_b200 import capi as G
    mo, g, psi = _pair(N, nl, px, py, agg_n, gpu)
    assert np.array_equal(g.get_global(G.Q), mo.get(O.Q))
    z = np.zeros_like(psi)
    mo.set(O.PSI, z); g.set_global(G.PSI, z)
    mo.invertq(); g.invertq()
    so, sg = mo.mgstats(), g.mgstats()
    assert (sg.i, sg.nrelax, sg.resb, sg.resa) == (so.i, so.nrelax, so.resb, so.resa)
    assert np.array_equal(g.get_global(G.PSI), mo.get(O.PSI))
    for _ in range(3):
        assert g.step() == mo.step()
    assert g.total_cycles == mo.L.orc_total_cycles(mo.h)
    assert np.array_equal(g.get_global(G.Q), mo.get(O.Q))
    assert np.array_equal(g.get_global(G.PSI), mo.get(O.PSI))
    assert g.exchanges > 0


def test_decomposition_differs_from_serial_only_at_solver_tolerance(gpu):
    """block Gauss-Seidel is a different iterate from the serial sweep (poisson_layer.h:55-65) but the
    same solution to the solver tolerance"""
    from msom_b200 import capi as G
    from msom_b200.dist import Group
    N, nl = 128, 2
    kw = base_kw(N, nl)
    psi = synth_psi(N, nl)
    s = G.Model(G.make_params(**kw), gpu); s.set(G.PSI, psi); s.set_const()
    g = Group(G.make_params(**kw), 2, 2, 32, gpu); g.set_global(G.PSI, psi); g.set_const()
    for _ in range(2):
        s.step(); g.step()
    a, b = s.get(G.PSI), g.get_global(G.PSI)
    assert not np.array_equal(a, b)
    assert np.abs(a - b).max() < 1e-3 * np.abs(a).max()
