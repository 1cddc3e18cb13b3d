"""Doubly periodic boundaries, sbc = -1 (msqg/qg.h:80, :842-846: periodic(right); periodic(top), bc_type = -2).

On the GPU the periodic domain runs on the tile machinery of the multi-GPU path: every side of every tile is an internal
side whose halo comes from the tile across the seam -- the tile itself for px = 1 or py = 1 -- levels up to 32^2 are
swept by the single-CTA coarse kernel with wrap-around neighbours.  Red-black ordering only (a half-sweep does not
depend on where the seam is).  Bit-exact against the oracle, whose boundary() copies the opposite side into the ghost
ring on every level ([BASILISK] periodic box boundaries)."""
import numpy as np
import pytest

from common import base_kw

pytestmark = pytest.mark.gpu


def periodic_psi(N, nl, L0=80., seed=5):
    rng = np.random.default_rng(seed)
    x = (np.arange(N) + 0.5) * L0 / N
    X, Y = np.meshgrid(x, x)
    psi = np.zeros((nl, N, N))
    for l in range(nl):
        A = 1.0 / (l + 1)
        psi[l] = A * np.sin(2 * np.pi * X / L0) * np.sin(4 * np.pi * Y / L0) + 0.3 * A * np.cos(2 * np.pi * (X + 2 * Y) / L0)
        nz = rng.uniform(-1, 1, (N, N))
        psi[l] += 1e-3 * A * (nz - nz.mean())
    return psi


@pytest.mark.parametrize("N,nl,px,py,over", [(64, 2, 1, 1, {}), (128, 3, 1, 1, dict(Re=200.)), (128, 4, 2, 1, {}), (128, 3, 2, 2, {}),
                                              (256, 2, 4, 2, dict(Eks=0.001)), (32, 3, 1, 1, {}), (64, 5, 1, 2, {})])
def test_periodic_domain_matches_oracle(gpu, N, nl, px, py, over):
    from oracle import oracle as O
    from msom_b200 import capi as G
    from msom_b200.dist import Group
    kw = base_kw(N, nl, sbc=-1., **over)
    psi = periodic_psi(N, nl)
    mo = O.Model(O.make_params(**kw)); mo.set_smoother("rb")
    g = Group(G.make_params(**kw), px, py, 0, gpu, smoother="rb")
    mo.set(O.PSI, psi); g.set_global(G.PSI, psi)
    mo.set_const(); g.set_const()
    assert np.array_equal(g.get_global(G.Q), mo.get(O.Q))
    z = np.zeros_like(psi)
    mo.set(O.PSI, z); g.set_global(G.PSI, z)
    mo.invertq(); g.invertq()
    so, sg = mo.mgstats(), g.mgstats()
    assert (sg.i, sg.nrelax, sg.resb, sg.resa) == (so.i, so.nrelax, so.resb, so.resa)
    assert np.array_equal(g.get_global(G.PSI), mo.get(O.PSI))
    mo.set(O.PSI, psi); g.set_global(G.PSI, psi)
    for _ in range(4):
        assert g.step() == mo.step()
    assert g.total_cycles == mo.L.orc_total_cycles(mo.h)
    assert np.array_equal(g.get_global(G.Q), mo.get(O.Q))
    assert np.array_equal(g.get_global(G.PSI), mo.get(O.PSI))
    # periodic and closed-basin runs of the same initial state differ from the first step on
    kw2 = base_kw(N, nl, **over)
    mc = O.Model(O.make_params(**kw2)); mc.set_smoother("rb")
    mc.set(O.PSI, psi); mc.set_const()
    assert not np.array_equal(mc.get(O.Q), mo.get(O.Q))


def test_periodic_needs_the_tile_path(gpu):
    from msom_b200 import capi as G
    with pytest.raises(G.MsqgError):
        G.Model(G.make_params(**base_kw(64, 2, sbc=-1.)), gpu)
    from msom_b200.dist import Group
    with pytest.raises(G.MsqgError):
        Group(G.make_params(**base_kw(64, 2, sbc=-1.)), 1, 1, 0, gpu, smoother="lex")
    with pytest.raises(G.MsqgError):
        Group(G.make_params(**base_kw(64, 2, sbc=-1., upg=[0.1, 0.], vpg=[0., 0.])), 1, 1, 0, gpu, smoother="rb")
