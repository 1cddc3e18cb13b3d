"""Oracle parity at the REAL BASELINE sizes (not only size-independent properties): the CUDA path against the serial
CPU oracle, bit for bit, with equal multigrid cycle counts.

  * config 2: 1024^2 x 3, 3 steps, both sweep orderings;
  * config 3: 2048^2 x 10 with the vertical-mode inversion (MODE_PV_INVERT 1), 1 step;
  * the metric shape: 4096^2 x 4, 1 step, reference sweep order (serial oracle) and red-black;
  * an 8192-wide level through the column-panel path of the reference-order relax kernel (more strips than the device
    holds co-resident: consecutive launches hand the boundary column over), panels occurring naturally.
The oracle needs seconds to a minute per case on one core; the whole file stays within a few minutes."""
import numpy as np
import pytest

from common import base_kw, make_pair, synth_psi

pytestmark = pytest.mark.gpu


def _steps_equal(N, nl, nsteps, smoother, **over):
    from oracle import oracle as O
    from msom_b200 import capi as G
    mo, mg, _ = make_pair(N, nl, smoother=smoother, **over)
    mo.set_const(); mg.set_const()
    for _ in range(nsteps):
        assert mg.step() == mo.step()
    assert mg.total_cycles == mo.L.orc_total_cycles(mo.h)
    so, sg = mo.mgstats(), mg.mgstats()
    assert (sg.i, sg.nrelax, sg.resb, sg.resa) == (so.i, so.nrelax, so.resb, so.resa)
    for fg, fo in ((G.PSI, O.PSI), (G.Q, O.Q)):
        a, b = mg.get(fg), mo.get(fo)
        assert np.array_equal(a, b), np.abs(a - b).max()
    mo.close(); mg.close()


@pytest.mark.parametrize("smoother", ["lex", "rb"])
def test_config2_1024x3_three_steps(gpu, smoother):
    _steps_equal(1024, 3, 3, smoother)


@pytest.mark.parametrize("smoother", ["lex", "rb"])
def test_config3_2048x10_modal_one_step(gpu, smoother):
    _steps_equal(2048, 10, 1, smoother, mode_pv_invert=1)


@pytest.mark.parametrize("smoother", ["lex", "rb"])
def test_metric_shape_4096x4_one_step(gpu, smoother):
    _steps_equal(4096, 4, 1, smoother)


def test_relax_8192_wide_level_through_column_panels(gpu):
    """reference-order relax kernel on a level wider than the device holds co-resident strips (nx > ~4700 at nl = 4)"""
    from oracle import oracle as O
    from msom_b200 import capi as G
    N, nl, level, nsweeps = 8192, 4, 13, 2
    kw = base_kw(N, nl)
    m = G.Model(G.make_params(**kw), gpu)
    m.set(G.PSI, np.zeros((nl, N, N)))
    m.set_const()
    rng = np.random.default_rng(13)
    a = rng.standard_normal((nl, N, N))
    b = rng.standard_normal((nl, N, N))
    s = np.zeros((nl - 1, N, N))
    for l in range(nl - 1):
        s[l] = (kw["Fr"][l] / kw["Rom"]) ** 2
    l0 = m.launches
    a_gpu = a.copy()
    G.check(G.lib().msqg_test_relax(m.h, level, a_gpu, b, nsweeps))
    assert m.launches - l0 >= 2          # more than one panel launch (+ mailbox arming)
    m.close()
    O.lib().orc_test_relax(nl, level, kw["L0"], np.array(kw["dh"], dtype=np.float64), s, a, b, nsweeps, 1, 1)
    assert np.array_equal(a_gpu, a), np.abs(a_gpu - a).max()
