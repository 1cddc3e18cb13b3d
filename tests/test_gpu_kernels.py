"""GPU parity of the individual multigrid kernels against the CPU oracle,
bit-exact (the kernels restate the reference arithmetic in IEEE fp64 without
FMA contraction)."""
import numpy as np
import pytest

from common import DH, FR, base_kw, make_pair, synth_psi

pytestmark = pytest.mark.gpu


def _model(gpu, N, nl):
    from msom_b200 import capi as G
    m = G.Model(G.make_params(**base_kw(N, nl)), gpu)
    m.set(G.PSI, synth_psi(N, nl))
    m.set_const()
    return m


def test_div_by_is_ieee_division(gpu):
    from msom_b200 import capi as G
    rng = np.random.default_rng(7)
    n = 1 << 22
    x = rng.standard_normal(n) * 10.0 ** rng.uniform(-8, 8, n)
    d = (4.0 + rng.uniform(0, 3, n)) * rng.choice([1.0, -1.0], n)
    d[: n // 4] = rng.standard_normal(n // 4) * 10.0 ** rng.uniform(-6, 6, n // 4)
    d[d == 0] = 1.0
    qf, qi = np.zeros(n), np.zeros(n)
    G.check(G.lib().msqg_test_div(gpu, x, d, qf, qi, n))
    assert np.array_equal(qi, x / d)          # device IEEE division == host
    assert np.array_equal(qf, qi)             # reciprocal-based exact division == IEEE


@pytest.mark.parametrize("nl,level,nsweeps", [(2, 3, 1), (2, 5, 4), (3, 6, 4), (4, 7, 4), (4, 6, 3), (4, 5, 2),
                                               (3, 7, 7), (2, 6, 13), (10, 5, 4), (4, 1, 4), (4, 2, 5), (2, 8, 4)])
def test_relax_matches_oracle(gpu, nl, level, nsweeps):
    from oracle import oracle as O
    from msom_b200 import capi as G
    N = max(1 << level, 32)
    m = _model(gpu, N, nl)
    n = 1 << level
    rng = np.random.default_rng(100 + level)
    a = rng.standard_normal((nl, n, n))
    b = rng.standard_normal((nl, n, n))
    kw = base_kw(N, nl)
    # stretching on this level = restriction of the uniform finest-level field (same summation order)
    s = np.zeros((nl - 1, n, n))
    for l in range(nl - 1):
        v = (FR[nl][l] / kw["Rom"]) ** 2
        for _ in range(int(np.log2(N)) - level):
            v = (((0.0 + v) + v) + v + v) / 4
        s[l] = v
    a_ref = a.copy()
    O.lib().orc_test_relax(nl, level, kw["L0"], np.array(DH[nl], dtype=np.float64), s, a_ref, b, nsweeps, 1, 1)
    a_gpu = a.copy()
    G.check(G.lib().msqg_test_relax(m.h, level, a_gpu, b, nsweeps))
    assert np.array_equal(a_gpu, a_ref), np.abs(a_gpu - a_ref).max()


@pytest.mark.parametrize("level,nsweeps,lam", [(4, 4, 0.0), (6, 4, -3.7), (7, 2, -50.0), (5, 8, -0.3)])
def test_relax_scalar_matches_oracle(gpu, level, nsweeps, lam):
    from oracle import oracle as O
    from msom_b200 import capi as G
    N = max(1 << level, 32)
    m = _model(gpu, N, 2)
    n = 1 << level
    rng = np.random.default_rng(5 + level)
    a = rng.standard_normal((1, n, n))
    b = rng.standard_normal((1, n, n))
    lamf = np.full((1, n, n), lam)
    a_ref = a.copy()
    O.lib().orc_test_relax_scalar(level, 80., lamf, a_ref, b, nsweeps)
    a_gpu = a.copy()
    G.check(G.lib().msqg_test_relax_scalar(m.h, level, lam, a_gpu, b, nsweeps))
    assert np.array_equal(a_gpu, a_ref), np.abs(a_gpu - a_ref).max()


@pytest.mark.parametrize("nl,N", [(2, 64), (3, 128), (4, 256), (10, 64)])
def test_residual_matches_oracle(gpu, nl, N):
    import ctypes as C
    from oracle import oracle as O
    from msom_b200 import capi as G
    m = _model(gpu, N, nl)
    level = int(np.log2(N))
    rng = np.random.default_rng(3)
    a = rng.standard_normal((nl, N, N))
    b = rng.standard_normal((nl, N, N))
    kw = base_kw(N, nl)
    s = np.zeros((nl - 1, N, N))
    for l in range(nl - 1):
        s[l] = (FR[nl][l] / kw["Rom"]) ** 2
    r_ref = np.zeros_like(a)
    mx_ref = O.lib().orc_test_residual(nl, level, kw["L0"], np.array(DH[nl], dtype=np.float64), s, a, b, r_ref)
    r_gpu = np.zeros_like(a)
    mx = C.c_double()
    G.check(G.lib().msqg_test_residual(m.h, a, b, r_gpu, C.byref(mx)))
    assert np.array_equal(r_gpu, r_ref)
    assert mx.value == mx_ref


@pytest.mark.parametrize("nl,level", [(2, 5), (4, 7)])
def test_restrict_prolong_match_oracle(gpu, nl, level):
    from oracle import oracle as O
    from msom_b200 import capi as G
    N = 1 << level
    m = _model(gpu, N, nl)
    rng = np.random.default_rng(11)
    fine = rng.standard_normal((nl, N, N))
    c_ref = np.zeros((nl, N // 2, N // 2))
    c_gpu = np.zeros_like(c_ref)
    O.lib().orc_test_restrict(nl, level, fine, c_ref)
    G.check(G.lib().msqg_test_restrict(m.h, level, fine, c_gpu))
    assert np.array_equal(c_gpu, c_ref)
    coarse = rng.standard_normal((nl, N // 2, N // 2))
    f_ref = np.zeros((nl, N, N))
    f_gpu = np.zeros_like(f_ref)
    O.lib().orc_test_prolong(nl, level, coarse, f_ref)
    G.check(G.lib().msqg_test_prolong(m.h, level, coarse, f_gpu))
    assert np.array_equal(f_gpu, f_ref)


@pytest.mark.parametrize("nl,level,nsweeps,cap", [(4, 8, 4, 3), (2, 7, 3, 1), (3, 8, 6, 5)])
def test_relax_in_column_panels_matches_oracle(gpu, nl, level, nsweeps, cap, monkeypatch):
    """levels wider than the device holds co-resident strips are swept in column panels (several launches,
    the global mailbox carries the boundary column): same bits as one launch and as the oracle"""
    monkeypatch.setenv("MSQG_RELAX_CAP", str(cap))
    test_relax_matches_oracle(gpu, nl, level, nsweeps)
