"""MODE_PV_INVERT 1 (msqg/qg.h:116-157, eigmode.h:65-308) restated independently in numpy and run beside the C oracle:
the vertical modes from numpy.linalg.eig with the reference's normalisation (Flierl 1978: sum_k dh_k vr_km^2 = 1,
positive at the surface, left vectors scaled to vl.vr = 1, eigenvalues ascending, iBu = -lambda with iBu[0] = 0), the
projection q -> modes, one scalar Helmholtz multigrid solve per mode ([BASILISK] poisson(): relax() / residual() with
lambda = iBu, red-black order) and the projection back.  (CPU test, no GPU.)"""
import numpy as np
import pytest

from common import DH, FR, base_kw, synth_psi
from oracle import oracle as O
from test_oracle_numpy_mg import prolong, restrict
from test_oracle_numpy_rhs import lap, pad


def numpy_modes(dh, fr, Ro):
    nl = len(dh)
    dhc = 0.5 * (dh[:-1] + dh[1:])
    A = np.zeros((nl, nl))
    for l in range(nl):                       # eigmode.h:86-106
        if l < nl - 1:
            A[l, l + 1] = -(fr[l] / Ro) ** 2 / (dhc[l] * dh[l])
        if l > 0:
            A[l, l - 1] = -(fr[l - 1] / Ro) ** 2 / (dhc[l - 1] * dh[l])
        A[l, l] = -A[l].sum()
    w, V = np.linalg.eig(A)
    order = np.argsort(w.real)
    w, V = w.real[order], V.real[:, order]
    for mm in range(nl):
        V[:, mm] *= np.sign(V[0, mm]) * np.sqrt(1. / (dh * V[:, mm] ** 2).sum())
    ib = -w
    ib[0] = 0.
    return np.linalg.inv(V), V, ib            # cl2m (rows = left eigenvectors), cm2l (columns = modes), iBu


def test_vertical_modes_against_numpy_eig():
    if O._lapack_path() is None:
        pytest.skip("no LAPACK dgeev available")
    for nl in (2, 3, 4, 10):
        dh = np.array(DH[nl]); fr = np.array(FR[nl] + [0.]); Ro = 0.025
        cl, cm, ib = np.zeros(nl * nl), np.zeros(nl * nl), np.zeros(nl)
        assert O.lib().orc_eigmod_column(nl, dh, fr, Ro, cl, cm, ib) == 0
        L2M, M2L, IB = numpy_modes(dh, fr, Ro)
        assert np.allclose(cm.reshape(nl, nl), M2L, rtol=1e-9, atol=1e-12)
        assert np.allclose(cl.reshape(nl, nl), L2M, rtol=1e-8, atol=1e-11)
        assert np.allclose(ib[1:], IB[1:], rtol=1e-10) and ib[0] == 0.


def scalar_sweeps(a, b, lam, D, nsweeps):
    """[BASILISK] poisson.h relax() (in-tree copy mspg/elliptic.h:294-301), alpha = 1: a = (-D^2 b + sum of the four
    neighbours) / (4 - lam D^2), red cells then black cells, homogeneous dirichlet ghosts"""
    n = a.shape[0]
    yy, xx = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    a = a.copy()
    for _ in range(nsweeps):
        for colour in (0, 1):
            P = pad(a, -1)
            new = (-D * D * b + P[1:-1, 2:] + P[1:-1, :-2] + P[2:, 1:-1] + P[:-2, 1:-1]) / (4. - lam * D * D)
            mask = ((xx + yy) & 1) == colour
            a[mask] = new[mask]
    return a


def scalar_solve(a, b, lam, L0, tol=1e-3):
    N = a.shape[0]
    depth = int(np.log2(N))
    resid = lambda x: b - lam * x - lap(pad(x, -1), L0 / N)      # residual(): b - lambda a - laplacian(a)
    a = a.copy()
    res = resid(a); resb = float(np.abs(res).max())
    st = dict(i=0, nrelax=4, resb=resb, resa=resb)
    while st["i"] < 100 and (st["i"] < 1 or st["resa"] > tol):
        r = {depth: res}
        for l in range(depth - 1, 0, -1):
            r[l] = restrict(r[l + 1][None])[0]
        da = None
        for l in range(1, depth + 1):
            n = 1 << l
            da = np.zeros((n, n)) if l == 1 else prolong(da[None])[0]
            da = scalar_sweeps(da, r[l], lam, L0 / n, st["nrelax"])   # a uniform lambda restricts to itself
        a = a + da
        res = resid(a); resa = float(np.abs(res).max())
        st["resa"] = resa
        if resa > tol:
            if resb / resa < 1.2 and st["nrelax"] < 100:
                st["nrelax"] += 1
            elif resb / resa > 10 and st["nrelax"] > 2:
                st["nrelax"] -= 1
        resb = resa
        st["i"] += 1
    return a, st


@pytest.mark.parametrize("N,nl", [(32, 3), (64, 2)])
def test_modal_inversion_against_numpy(N, nl):
    if O._lapack_path() is None:
        pytest.skip("no LAPACK dgeev available")
    kw = base_kw(N, nl, mode_pv_invert=1)
    m = O.Model(O.make_params(**kw)); m.set_smoother("rb")
    psi0 = synth_psi(N, nl)
    m.set(O.PSI, psi0); m.set_const()
    q = m.get(O.Q)
    L2M, M2L, IB = numpy_modes(np.array(kw["dh"], dtype=float), np.array(list(kw["Fr"]) + [0.]), kw["Rom"])
    m.set(O.PSI, np.zeros_like(psi0)); m.set(O.PM, np.zeros_like(psi0))     # cold start of every mode
    m.invertq()
    qm = np.einsum("ml,lyx->myx", L2M, q)                         # qg.h:118-131
    pm = np.zeros_like(qm)
    for mode in range(nl):                                        # qg.h:136-141: poisson(pm, qm, lambda = iBu, tolerance = 1e-3)
        pm[mode], st = scalar_solve(np.zeros((N, N)), qm[mode], IB[mode], kw["L0"])
        so = m.mgstats(mode)
        assert (so.i, so.nrelax) == (st["i"], st["nrelax"]), (mode, so.i, so.nrelax, st)
        assert so.resa == pytest.approx(st["resa"], rel=1e-6)
    psi = np.einsum("lm,myx->lyx", M2L, pm)                       # qg.h:144-157
    got = m.get(O.PSI)
    assert np.abs(got - psi).max() <= 1e-8 * np.abs(psi).max()
