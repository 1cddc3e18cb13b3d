"""An INDEPENDENT restatement of one right-hand-side evaluation of the reference (update_qg after invertq,
msqg/qg.h:609-650), written in vectorised numpy straight from the reference formulas -- laplacian (qg.h:169),
comp_stretch (:202-246), the Arakawa jacobian macro (:252-262), beta_effect (:269), advection_pv (:287-380, default
build), dissip (:406-422), ekman_friction (:428-440), surface_forcing (:446-459), qforcing (:465-474),
bottom_topography (:480-488), comp_vel + [BASILISK] timestep (:275-283, :383-391) -- and compared with the C oracle.
Two restatements by different routes (scalar C loops in traversal order / whole-array numpy) agreeing to round-off
pin the oracle against transcription errors: sign conventions, argument order of the Jacobians, which stretching
coefficient multiplies which interface, the ghost rings every stencil reads.  (CPU test, no GPU.)"""
import numpy as np
import pytest

from common import base_kw, synth_psi
from oracle import oracle as O


def pad(a, sign):
    """one ghost ring: ghost = sign * interior on the four sides ([BASILISK] dirichlet(0): -1, default: +1), the y sides
    applied over the x ghosts, so a corner holds sign*sign*interior; sign = 0: periodic(right), periodic(top) -- the ring
    holds the opposite side, corners the diagonally opposite cell"""
    if sign == 0:
        return np.pad(a, 1, mode="wrap")
    p = np.zeros((a.shape[0] + 2, a.shape[1] + 2))
    p[1:-1, 1:-1] = a
    p[1:-1, 0] = sign * a[:, 0]; p[1:-1, -1] = sign * a[:, -1]
    p[0, :] = sign * p[1, :]; p[-1, :] = sign * p[-2, :]
    return p


def sh(p, dx, dy):
    """p[x + dx, y + dy] on the interior of a padded [y][x] array (the reference's po[dx, dy])"""
    n = p.shape[0] - 2
    return p[1 + dy:1 + dy + n, 1 + dx:1 + dx + n]


def lap(p, D):
    return (sh(p, 1, 0) + sh(p, -1, 0) + sh(p, 0, 1) + sh(p, 0, -1) - 4 * sh(p, 0, 0)) / D ** 2


def jac(P, Q, D):
    """the jacobian(po, qo) macro: -J(p, q), Arakawa 1966"""
    return ((sh(Q, 1, 0) - sh(Q, -1, 0)) * (sh(P, 0, 1) - sh(P, 0, -1))
            + (sh(Q, 0, -1) - sh(Q, 0, 1)) * (sh(P, 1, 0) - sh(P, -1, 0))
            + sh(Q, 1, 0) * (sh(P, 1, 1) - sh(P, 1, -1))
            - sh(Q, -1, 0) * (sh(P, -1, 1) - sh(P, -1, -1))
            - sh(Q, 0, 1) * (sh(P, 1, 1) - sh(P, -1, 1))
            + sh(Q, 0, -1) * (sh(P, 1, -1) - sh(P, -1, -1))
            + sh(P, 0, 1) * (sh(Q, 1, 1) - sh(Q, -1, 1))
            - sh(P, 0, -1) * (sh(Q, 1, -1) - sh(Q, -1, -1))
            - sh(P, 1, 0) * (sh(Q, 1, 1) - sh(Q, 1, -1))
            + sh(P, -1, 0) * (sh(Q, -1, 1) - sh(Q, -1, -1))) / (12. * D * D)


def stretch(f, s, idh0, idh1):
    """comp_stretch: Gamma(f)_l = s_{l-1} (f_{l-1} - f_l) idh0_l + s_l (f_{l+1} - f_l) idh1_l"""
    nl = f.shape[0]
    out = np.zeros_like(f)
    for l in range(nl):
        if l > 0:
            out[l] += s[l - 1] * (f[l - 1] - f[l]) * idh0[l]
        if l < nl - 1:
            out[l] += s[l] * (f[l + 1] - f[l]) * idh1[l]
    return out


def numpy_tendency(kw, psi, q_ev, qf, topo, iRe, iRe4, econs=False):
    """one evaluation of the tendencies of update_qg from the stream function psi (what invertq left) and, for the
    -DENERGY_CONSERV=1 build, the evolving PV q_ev; returns (dq, zeta, padded psi lists)"""
    nl, N = psi.shape[0], psi.shape[1]
    L0, Rom, beta = kw["L0"], kw["Rom"], kw["beta"]
    D = L0 / N
    dh = np.array(kw["dh"], dtype=float)
    dhc = 0.5 * (dh[:-1] + dh[1:])
    idh0 = np.zeros(nl); idh1 = np.zeros(nl)
    idh1[:-1] = 1. / (dhc * dh[:-1]); idh0[1:] = 1. / (dhc * dh[1:])
    s = [(fr / Rom) ** 2 * np.ones((N, N)) for fr in kw["Fr"]]
    x = (np.arange(N) + 0.5) * D
    X, Y = np.meshgrid(x, x)                       # [y][x]
    upg, vpg = kw.get("upg", [0.] * nl), kw.get("vpg", [0.] * nl)
    pp = np.array([vpg[l] * X - upg[l] * Y for l in range(nl)])
    P = [pad(psi[l], -1) for l in range(nl)]
    PP = [pad(pp[l], -1) for l in range(nl)]
    zeta = np.array([lap(P[l], D) for l in range(nl)])
    Z = [pad(zeta[l], -1) for l in range(nl)]
    zpg = np.array([lap(PP[l], D) for l in range(nl)]) if kw.get("flsrv", 0) == 1 else np.zeros_like(psi)
    ZP = [pad(zpg[l], -1) for l in range(nl)]
    QE = [pad(q_ev[l], -1) for l in range(nl)]
    # advection_pv (_LS_RV = 1); the -DENERGY_CONSERV=1 build (qg.h:310-312, :338-340, :366-367) advects the full PV and
    # keeps only the large-scale parts of the stretching Jacobians
    out = np.zeros_like(psi)
    if econs:
        jd = [jac(PP[l], P[l + 1], D) + jac(P[l], PP[l + 1], D) for l in range(nl - 1)]
    else:
        jd = [jac(P[l], P[l + 1], D) + jac(PP[l], P[l + 1], D) + jac(P[l], PP[l + 1], D) for l in range(nl - 1)]
    for l in range(nl):
        t = jac(P[l], QE[l] if econs else Z[l], D) + jac(PP[l], Z[l], D) + beta * (sh(P[l], -1, 0) - sh(P[l], 1, 0)) / (2 * D)
        if l > 0:
            t = t + s[l - 1] * (-jd[l - 1]) * idh0[l]
        if l < nl - 1:
            t = t + s[l] * jd[l] * idh1[l]
        out[l] = t + jac(P[l], ZP[l], D)
    # dissip
    tmp = np.array([lap(Z[l], D) for l in range(nl)])
    T = [pad(tmp[l], -1) for l in range(nl)]
    out += iRe * stretch(zeta, s, idh0, idh1) + iRe * tmp + iRe4 * stretch(tmp, s, idh0, idh1)
    out += iRe4 * np.array([lap(T[l], D) for l in range(nl)])
    # ekman_friction, surface_forcing, qforcing, bottom_topography
    out[0] -= kw.get("Eks", 0.) / (Rom * 2 * dh[0]) * zeta[0]
    out[-1] -= kw["Ekb"] / (Rom * 2 * dh[-1]) * zeta[-1]
    out[0] -= kw["tau0"] / (Rom * dh[0]) * np.sin(2 * np.pi * Y / L0) * np.sin(np.pi * Y / L0)
    out += qf
    out[-1] += jac(P[-1], pad(topo[0], +1), D) / (Rom * dh[-1])
    return out, zeta, P, PP


@pytest.mark.parametrize("N,nl,over", [(32, 3, dict(Re=200., Eks=0.001, flsrv=1, upg=[0.3, 0.1, 0.], vpg=[0.05, 0., -0.02])),
                                        (64, 2, {}), (32, 4, dict(Re=50.))])
def test_rhs_against_numpy_restatement(N, nl, over):
    kw = base_kw(N, nl, **over)
    m = O.Model(O.make_params(**kw))
    rng = np.random.default_rng(8)
    m.set(O.PSI, synth_psi(N, nl))
    qf = 1e-3 * rng.standard_normal((nl, N, N))
    topo = 0.1 * rng.standard_normal((1, N, N))
    m.set(O.QFORC, qf); m.set(O.TOPO, topo); m.L.orc_set_flag_topo(m.h, 1)
    m.set_const()
    D = kw["L0"] / N
    for econs in (False, True):                    # the default build, then the -DENERGY_CONSERV=1 build
        m.set_energy_conserv(econs)
        dt_oracle = m.update(kw["DT"])
        psi, dq = m.get(O.PSI), m.get(O.DQ)        # psi is what invertq left: the stream function the RHS was built from
        out, zeta, P, PP = numpy_tendency(kw, psi, m.get(O.Q), qf, topo, m.p.iRe, m.p.iRe4, econs)
        assert np.abs(zeta - m.get(O.ZETA)).max() <= 1e-12 * np.abs(zeta).max()
        scale = np.abs(dq).max()
        assert scale > 0 and np.abs(out - dq).max() <= 2e-12 * scale, (econs, float(np.abs(out - dq).max() / scale))
        if not econs:
            dq_default, dt_first = dq, dt_oracle
    assert np.abs(dq - dq_default).max() > 0       # the two builds are different discretisations

    # comp_vel + timestep() of the FIRST update: dt = CFL * min Delta/|u| over the faces of psi and psi_pg of every layer,
    # chained through the static `previous` (0 at the first call: the first dt is ramped to 0.1/1.1 of its value)
    m2 = O.Model(O.make_params(**kw))
    m2.set(O.PSI, synth_psi(N, nl)); m2.set(O.QFORC, qf); m2.set(O.TOPO, topo); m2.L.orc_set_flag_topo(m2.h, 1)
    m2.set_const()
    assert m2.update(kw["DT"]) == dt_first
    psi = m2.get(O.PSI)
    _, _, P, PP = numpy_tendency(kw, psi, m2.get(O.Q), qf, topo, m2.p.iRe, m2.p.iRe4)

    def umax(Pl):
        n = Pl.shape[0] - 2
        ux = -0.25 * (Pl[2:n + 2, 1:n + 2] - Pl[0:n, 1:n + 2] + Pl[2:n + 2, 0:n + 1] - Pl[0:n, 0:n + 1]) / D   # faces i = 0..n
        uy = 0.25 * (Pl[1:n + 2, 2:n + 2] - Pl[1:n + 2, 0:n] + Pl[0:n + 1, 2:n + 2] - Pl[0:n + 1, 0:n]) / D    # faces j = 0..n
        return max(np.abs(ux).max(), np.abs(uy).max())

    CFL, prev, dtmax = kw["CFL"], 0., kw["DT"]
    for l in range(nl):
        for Pl in (P[l], PP[l]):
            u = umax(Pl)
            dtmax = dtmax / CFL
            if u != 0 and D / u < dtmax:
                dtmax = D / u
            dtmax *= CFL
            if dtmax > prev:
                dtmax = (prev + 0.1 * dtmax) / 1.1
            prev = dtmax
    assert dt_first == pytest.approx(dtmax, rel=1e-12)


def test_multigrid_operators_against_numpy_restatement():
    """residual_layer (poisson_layer.h:157-258: res = b - laplacian(a) - Gamma(a), the face-gradient form written out),
    [BASILISK] restriction (mean of the four children) and bilinear prolongation ((9, 3, 3, 1)/16 on the coarse ring with
    its dirichlet ghosts), whole-array numpy against the oracle's loops."""
    L = O.lib()
    rng = np.random.default_rng(21)
    nl, level, L0 = 3, 5, 80.
    n = 1 << level
    D = L0 / n
    dh = np.array([0.06, 0.14, 0.8])
    dhc = 0.5 * (dh[:-1] + dh[1:])
    idh0 = np.zeros(nl); idh1 = np.zeros(nl)
    idh1[:-1] = 1. / (dhc * dh[:-1]); idh0[1:] = 1. / (dhc * dh[1:])
    s = np.abs(rng.standard_normal((nl - 1, n, n))) * 10 + 5
    a = rng.standard_normal((nl, n, n)); b = rng.standard_normal((nl, n, n))
    res = np.zeros_like(a)
    mx = L.orc_test_residual(nl, level, L0, dh, np.ascontiguousarray(s), a, b, res)
    A = [pad(a[l], -1) for l in range(nl)]
    ref = b - np.array([lap(A[l], D) for l in range(nl)]) - stretch(a, list(s), idh0, idh1)
    assert np.abs(ref - res).max() <= 1e-12 * np.abs(res).max()
    assert mx == pytest.approx(np.abs(res).max(), rel=1e-15)
    # restriction: mean of the 2 x 2 children
    coarse = np.zeros((nl, n // 2, n // 2))
    L.orc_test_restrict(nl, level, a, coarse)
    ref = 0.25 * (a[:, 0::2, 0::2] + a[:, 1::2, 0::2] + a[:, 0::2, 1::2] + a[:, 1::2, 1::2])
    assert np.abs(ref - coarse).max() <= 1e-15 * np.abs(a).max()
    # bilinear prolongation of the coarse field with homogeneous dirichlet ghosts
    fine = np.zeros_like(a)
    L.orc_test_prolong(nl, level, coarse, fine)
    for l in range(nl):
        C = pad(coarse[l], -1)
        nc = n // 2
        ref = np.zeros((n, n))
        for py in (0, 1):
            for px in (0, 1):
                cx, cy = (1 if px else -1), (1 if py else -1)      # child.x, child.y
                ref[py::2, px::2] = (9. * sh(C, 0, 0) + 3. * (sh(C, cx, 0) + sh(C, 0, cy)) + sh(C, cx, cy)) / 16.
        assert np.abs(ref - fine[l]).max() <= 1e-14 * np.abs(coarse).max()
