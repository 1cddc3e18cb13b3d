"""Energy diagnostics (msqg/qg_energy.h:28-242) and passive tracers (qg.h:573-588) restated in whole-array numpy from
the reference formulas and compared with the C oracle.  (CPU tests, no GPU.)"""
import numpy as np
import pytest

from common import base_kw, synth_psi
from oracle import oracle as O
from test_oracle_numpy_rhs import jac, lap, pad, sh, stretch


@pytest.mark.parametrize("ediag,over", [(0, dict(Re=200., Eks=0.001)), (1, dict(Re=50., flsrv=1, upg=[0.3, 0.1, 0.], vpg=[0.05, 0., -0.02]))])
def test_energy_tend_against_numpy(ediag, over):
    """energy_tend: advection_de (J1 = j(psi, q) with its stretching part, J2 = j(psi_pg, q), J3 = beta + j(psi, q_pg)),
    dissip_de, ekman_friction_de, every term times dt*(-psi*(1 - ediag) + ediag), and the running mean po_mft"""
    N, nl = 32, 3
    kw = base_kw(N, nl, ediag=ediag, **over)
    m = O.Model(O.make_params(**kw))
    psi = synth_psi(N, nl)
    m.set(O.PSI, psi); m.set_const()
    dt = 0.0123
    m.reset_energy()
    m.energy_tend(dt)
    L0, Rom, beta = kw["L0"], kw["Rom"], kw["beta"]
    D = L0 / N
    dh = np.array(kw["dh"], dtype=float)
    dhc = 0.5 * (dh[:-1] + dh[1:])
    idh0 = np.zeros(nl); idh1 = np.zeros(nl)
    idh1[:-1] = 1. / (dhc * dh[:-1]); idh0[1:] = 1. / (dhc * dh[1:])
    s = [(fr / Rom) ** 2 * np.ones((N, N)) for fr in kw["Fr"]]
    x = (np.arange(N) + 0.5) * D
    X, Y = np.meshgrid(x, x)
    upg, vpg = kw.get("upg", [0.] * nl), kw.get("vpg", [0.] * nl)
    pp = np.array([vpg[l] * X - upg[l] * Y for l in range(nl)])
    P = [pad(psi[l], -1) for l in range(nl)]; PP = [pad(pp[l], -1) for l in range(nl)]
    zeta = np.array([lap(P[l], D) for l in range(nl)]); Z = [pad(zeta[l], -1) for l in range(nl)]
    zpg = np.array([lap(PP[l], D) for l in range(nl)]) if kw.get("flsrv", 0) == 1 else np.zeros_like(psi)
    ZP = [pad(zpg[l], -1) for l in range(nl)]
    w = dt * (-psi * (1 - ediag) + ediag)
    j1 = np.zeros_like(psi); j2 = np.zeros_like(psi); j3 = np.zeros_like(psi)
    jd1 = [jac(P[l], P[l + 1], D) for l in range(nl - 1)]
    jd2 = [jac(PP[l], P[l + 1], D) for l in range(nl - 1)]
    jd3 = [jac(P[l], PP[l + 1], D) for l in range(nl - 1)]
    for l in range(nl):
        jc = jac(P[l], PP[l], D)
        a1 = jac(P[l], Z[l], D); a2 = jac(PP[l], Z[l], D); a3 = beta * (sh(P[l], -1, 0) - sh(P[l], 1, 0)) / (2 * D)
        if l > 0:        # ju_1 = -jd_1, ju_2 = -jd_3 (swap), ju_3 = -jd_2 (swap), qg_energy.h:88-90
            a1 = a1 + s[l - 1] * (-jd1[l - 1]) * idh0[l]
            a2 = a2 + s[l - 1] * (-jd3[l - 1] + jc) * idh0[l]
            a3 = a3 + s[l - 1] * (-jd2[l - 1] - jc) * idh0[l]
        if l < nl - 1:
            a1 = a1 + s[l] * jd1[l] * idh1[l]
            a2 = a2 + s[l] * (jd2[l] + jc) * idh1[l]
            a3 = a3 + s[l] * (jd3[l] - jc) * idh1[l]
        j1[l] = a1 * w[l]; j2[l] = a2 * w[l]; j3[l] = (a3 + jac(P[l], ZP[l], D)) * w[l]
    iRe, iRe4 = m.p.iRe, m.p.iRe4
    p4 = np.array([lap(Z[l], D) for l in range(nl)]); P4 = [pad(p4[l], -1) for l in range(nl)]
    vd = ((p4 + stretch(zeta, s, idh0, idh1)) * iRe + iRe4 * np.array([lap(P4[l], D) for l in range(nl)])
          + iRe4 * stretch(p4, s, idh0, idh1)) * w
    bf = np.zeros_like(psi)
    bf[0] -= kw.get("Eks", 0.) / (Rom * 2 * dh[0]) * zeta[0] * w[0]
    bf[-1] -= kw["Ekb"] / (Rom * 2 * dh[-1]) * zeta[-1] * w[-1]
    for name, fid, ref in (("j1", O.DE_J1, j1), ("j2", O.DE_J2, j2), ("j3", O.DE_J3, j3), ("vd", O.DE_VD, vd), ("bf", O.DE_BF, bf)):
        got = m.get(fid)
        scale = max(np.abs(ref).max(), 1e-300)
        assert np.abs(got - ref).max() <= 5e-12 * scale, (name, float(np.abs(got - ref).max() / scale))
    assert np.array_equal(m.get(O.PO_MFT), psi)     # running mean after one call: (0*0 + psi)/1


def test_tracer_tendency_against_numpy():
    """ptr_rhs + advance_qg on the tracer tail of `evolving`: dpdt = jacobian(po, ptr) + iPe laplacian(ptr)
    + ptr_ir (ptr_relax - ptr), tracers carry [BASILISK]'s default (zero-gradient) boundaries (bc_type + 1, qg.h:867-870)"""
    N, nl, nptr = 32, 2, 2
    kw = base_kw(N, nl, nptr=nptr, Pe=[30., 0.], ptr_r=[0., 0.05])
    m = O.Model(O.make_params(**kw))
    rng = np.random.default_rng(4)
    psi = synth_psi(N, nl)
    tr = rng.standard_normal((nl * nptr, N, N)); relax = rng.standard_normal((nl * nptr, N, N))
    m.set(O.PSI, psi); m.set(O.PTR, tr); m.set(O.PTR_RELAX, relax); m.set_const()
    m.update(kw["DT"])
    psi_i = m.get(O.PSI)                            # what invertq left
    D = kw["L0"] / N
    iPe = [1. / 30., 0.]; ir = [0., 1. / 0.05]
    got = m.get(O.DPTR)
    for l in range(nl):
        P = pad(psi_i[l], -1)
        for nt in range(nptr):
            f = l * nptr + nt
            T = pad(tr[f], +1)
            ref = jac(P, T, D) + iPe[nt] * lap(T, D) + ir[nt] * (relax[f] - tr[f])
            assert np.abs(got[f] - ref).max() <= 5e-12 * np.abs(ref).max(), f


def test_wavelet_filter_against_numpy():
    """The multiple-scale filter (qg.h:509-560) with [BASILISK]'s wavelet() / inverse_wavelet() and the sig_lev mask of
    set_const (qg.h:1057-1090): psi = invertq(q); w_l = psi_l - bilinear(restriction(psi)_l-1) on every level, times the
    high-pass mask, summed back up; q = comp_q(psi); qof = (q_before - q_after)/dtflt."""
    from test_oracle_numpy_mg import numpy_solve, prolong, restrict
    N, nl, afilt, dtflt = 64, 2, 4.0, 0.05
    kw = base_kw(N, nl, afilt=afilt, dtflt=dtflt)
    m = O.Model(O.make_params(**kw)); m.set_smoother("rb")
    psi0 = synth_psi(N, nl)
    m.set(O.PSI, psi0); m.set_const()
    q0 = m.get(O.Q)
    m.wavelet_filter(dtflt)
    L0 = kw["L0"]
    depth = int(np.log2(N))
    dh = np.array(kw["dh"], dtype=float)
    dhc = 0.5 * (dh[:-1] + dh[1:])
    idh0 = np.zeros(nl); idh1 = np.zeros(nl)
    idh1[:-1] = 1. / (dhc * dh[:-1]); idh0[1:] = 1. / (dhc * dh[1:])
    s = np.array([(fr / kw["Rom"]) ** 2 * np.ones((N, N)) for fr in kw["Fr"]])
    # sig_filt = min(afilt*Rd, Lfmax) with Rd = 1, restricted; low-pass mask per level, then 1 - mask
    sf = {depth: np.full((N, N), min(afilt * 1., kw.get("Lfmax", m.p.Lfmax)))}
    for l in range(depth - 1, -1, -1):
        sf[l] = restrict(sf[l + 1][None])[0]
    sl = {}
    for l in range(depth, -1, -1):
        n = 1 << l
        Dl = L0 / n
        low = np.where(sf[l] > 2 * Dl, 0., np.where(sf[l] > Dl, 1 - (sf[l] - Dl) / Dl, 1.))
        if l < depth:
            c = sl[l + 1]
            flag = c[0::2, 0::2] + c[1::2, 0::2] + c[0::2, 1::2] + c[1::2, 1::2]
            low = np.where(flag > 0, 1., low)
        sl[l] = low
    for l in sl:
        sl[l] = 1 - sl[l]
        assert np.abs(sl[l] - m.siglev(l)).max() <= 1e-15, l
    psi, st, _ = numpy_solve(psi0, q0, s, dh, L0)                      # invertq, warm start
    lev = {depth: psi}
    for l in range(depth - 1, -1, -1):
        lev[l] = restrict(lev[l + 1])
    w = {0: lev[0].copy()}
    for l in range(1, depth + 1):
        w[l] = lev[l] - prolong(lev[l - 1])
    for l in w:
        w[l] = w[l] * sl[l][None]
    rec = w[0]
    for l in range(1, depth + 1):
        rec = prolong(rec) + w[l]
    D = L0 / N
    q1 = np.array([lap(pad(rec[l], -1), D) for l in range(nl)]) + stretch(rec, list(s), idh0, idh1)
    assert np.abs(m.get(O.PSI) - rec).max() <= 1e-10 * np.abs(psi0).max()
    assert np.abs(m.get(O.Q) - q1).max() <= 1e-9 * np.abs(q0).max()
    assert np.abs(m.get(O.QOF) - (q0 - q1) / dtflt).max() <= 1e-8 * np.abs(q0).max() / dtflt
    assert 0 < np.abs(q1).max() < np.abs(q0).max()
