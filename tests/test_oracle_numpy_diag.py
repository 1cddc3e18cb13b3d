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
