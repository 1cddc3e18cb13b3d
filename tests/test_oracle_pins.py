"""Known-answer pins of the CPU oracle: the reference ships
no golden vectors and cannot be built here, so the oracle is pinned by identities
that follow from the reference's own formulas, plus regression fixtures generated
by the oracle itself (tests/golden, made by tests/golden/make_golden.py)."""
import ctypes as C
import os

import numpy as np
import pytest

from common import DH, FR, base_kw, rel_l2, synth_psi
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _model(N, nl, **over):
    m = O.Model(O.make_params(**base_kw(N, nl, **over)))
    m.set(O.PSI, synth_psi(N, nl))
    m.set_const()
    return m


def test_read_params_double_gyre_fixture(tmp_path):
    """read_params on the reference's only shipped configuration (values of
    msqg/test/params.double_gyre.in) incl. the derived values of qg.h:739-746."""
    f = tmp_path / "params.in"
    f.write_text("#!sh\n# Double gyre configuration of Verron 1992\n\nN  = 256\nnl = 3\nL0 = 80\n\nRom   = 0.025\n"
                 "Ekb   = 0.002\ntau0  = 0.0001\nRe4   = 1563\nbeta  = 0.5\nFr = [0.0023669,0.0076173]\n"
                 "dh = [0.06,0.14,0.8]\n\nDT    = 5.e-2\ntend  = 500.\ndtout = 1.\nCFL   = 0.6\n")
    p = O.Params()
    O.lib().orc_default_params(p)
    assert O.lib().orc_read_params(str(f).encode(), p) == 0
    assert (p.N, p.nl, p.L0, p.Rom, p.Ekb, p.tau0, p.Re4, p.beta) == (256, 3, 80., 0.025, 0.002, 1e-4, 1563., 0.5)
    assert list(p.Fr)[:2] == [0.0023669, 0.0076173] and list(p.dh)[:3] == [0.06, 0.14, 0.8]
    assert (p.tend, p.dtout, p.CFL) == (500., 1., 0.6)
    assert p.iRe == 0. and p.iRe4 == -1 / 1563.
    # DT = 0.5*min(DT, Delta^4*Re4/32) -- the 0.5 applies even when min picks DT
    d2 = (80. / 256) * (80. / 256)
    assert p.DT == 0.5 * min(5e-2, d2 * d2 * 1563. / 32.)
    assert O.lib().orc_read_params(b"/nonexistent/params.in", p) == -1


def test_arakawa_invariants():
    """Arakawa (1966) Jacobian (qg.h:252-262): for fields supported away from the walls,
    sum J = sum psi*J = sum zeta*J = 0 to round-off.  Checked through update_qg with every
    other tendency switched off (beta = Ekb = tau0 = Re4 = 0, negligible stretching)."""
    N, nl = 64, 2
    m = O.Model(O.make_params(**base_kw(N, nl, beta=0., Ekb=0., tau0=0., Re4=0., Fr=[1e-30], DT=1e-3)))
    rng = np.random.default_rng(0)
    psi = np.zeros((nl, N, N))
    psi[:, 6:-6, 6:-6] = rng.standard_normal((nl, N - 12, N - 12))
    for _ in range(3):  # smooth a little, support stays >= 3 cells away from the walls
        psi[:, 1:-1, 1:-1] = 0.25 * (psi[:, :-2, 1:-1] + psi[:, 2:, 1:-1] + psi[:, 1:-1, :-2] + psi[:, 1:-1, 2:])
    m.set(O.PSI, psi)
    m.set_const()          # q = comp_q(psi): invertq inside update_qg then returns psi to round-off
    m.update(1e-3)
    p, z, dq = m.get(O.PSI), m.get(O.ZETA), m.get(O.DQ)
    assert rel_l2(p, psi) < 1e-9
    for l in range(nl):
        assert np.abs(dq[l]).max() > 0
        assert abs(dq[l].sum()) < 1e-9 * np.abs(dq[l]).sum()
        assert abs((z[l] * dq[l]).sum()) < 1e-9 * np.abs(z[l] * dq[l]).sum()
        assert abs((p[l] * dq[l]).sum()) < 1e-9 * np.abs(p[l] * dq[l]).sum()


def test_thomas_line_solve_matches_dense():
    """One relax_layer sweep on a single interior cell == dense solve of the nl x nl system
    (poisson_layer.h:80-146)."""
    nl, level = 4, 1
    n = 2
    kw = base_kw(32, nl)
    dh = np.array(DH[nl])
    s = np.array([(f / kw["Rom"]) ** 2 for f in FR[nl]])
    rng = np.random.default_rng(1)
    b = rng.standard_normal((nl, n, n))
    a = np.zeros((nl, n, n))
    sf = np.stack([np.full((n, n), v) for v in s])
    O.lib().orc_test_relax(nl, level, kw["L0"], dh, sf, a, b, 1, 1, 1)
    # first cell in sweep order (x=0,y=0): all neighbours are zero (ghost = -0, others not yet updated = 0)
    D = kw["L0"] / n
    dhc = 0.5 * (dh[:-1] + dh[1:])
    A = np.zeros((nl, nl))
    for l in range(nl):
        t0 = -D * D * s[l - 1] / (dhc[l - 1] * dh[l]) if l > 0 else 0.
        t2 = -D * D * s[l] / (dhc[l] * dh[l]) if l < nl - 1 else 0.
        A[l, l] = 4 - t0 - t2
        if l > 0:
            A[l, l - 1] = t0
        if l < nl - 1:
            A[l, l + 1] = t2
    ref = np.linalg.solve(A, -D * D * b[:, 0, 0])
    assert np.allclose(a[:, 0, 0], ref, rtol=1e-12, atol=0)


def test_eigenmode_identities():
    """eigmode.h:213-266: cl2m*cm2l = I, sum_k dh_k vr_km^2 = 1, surface-positive, iBu[0] = 0,
    and the 2-layer analytic deformation radius."""
    for nl in (2, 3, 4, 10):
        dh = np.array(DH[nl]); fr = np.array(FR[nl] + [0.]); Ro = 0.025
        cl, cm, ib = np.zeros(nl * nl), np.zeros(nl * nl), np.zeros(nl)
        rc = O.lib().orc_eigmod_column(nl, dh, fr, Ro, cl, cm, ib)
        if rc == -2:
            pytest.skip("no LAPACK dgeev available")
        assert rc == 0
        L2M, M2L = cl.reshape(nl, nl).T, cm.reshape(nl, nl)   # cl2m[k*nl+m] = vl[m][k]; cm2l[k*nl+m] = vr[k][m]
        # reference layout: qm[m] = sum_l cl2m[m*nl+l] q_l ; po[l] = sum_m cm2l[l*nl+m] pm
        A = cl.reshape(nl, nl) @ cm.reshape(nl, nl)
        assert np.allclose(A, np.eye(nl), atol=1e-10)
        vr = cm.reshape(nl, nl)
        assert np.allclose((dh[:, None] * vr ** 2).sum(0), 1., rtol=1e-12)
        assert (vr[0] > 0).all()
        assert ib[0] == 0. and (ib[1:] < 0).all() and np.all(np.diff(ib) <= 0)
    dh = np.array(DH[2]); F = FR[2][0]; Ro = 0.025
    cl, cm, ib = np.zeros(4), np.zeros(4), np.zeros(2)
    O.lib().orc_eigmod_column(2, dh, np.array([F, 0.]), Ro, cl, cm, ib)
    dhc = 0.5 * (dh[0] + dh[1])
    lam = (F / Ro) ** 2 / dhc * (1 / dh[0] + 1 / dh[1])
    assert np.isclose(-ib[1], lam, rtol=1e-12)


def test_q2p_p2q_round_trip():
    """comp_q(invertq(q)) == q to the solver tolerance (max-norm residual <= 1e-3, qg.h:159)."""
    N, nl = 64, 3
    m = _model(N, nl)
    q = m.get(O.Q)
    p = np.zeros_like(q); q2 = np.zeros_like(q)
    m.L.orc_pyq2p(m.h, p, q)
    s = m.mgstats()
    assert s.resa <= 1e-3 and s.i >= 1
    m.L.orc_pyp2q(m.h, p, q2)
    assert np.abs(q2 - q).max() <= 1e-3 * 1.0000001
    assert np.abs(q2 - q).max() == pytest.approx(s.resa, rel=1e-6)


def test_bas_format_round_trip(tmp_path):
    """.bas records (auxiliar_input.h:128-141) as the readers of msqg/scripts/read_data.py:22-46 see them."""
    N, nl, L0 = 32, 3, 80.
    v = synth_psi(N, nl)
    f = str(tmp_path / "po.bas")
    assert O.lib().orc_write_bas(f.encode(), nl, N, L0, v) == 0
    raw = np.fromfile(f, "f4").reshape(nl, N + 1, N + 1)
    assert (raw[:, 0, 0] == N).all()
    assert np.allclose(raw[0, 0, 1:], (np.arange(N) + 0.5) * L0 / N, rtol=1e-6)
    a = raw.transpose(0, 2, 1)[:, 1:, 1:]
    assert np.array_equal(a, v.astype("f4"))
    back = np.zeros_like(v)
    assert O.lib().orc_read_bas(f.encode(), nl, N, L0, back) == 0
    assert np.array_equal(back, v.astype("f4").astype("f8"))


def test_timestep_ramp_and_ke():
    """timestep() static state ([BASILISK] timestep.h): first dt = DT/11, geometric ramp toward DT."""
    m = _model(64, 2)
    DT = m.p.DT
    dts = [m.step() for _ in range(4)]
    assert dts[0] == pytest.approx(DT / 11, rel=1e-14)
    assert dts[1] == pytest.approx((dts[0] + 0.1 * DT) / 1.1, rel=1e-14)
    assert all(b > a for a, b in zip(dts, dts[1:]))
    assert m.ke1() > 0


def test_golden_regression():
    """The oracle reproduces its committed outputs (guards the oracle against accidental edits)."""
    g = np.load(os.path.join(HERE, "golden", "oracle_32x2_3steps.npz"))
    m = _model(32, 2)
    dts = [m.step() for _ in range(3)]
    assert np.array_equal(np.array(dts), g["dts"])
    assert np.array_equal(m.get(O.PSI), g["psi"]) and np.array_equal(m.get(O.Q), g["q"])
    g = np.load(os.path.join(HERE, "golden", "oracle_32x3_modal_2steps.npz"))
    if O._lapack_path() is None:
        pytest.skip("no LAPACK for eigmod")
    m = _model(32, 3, mode_pv_invert=1)
    dts = [m.step() for _ in range(2)]
    assert rel_l2(m.get(O.PSI), g["psi"]) < 1e-11   # LAPACK builds may differ in the last bits


def test_golden_regression_round2_variants():
    """periodic domain (red-black), ENERGY_CONSERV, per-column vertical modes: the oracle reproduces the fixtures
    tests/golden/make_golden.py wrote for them (regression only, like every fixture here)."""
    from common import periodic_psi
    g = np.load(os.path.join(HERE, "golden", "oracle_rb_periodic_32x3_3steps.npz"))
    m = O.Model(O.make_params(**base_kw(32, 3, sbc=-1.))); m.set_smoother("rb")
    m.set(O.PSI, periodic_psi(32, 3)); m.set_const()
    dts = [m.step() for _ in range(3)]
    assert np.array_equal(np.array(dts), g["dts"])
    assert np.array_equal(m.get(O.PSI), g["psi"]) and np.array_equal(m.get(O.Q), g["q"])
    g = np.load(os.path.join(HERE, "golden", "oracle_econs_32x2_3steps.npz"))
    m = _model(32, 2); m.set_energy_conserv(1)
    dts = [m.step() for _ in range(3)]
    assert np.array_equal(np.array(dts), g["dts"])
    assert np.array_equal(m.get(O.PSI), g["psi"]) and np.array_equal(m.get(O.Q), g["q"])
    if O._lapack_path() is None:
        pytest.skip("no LAPACK for eigmod")
    g = np.load(os.path.join(HERE, "golden", "oracle_32x3_modal_varRo_2steps.npz"))
    m = _model(32, 3, mode_pv_invert=1, varRo=1)
    dts = [m.step() for _ in range(2)]
    assert rel_l2(m.get(O.PSI), g["psi"]) < 1e-11   # LAPACK builds may differ in the last bits


def test_energy_budget_identities():
    """qg_energy.h: with the weight -psi (ediag = 0) the advective terms integrate to zero -- the Arakawa
    Jacobian conserves energy and the stretching Jacobians of neighbouring layers cancel in the
    thickness-weighted sum (psi supported away from the walls); viscous and drag terms dissipate."""
    N, nl = 64, 3
    kw = base_kw(N, nl, beta=0., tau0=0., Re=500., Eks=0.002, ediag=0)
    m = O.Model(O.make_params(**kw))
    rng = np.random.default_rng(1)
    psi = np.zeros((nl, N, N))
    psi[:, 8:-8, 8:-8] = rng.standard_normal((nl, N - 16, N - 16))
    for _ in range(4):
        psi[:, 1:-1, 1:-1] = 0.25 * (psi[:, :-2, 1:-1] + psi[:, 2:, 1:-1] + psi[:, 1:-1, :-2] + psi[:, 1:-1, 2:])
    m.set(O.PSI, psi); m.set_const()
    m.energy_tend(1.0)
    dh = np.array(kw["dh"])[:, None, None]
    j1 = m.get(O.DE_J1)
    assert np.abs(j1).max() > 0
    assert abs((j1 * dh).sum()) < 1e-9 * (np.abs(j1) * dh).sum()       # sum_l dh_l sum psi_l * (J + stretching) = 0
    assert not m.get(O.DE_J2).any()                                     # no background flow
    assert (m.get(O.DE_BF) * dh).sum() < 0                              # Ekman drag removes energy
    assert np.array_equal(m.get(O.PO_MFT), m.get(O.PSI))                # running mean after one sample
    m.energy_tend(1.0)
    assert np.allclose(m.get(O.DE_J1), 2 * j1, rtol=1e-13, atol=0)      # accumulates
    m.reset_energy()
    assert not m.get(O.DE_J1).any()


def test_tracer_mean_is_conserved_by_advection_and_diffusion():
    """ptr_rhs (qg.h:573-588): the Arakawa Jacobian with psi = 0 on the walls and the 5-point laplacian with
    zero-gradient ghosts both conserve the domain mean of a tracer; relaxation pulls it to the mean of ptr_relax."""
    N, nl, nptr = 64, 2, 2
    m = O.Model(O.make_params(**base_kw(N, nl, nptr=nptr, Pe=[30., 0.], ptr_r=[0., 0.05])))
    rng = np.random.default_rng(2)
    psi = np.zeros((nl, N, N))
    psi[:, 6:-6, 6:-6] = rng.standard_normal((nl, N - 12, N - 12))
    for _ in range(3):
        psi[:, 1:-1, 1:-1] = 0.25 * (psi[:, :-2, 1:-1] + psi[:, 2:, 1:-1] + psi[:, 1:-1, :-2] + psi[:, 1:-1, 2:])
    tr = rng.standard_normal((nl * nptr, N, N))
    m.set(O.PSI, 1e-2 * psi); m.set(O.PTR, tr); m.set(O.PTR_RELAX, np.full_like(tr, 3.0)); m.set_const()
    for _ in range(5):
        m.step()
    t2 = m.get(O.PTR)
    for l in range(nl):
        a, b = tr[l * nptr + 0], t2[l * nptr + 0]            # advected + diffused, not relaxed
        assert np.abs(b - a).max() > 0 and abs(b.mean() - a.mean()) < 1e-12 * np.abs(a).mean()
        a, b = tr[l * nptr + 1], t2[l * nptr + 1]            # relaxed towards 3
        assert abs(b.mean() - 3.0) < abs(a.mean() - 3.0)


def test_wavelet_filter_limits():
    """wavelet_filter (qg.h:509-560) with [BASILISK]'s wavelet()/inverse_wavelet(): sig_lev is a high-pass mask built
    from sig_filt = min(afilt*Rd, Lfmax) (qg.h:1057-1090).  A filter scale above every level keeps all coefficients
    (the reconstruction is the identity to round-off), a scale below the grid removes the whole field, and in
    between the filter is a linear projector-like operator: filtering twice changes little more."""
    N, nl = 64, 2
    psi = synth_psi(N, nl)
    out = {}
    for afilt in (1e9, 1e-3, 4.0):
        m = O.Model(O.make_params(**base_kw(N, nl, afilt=afilt, dtflt=0.05)))
        m.set(O.PSI, psi); m.set_const()
        q0 = m.get(O.Q)
        m.wavelet_filter(0.05)
        out[afilt] = (q0, m.get(O.Q), m.get(O.PSI), m.get(O.QOF), [float(m.siglev(l).min()) for l in range(7)],
                      [float(m.siglev(l).max()) for l in range(7)])
    q0, q1, p1, qof, lo, hi = out[1e9]
    assert lo == [1.0] * 7 and np.abs(q1 - q0).max() < 1e-13 and np.abs(p1 - psi).max() < 1e-12
    q0, q1, p1, qof, lo, hi = out[1e-3]
    assert hi == [0.0] * 7 and not p1.any() and not q1.any()
    assert np.allclose(qof, (q0 - q1) / 0.05, rtol=0, atol=1e-15)      # qof = (q_before - q_after)/dtflt with nbar = 0
    q0, q1, p1, qof, lo, hi = out[4.0]
    assert 0 < hi[6] and lo[0] == 0.0 and 0 < np.abs(q1).max() < np.abs(q0).max()


def test_energy_conserv_variant_is_the_same_operator_for_uniform_stretching():
    """ENERGY_CONSERV (qg.h:310-373) advects the full PV: jacobian(po, qot) instead of jacobian(po, zeta) plus the
    stretching Jacobians J(psi_l, psi_l+-1).  The discrete Jacobian is bilinear and J(psi, psi) == 0, so with
    horizontally uniform stretching q_l = zeta_l + s (psi_l+1 - psi_l) idh1 - ... gives the SAME tendency up to
    round-off -- an identity that pins the branch (argument order, sign, which list is advected).  With a
    stretching field s(x, y) the two builds differ by the terms J(psi, s) that only the energy-conserving form
    keeps.  (_LS_RV = 0, the other switch of qg.h, is the arithmetic of flsrv = 0: zetapl stays zero and the
    term adds exact zeros.)"""
    N, nl = 64, 3
    psi = synth_psi(N, nl)

    def tendency(econs, fr=None, **over):
        m = O.Model(O.make_params(**base_kw(N, nl, **over)))
        m.set_energy_conserv(econs)
        m.set(O.PSI, psi)
        if fr is not None:
            m.set(O.FR, fr)
        m.set_const()
        m.update(1e10)
        return m.get(O.DQ)

    a, b = tendency(0), tendency(1)
    assert not np.array_equal(a, b)
    assert np.abs(a - b).max() < 1e-11 * np.abs(a).max()
    over = dict(upg=[0.3, 0.1, 0.], vpg=[0.05, 0., 0.], flsrv=1, Re=200.)
    a, b = tendency(0, **over), tendency(1, **over)
    assert np.abs(a - b).max() < 1e-11 * np.abs(a).max()
    y, x = np.meshgrid((np.arange(N) + 0.5) / N, (np.arange(N) + 0.5) / N, indexing="ij")
    fr = np.zeros_like(psi)
    for l in range(nl - 1):
        fr[l] = (0.003 + 0.002 * l) * (1 + 0.3 * np.sin(2 * np.pi * x) * np.cos(np.pi * y))
    a, b = tendency(0, fr), tendency(1, fr)
    assert np.abs(a - b).max() > 1e-6 * np.abs(a).max()


def test_varying_vertical_modes_fields():
    """MODE_PV_INVERT with varRo > 0 (eigmode.h:74-299: one dgeev per column): the mode matrices and lambda = iBu are
    fields.  In every column cl2m*cm2l = I, the first baroclinic lambda is the analytic 2-layer value built from that
    column's Ro(y) (qg.h:1032-1037) and the columns of one row (same Ro) hold identical matrices."""
    if O._lapack_path() is None:
        pytest.skip("no LAPACK dgeev available")
    N, nl = 32, 2
    kw = base_kw(N, nl, mode_pv_invert=1, varRo=1)
    m = O.Model(O.make_params(**kw))
    m.set(O.PSI, synth_psi(N, nl)); m.set_const()
    ib, cl, cm = m.get(O.IBU), m.get(O.CL2M), m.get(O.CM2L)
    dh = np.array(kw["dh"]); F = kw["Fr"][0]; L0, Rom, beta = kw["L0"], kw["Rom"], kw["beta"]
    y = (np.arange(N) + 0.5) * L0 / N
    Ro = Rom / (1 + Rom * beta * (y - 0.5 * L0))
    dhc = 0.5 * (dh[0] + dh[1])
    lam = (F / Ro) ** 2 / dhc * (1 / dh[0] + 1 / dh[1])
    assert np.ptp(lam) > 0.1 * lam.mean()
    assert np.allclose(-ib[1], lam[:, None] * np.ones((1, N)), rtol=1e-11)
    assert not ib[0].any()
    for j in (0, N // 2, N - 1):
        for i in (0, 7):
            A = cl[:, j, i].reshape(nl, nl) @ cm[:, j, i].reshape(nl, nl)
            assert np.allclose(A, np.eye(nl), atol=1e-10)
        assert np.array_equal(cl[:, j, 0], cl[:, j, N - 1]) and np.array_equal(ib[:, j, 0], ib[:, j, 5])
