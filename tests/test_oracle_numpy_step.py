"""Whole timesteps of the reference -- [BASILISK] predictor-corrector run() with update = update_qg and
advance = advance_qg (msqg/qg.h:594-650, :922-923), i.e. two PV inversions, two right-hand sides, the midpoint
update and the dt chain of timestep() with its static `previous` -- assembled from the independent numpy restatements
of tests/test_oracle_numpy_rhs.py and tests/test_oracle_numpy_mg.py and run beside the C oracle (red-black order):
same dt sequence, same cycle counts, q and psi equal to round-off after several steps.  (CPU test, no GPU.)"""
import numpy as np
import pytest

from common import base_kw, synth_psi
from oracle import oracle as O
from test_oracle_numpy_mg import numpy_solve
from test_oracle_numpy_rhs import jac, lap, pad, sh, stretch


class NumpyModel:
    def __init__(self, kw, psi):
        self.kw = kw
        self.nl, self.N = psi.shape[0], psi.shape[1]
        self.L0, self.D = kw["L0"], kw["L0"] / psi.shape[1]
        self.dh = np.array(kw["dh"], dtype=float)
        dhc = 0.5 * (self.dh[:-1] + self.dh[1:])
        self.idh0 = np.zeros(self.nl); self.idh1 = np.zeros(self.nl)
        self.idh1[:-1] = 1. / (dhc * self.dh[:-1]); self.idh0[1:] = 1. / (dhc * self.dh[1:])
        self.s = np.array([(fr / kw["Rom"]) ** 2 * np.ones((self.N, self.N)) for fr in kw["Fr"]])
        # read_params, qg.h:739-746: iRe4 = -1/Re4 and the viscous limit on DT (CFL 0.5)
        self.iRe4 = -1. / kw["Re4"] if kw["Re4"] else 0.
        self.DT = 0.5 * min(kw["DT"], (self.L0 / self.N) ** 4 * kw["Re4"] / 32.) if kw["Re4"] else kw["DT"]
        self.psi = psi.copy()
        P = [pad(psi[l], -1) for l in range(self.nl)]
        self.q = np.array([lap(P[l], self.D) for l in range(self.nl)]) + stretch(psi, list(self.s), self.idh0, self.idh1)  # comp_q
        self.prev = 0.          # timestep()'s static `previous`
        self.cycles = 0
        y = (np.arange(self.N) + 0.5) * self.D
        self.wind = kw["tau0"] / (kw["Rom"] * self.dh[0]) * (np.sin(2 * np.pi * y / self.L0) * np.sin(np.pi * y / self.L0))[:, None]

    def update(self, q, dtmax):
        kw, nl, D = self.kw, self.nl, self.D
        self.psi, st, _ = numpy_solve(self.psi, q, self.s, self.dh, self.L0)          # invertq: warm start
        self.cycles += st["i"]
        P = [pad(self.psi[l], -1) for l in range(nl)]
        zeta = np.array([lap(P[l], D) for l in range(nl)])
        Z = [pad(zeta[l], -1) for l in range(nl)]
        jd = [jac(P[l], P[l + 1], D) for l in range(nl - 1)]
        dq = np.zeros_like(q)
        for l in range(nl):
            t = jac(P[l], Z[l], D) + kw["beta"] * (sh(P[l], -1, 0) - sh(P[l], 1, 0)) / (2 * D)
            if l > 0:
                t = t + self.s[l - 1] * (-jd[l - 1]) * self.idh0[l]
            if l < nl - 1:
                t = t + self.s[l] * jd[l] * self.idh1[l]
            dq[l] = t
        tmp = np.array([lap(Z[l], D) for l in range(nl)])
        T = [pad(tmp[l], -1) for l in range(nl)]
        dq += self.iRe4 * stretch(tmp, list(self.s), self.idh0, self.idh1) + self.iRe4 * np.array([lap(T[l], D) for l in range(nl)])
        dq[-1] -= kw["Ekb"] / (kw["Rom"] * 2 * self.dh[-1]) * zeta[-1]
        dq[0] -= self.wind
        CFL = kw["CFL"]
        for l in range(nl):
            Pl = P[l]; n = self.N
            ux = -0.25 * (Pl[2:n + 2, 1:n + 2] - Pl[0:n, 1:n + 2] + Pl[2:n + 2, 0:n + 1] - Pl[0:n, 0:n + 1]) / D
            uy = 0.25 * (Pl[1:n + 2, 2:n + 2] - Pl[1:n + 2, 0:n] + Pl[0:n + 1, 2:n + 2] - Pl[0:n + 1, 0:n]) / D
            for u in (max(np.abs(ux).max(), np.abs(uy).max()), 0.):   # psi, then psi_pg (zero: no limit, but the chain runs)
                dtmax = dtmax / CFL
                if u != 0 and D / u < dtmax:
                    dtmax = D / u
                dtmax *= CFL
                if dtmax > self.prev:
                    dtmax = (self.prev + 0.1 * dtmax) / 1.1
                self.prev = dtmax
        return dq, dtmax

    def step(self):
        dq, dt = self.update(self.q, self.DT)
        qp = self.q + dq * (dt / 2.)
        dq, _ = self.update(qp, dt)
        self.q = self.q + dq * dt
        return dt


@pytest.mark.parametrize("N,nl,nsteps", [(32, 2, 4), (64, 3, 3)])
def test_timesteps_against_numpy_restatement(N, nl, nsteps):
    kw = base_kw(N, nl)
    psi = synth_psi(N, nl)
    m = O.Model(O.make_params(**kw)); m.set_smoother("rb")
    m.set(O.PSI, psi); m.set_const()
    ref = NumpyModel(kw, psi)
    assert np.abs(ref.q - m.get(O.Q)).max() <= 1e-12 * np.abs(ref.q).max()
    for k in range(nsteps):
        dto, dtn = m.step(), ref.step()
        assert dto == pytest.approx(dtn, rel=1e-11), k
    assert m.L.orc_total_cycles(m.h) == ref.cycles
    for a, b in ((m.get(O.Q), ref.q), (m.get(O.PSI), ref.psi)):
        assert np.abs(a - b).max() <= 1e-9 * np.abs(b).max()


class NumpyStochasticModel(NumpyModel):
    """the -D_STOCHASTIC=1 build (qg_stochastic.h): no J(psi, zeta) in the top layer and no J(psi_l, psi_l+1) anywhere
    (:35-40, :56-58), relaxation -q/tr_stoch (:44, :68, :91), and advance_qg adds n*dts with a FLOAT dts and a noise
    field drawn from libc rand() in traversal order (x outer, y inner, layers innermost) at every other call (:117-149)"""

    def __init__(self, kw, psi, sigma, libc):
        super().__init__(kw, psi)
        self.sigma, self.libc = sigma, libc
        self.corrector_step = 0
        self.noise = np.zeros_like(psi)
        self.itr = 1. / kw["tr_stoch"]

    def update(self, q, dtmax):
        dq, dtmax = super().update(q, dtmax)
        nl, D = self.nl, self.D
        P = [pad(self.psi[l], -1) for l in range(nl)]
        zeta = np.array([lap(P[l], D) for l in range(nl)])
        Z = [pad(zeta[l], -1) for l in range(nl)]
        jd = [jac(P[l], P[l + 1], D) for l in range(nl - 1)]
        for l in range(nl):          # take the terms of the deterministic build back out
            if l > 0:
                dq[l] -= self.s[l - 1] * (-jd[l - 1]) * self.idh0[l]
            if l < nl - 1:
                dq[l] -= self.s[l] * jd[l] * self.idh1[l]
        dq[0] -= jac(P[0], Z[0], D)
        dq += -q * self.itr
        return dq, dtmax

    def generate_noise(self):
        RAND_MAX = 2147483647
        n, nl = self.N, self.nl
        rand = self.libc.rand
        for i in range(n):
            for j in range(n):
                for l in range(nl):
                    r1 = rand(); r2 = rand()      # gcc evaluates the left operand of the product first
                    g = np.sqrt(-2. * np.log((float(r1) + 1.) / (float(RAND_MAX) + 2.))) * np.cos(2 * np.pi * r2 / float(RAND_MAX))
                    self.noise[l, j, i] = self.kw["amp_stoch"] * self.sigma[l, j, i] * g

    def advance(self, q_in, dq, dt):
        self.corrector_step = (self.corrector_step + 1) % 2
        dts = np.float32(np.sqrt(dt))
        if self.corrector_step:
            self.generate_noise()
            dts = np.float32(dts / np.sqrt(2))    # float / double -> double, stored in a float
        return q_in + dq * dt + self.noise * float(dts)

    def step(self):
        dq, dt = self.update(self.q, self.DT)
        qp = self.advance(self.q, dq, dt / 2.)
        dq, _ = self.update(qp, dt)
        self.q = self.advance(self.q, dq, dt)
        return dt


def test_stochastic_timesteps_against_numpy_restatement():
    import ctypes as C
    libc = C.CDLL(None)
    N, nl, nsteps = 32, 3, 3
    kw = base_kw(N, nl, stochastic=1, tr_stoch=10., amp_stoch=1.)
    psi = synth_psi(N, nl)
    sigma = 1e-3 * (1 + np.arange(nl)[:, None, None]) * np.ones((nl, N, N))
    m = O.Model(O.make_params(**kw)); m.set_smoother("rb")
    m.set(O.PSI, psi); m.set(O.SSTOCH, sigma); m.set_const()
    libc.srand(77)
    dto = [m.step() for _ in range(nsteps)]
    ref = NumpyStochasticModel(kw, psi, sigma, libc)
    libc.srand(77)
    dtn = [ref.step() for _ in range(nsteps)]
    assert dto == pytest.approx(dtn, rel=1e-11)
    noise = m.get(O.NSTOCH)
    assert np.abs(noise - ref.noise).max() <= 1e-15 * np.abs(noise).max() and np.abs(noise).max() > 0
    for a, b in ((m.get(O.Q), ref.q), (m.get(O.PSI), ref.psi)):
        assert np.abs(a - b).max() <= 1e-9 * np.abs(b).max()
