"""CPU pins of what round 2 added to the oracle (no GPU needed):

  * the red-black ordering of relax_layer (orc_set_smoother / orc_test_relax_rb) against an independent numpy
    restatement, vectorised by colour -- which only works because a half-sweep does not depend on the traversal order,
    the property the throughput mode of the CUDA library rests on;
  * red-black and lexicographic sweeps reach the same solution to the solver tolerance with similar cycle counts;
  * Philox4x32-10 against the known-answer vectors of the Random123 distribution (kat_vectors), and the statistics and
    reproducibility of the noise field built on it."""
import ctypes as C

import numpy as np

from common import DH, FR, base_kw, rel_l2, synth_psi
from oracle import oracle as O


def _numpy_rb_sweeps(nl, n, L0, dh, s, a, b, nsweeps, periodic=False):
    """poisson_layer.h:80-146 per cell, all cells of one colour at once; same association order, IEEE ops only"""
    Delta = L0 / n
    dhc = [0.5 * (dh[l] + dh[l + 1]) for l in range(nl - 1)]
    idh0 = [0.0] + [1.0 / (dhc[l - 1] * dh[l]) for l in range(1, nl)]
    idh1 = [1.0 / (dhc[l] * dh[l]) for l in range(nl - 1)] + [0.0]
    a = a.copy()
    yy, xx = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    for _ in range(nsweeps):
        for colour in (0, 1):
            g = np.zeros((nl, n + 2, n + 2)) if not periodic else np.pad(a, ((0, 0), (1, 1), (1, 1)), mode="wrap")
            g[:, 1:-1, 1:-1] = a
            if not periodic:
                g[:, 0, 1:-1] = -a[:, 0, :]; g[:, -1, 1:-1] = -a[:, -1, :]   # homogeneous dirichlet ghosts
                g[:, 1:-1, 0] = -a[:, :, 0]; g[:, 1:-1, -1] = -a[:, :, -1]
            E, W = g[:, 1:-1, 2:], g[:, 1:-1, :-2]
            Nn, S = g[:, 2:, 1:-1], g[:, :-2, 1:-1]
            t0 = [None] * nl; t1 = [None] * nl; t2 = [None] * nl; rhs = [None] * nl
            for l in range(nl):
                rhs[l] = -(Delta * Delta) * b[l]
                if l == 0:
                    t2[l] = -(Delta * Delta) * s[l] * idh1[l]; t1[l] = -t2[l]
                elif l < nl - 1:
                    t0[l] = -(Delta * Delta) * s[l - 1] * idh0[l]; t2[l] = -(Delta * Delta) * s[l] * idh1[l]
                    t1[l] = -t0[l] - t2[l]
                else:
                    t0[l] = -(Delta * Delta) * s[l - 1] * idh0[l]; t1[l] = -t0[l]
                rhs[l] = rhs[l] + (1. * E[l] + 1. * W[l]); t1[l] = t1[l] + (1. + 1.)
                rhs[l] = rhs[l] + (1. * Nn[l] + 1. * S[l]); t1[l] = t1[l] + (1. + 1.)
            for l in range(1, nl):
                t1[l] = t1[l] - t0[l] * t2[l - 1] / t1[l - 1]
                rhs[l] = rhs[l] - t0[l] * rhs[l - 1] / t1[l - 1]
            new = [None] * nl
            new[nl - 1] = rhs[nl - 1] / t1[nl - 1]
            for l in range(nl - 2, -1, -1):
                new[l] = (rhs[l] - t2[l] * new[l + 1]) / t1[l]
            mask = ((xx + yy) & 1) == colour
            for l in range(nl):
                a[l][mask] = new[l][mask]
    return a


def test_red_black_sweep_matches_numpy_restatement():
    for nl, level, nsweeps in ((2, 3, 1), (3, 4, 3), (4, 5, 4)):
        n = 1 << level
        rng = np.random.default_rng(level)
        a = rng.standard_normal((nl, n, n)); b = rng.standard_normal((nl, n, n))
        s = np.empty((nl - 1, n, n))
        for l in range(nl - 1):
            s[l] = (FR[nl][l] / 0.025) ** 2 * (1 + 0.1 * rng.random((n, n)))     # horizontally varying stretching too
        dh = np.array(DH[nl], dtype=np.float64)
        ref = _numpy_rb_sweeps(nl, n, 80., list(dh), s, a, b, nsweeps)
        got = a.copy()
        O.lib().orc_test_relax_rb(nl, level, 80., dh, np.ascontiguousarray(s), got, b, nsweeps)
        assert np.array_equal(got, ref), np.abs(got - ref).max()
        lex = a.copy()
        O.lib().orc_test_relax(nl, level, 80., dh, np.ascontiguousarray(s), lex, b, nsweeps, 1, 1)
        assert not np.array_equal(lex, got)            # a different iterate from the reference order ...
        assert np.abs(lex - got).max() < np.abs(got).max()  # ... of the same size


def test_red_black_and_reference_order_agree_to_solver_tolerance():
    N, nl = 128, 3
    out = {}
    for sm in ("lex", "rb"):
        m = O.Model(O.make_params(**base_kw(N, nl)))
        m.set_smoother(sm)
        m.set(O.PSI, synth_psi(N, nl)); m.set_const()
        for _ in range(5):
            m.step()
        out[sm] = (m.get(O.PSI), m.get(O.Q), m.L.orc_total_cycles(m.h), m.mgstats().resa)
        m.close()
    assert out["lex"][3] <= 1e-3 and out["rb"][3] <= 1e-3
    assert 0 < rel_l2(out["rb"][0], out["lex"][0]) < 5e-3
    assert 0 < rel_l2(out["rb"][1], out["lex"][1]) < 5e-3
    assert abs(out["rb"][2] - out["lex"][2]) <= max(2, out["lex"][2] // 2)


def test_red_black_ignores_the_decomposition_emulation():
    """orc_set_decomp changes the lexicographic iterate (block Gauss-Seidel) but not the red-black one"""
    N, nl = 64, 2
    res = []
    for px, py in ((1, 1), (2, 2)):
        m = O.Model(O.make_params(**base_kw(N, nl)))
        m.set_smoother("rb"); m.L.orc_set_decomp(m.h, px, py, 16)
        m.set(O.PSI, synth_psi(N, nl)); m.set_const()
        m.step(); m.step()
        res.append(m.get(O.PSI)); m.close()
    assert np.array_equal(res[0], res[1])


def test_philox_known_answers_and_noise_statistics():
    L = O.lib()
    L.orc_test_philox.argtypes = [C.POINTER(C.c_uint), C.c_uint, C.c_uint]
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:                      # Random123 kat_vectors, philox4x32 10 rounds
        c = (C.c_uint * 4)(*ctr)
        L.orc_test_philox(c, key[0], key[1])
        assert tuple(c) == want
    N, nl = 64, 3
    kw = base_kw(N, nl, stochastic=1, tr_stoch=10., amp_stoch=2.)
    fields = []
    for seed in (7, 7, 8):
        m = O.Model(O.make_params(**kw))
        m.set(O.PSI, synth_psi(N, nl)); m.set(O.SSTOCH, np.full((nl, N, N), 0.5)); m.set_const()
        L.orc_set_noise_mode(m.h, 1, seed)
        m.step()
        fields.append(m.get(O.NSTOCH)); m.close()
    assert np.array_equal(fields[0], fields[1]) and not np.array_equal(fields[0], fields[2])
    z = fields[0] / (2. * 0.5)                       # amp * sigma
    assert abs(z.mean()) < 5 / np.sqrt(z.size) and abs(z.std() - 1) < 0.03
    assert abs(np.mean(z ** 3)) < 0.1 and abs(np.mean(z ** 4) - 3) < 0.3


def test_red_black_golden_regression():
    """the oracle with the red-black ordering reproduces its committed outputs (tests/golden/make_golden.py)"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_rb_32x2_3steps.npz"))
    m = O.Model(O.make_params(**base_kw(32, 2)))
    m.set_smoother("rb")
    m.set(O.PSI, synth_psi(32, 2)); m.set_const()
    dts = [m.step() for _ in range(3)]
    assert np.array_equal(np.array(dts), g["dts"])
    assert np.array_equal(m.get(O.PSI), g["psi"]) and np.array_equal(m.get(O.Q), g["q"])


def test_periodic_domain_is_translation_equivariant():
    """sbc = -1 (qg.h:842-846): with periodic(right); periodic(top) every operator of the step has constant
    coefficients once the wind forcing is off, and the multigrid hierarchy (minlevel = 1: two cells per side) maps onto
    itself under a diagonal shift by N/2 cells in x and y.  A red-black half-sweep does not depend on the traversal order, so the whole
    step -- ghost rings on every level, restriction, bilinear prolongation and the nine-point Jacobian across the seam
    -- must commute with that shift BIT FOR BIT.  This pins the wrap-around of the oracle's boundary_level()."""
    import numpy as np
    from oracle import oracle as O
    from common import base_kw
    N, nl = 64, 3
    rng = np.random.default_rng(11)
    x = (np.arange(N) + 0.5) / N
    X, Y = np.meshgrid(x, x)
    psi = np.zeros((nl, N, N))
    for l in range(nl):
        nz = rng.uniform(-1, 1, (N, N))
        psi[l] = (np.sin(2 * np.pi * X) * np.sin(4 * np.pi * Y) + 0.3 * np.cos(2 * np.pi * (X + 2 * Y)) + 1e-3 * (nz - nz.mean())) / (l + 1)

    def run(p0, nsteps=3):
        m = O.Model(O.make_params(**base_kw(N, nl, sbc=-1., tau0=0.)))
        m.set_smoother("rb")
        m.set(O.PSI, p0); m.set_const()
        q0 = m.get(O.Q)
        dts = [m.step() for _ in range(nsteps)]
        return q0, m.get(O.Q), m.get(O.PSI), dts, m.L.orc_total_cycles(m.h)

    q0, q1, p1, dts, cyc = run(psi)
    # the discrete PV of a periodic stream function integrates to zero in every layer (compatibility of the singular solve)
    assert np.abs(q0.sum(axis=(1, 2))).max() < 1e-9 * np.abs(q0).sum()
    # the shift is diagonal: on level 1 (2 x 2 cells) a shift along one axis alone would swap the two colours
    sh = lambda a: np.roll(a, (N // 2, N // 2), axis=(1, 2))
    r0, r1, rp, rdts, rcyc = run(sh(psi))
    assert rdts == dts and rcyc == cyc
    assert np.array_equal(r0, sh(q0))
    assert np.array_equal(r1, sh(q1))
    assert np.array_equal(rp, sh(p1))
    # a shift along x alone is the same solve with red and black swapped on level 1: equal to the solver tolerance only
    r0, r1, rp, rdts, rcyc = run(np.roll(psi, N // 2, axis=2))
    assert np.array_equal(r0, np.roll(q0, N // 2, axis=2)) and not np.array_equal(rp, np.roll(p1, N // 2, axis=2))
    d = rp - np.roll(p1, N // 2, axis=2)
    d -= d.mean(axis=(1, 2), keepdims=True)      # the constant is in the null space of the periodic operator
    assert np.abs(d).max() < 1e-3 * np.abs(p1).max()
    # and the seam is really open: the closed basin gives another PV for the same stream function
    m = O.Model(O.make_params(**base_kw(N, nl, tau0=0.)))
    m.set(O.PSI, psi); m.set_const()
    assert not np.array_equal(m.get(O.Q), q0)
