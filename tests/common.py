"""Shared synthetic inputs: deterministic, no file dependencies,
identical on the oracle and on the GPU."""
import numpy as np

DH = {2: [0.2, 0.8], 3: [0.06, 0.14, 0.8], 4: [0.05, 0.1, 0.25, 0.6], 10: [0.1] * 10}
FR = {2: [0.0045], 3: [0.0023669, 0.0076173], 4: [0.0024, 0.0050, 0.0076],
      10: list(np.linspace(0.002, 0.008, 9))}


def base_kw(N, nl, **over):
    kw = dict(N=N, nl=nl, L0=80., Rom=0.025, Ekb=0.002, tau0=1e-4, beta=0.5, CFL=0.6, DT=5e-2, Re=0.,
              Re4=1563. * (N / 256.) ** 4, dh=DH.get(nl, [1.0 / nl] * nl), Fr=FR.get(nl, list(np.linspace(0.002, 0.008, nl - 1))),
              tend=500., dtout=10.)
    kw.update(over)
    return kw


def synth_psi(N, nl, L0=80., seed=1234):
    rng = np.random.default_rng(seed)
    x = (np.arange(N) + 0.5) * L0 / N
    X, Y = np.meshgrid(x, x)  # [y][x]
    psi = np.zeros((nl, N, N))
    for l in range(nl):
        A = 1.0 / (l + 1)
        psi[l] = A * np.sin(np.pi * X / L0) * np.sin(2 * np.pi * Y / L0) + 1e-3 * A * rng.uniform(-1, 1, (N, N))
    return psi


def periodic_psi(N, nl, L0=80., seed=5):
    """a doubly periodic stream function (sbc = -1 cases): smooth waves + zero-mean noise"""
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


def rel_l2(a, b):
    d = np.linalg.norm((a - b).ravel())
    n = np.linalg.norm(b.ravel())
    return d / n if n > 0 else d


def make_pair(N, nl, smoother="lex", **over):
    """(oracle model, gpu model) with identical parameters, smoother ordering and initial psi."""
    from oracle import oracle as O
    from msom_b200 import capi as G
    kw = base_kw(N, nl, **over)
    po = O.make_params(**kw)
    pg = G.make_params(**kw)
    mo = O.Model(po)
    mg = G.Model(pg)
    mo.set_smoother(smoother)
    mg.set_smoother(smoother)
    psi = synth_psi(N, nl, kw["L0"])
    mo.set(O.PSI, psi)
    mg.set(G.PSI, psi)
    return mo, mg, psi
