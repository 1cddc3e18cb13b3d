"""Generates the regression fixtures of tests/golden from the CPU oracle.

The reference (Basilisk C) cannot be built or imported here, so these are NOT
reference outputs: they pin the oracle against accidental change (parity with
the reference itself stays "unpinned").  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402
from common import base_kw, periodic_psi, synth_psi  # noqa: E402
from oracle import oracle as O  # noqa: E402


def run(N, nl, steps, smoother="lex", econs=0, psi=None, **over):
    m = O.Model(O.make_params(**base_kw(N, nl, **over)))
    m.set_smoother(smoother)
    m.set_energy_conserv(econs)
    m.set(O.PSI, synth_psi(N, nl) if psi is None else psi)
    m.set_const()
    dts = np.array([m.step() for _ in range(steps)])
    return dict(dts=dts, psi=m.get(O.PSI), q=m.get(O.Q))


np.savez_compressed(os.path.join(HERE, "oracle_32x2_3steps.npz"), **run(32, 2, 3))
np.savez_compressed(os.path.join(HERE, "oracle_32x3_modal_2steps.npz"), **run(32, 3, 2, mode_pv_invert=1))
np.savez_compressed(os.path.join(HERE, "oracle_rb_32x2_3steps.npz"), **run(32, 2, 3, smoother="rb"))
# second half of round 2: periodic domain (red-black), ENERGY_CONSERV build variant, per-column vertical modes (Ro(y))
np.savez_compressed(os.path.join(HERE, "oracle_rb_periodic_32x3_3steps.npz"), **run(32, 3, 3, smoother="rb", psi=periodic_psi(32, 3), sbc=-1.))
np.savez_compressed(os.path.join(HERE, "oracle_econs_32x2_3steps.npz"), **run(32, 2, 3, econs=1))
np.savez_compressed(os.path.join(HERE, "oracle_32x3_modal_varRo_2steps.npz"), **run(32, 3, 2, mode_pv_invert=1, varRo=1))
print("wrote", os.listdir(HERE))
