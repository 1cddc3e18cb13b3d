"""CPU-side checks of the drop-in boundary: the C-ABI libraries load, export every
symbol include/msqg.h declares, the host logic (params, .bas files) agrees with
the oracle, and compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from common import base_kw, synth_psi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "msom_b200", "lib")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "msqg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b([a-z_][a-z0-9_]*)\s*\([^;{]*\)\s*;", hdr))
    names -= {"defined"}
    assert len(names) > 50
    cuda = C.CDLL(os.path.join(LIBDIR, "libmsqg_cuda.so"), mode=C.RTLD_GLOBAL)
    host = C.CDLL(os.path.join(LIBDIR, "libqg.so"))
    missing = [n for n in sorted(names) if not (hasattr(cuda, n) or hasattr(host, n))]
    assert not missing, missing


def test_params_match_oracle(tmp_path):
    from oracle import oracle as O
    from msom_b200 import capi as G
    f = tmp_path / "params.in"
    f.write_text("#!sh\nN = 128\nnl = 4\nL0 = 80\nRom = 0.025\nEkb = 0.002\nEks = 0.001\ntau0 = 0.0001\nRe = 300\n"
                 "Re4 = 97.6875\nbeta = 0.5\nFr = [0.0024, 0.0050,0.0076]\ndh = [0.05,0.1,0.25,0.6]\nupg = [0.1,0,0,0]\n"
                 "DT = 5.e-2\ntend = 3.\ndtout = 0.5\nCFL = 0.6\nflsrv = 1\nvarRo = 0\n")
    po = O.Params(); O.lib().orc_default_params(po); assert O.lib().orc_read_params(str(f).encode(), po) == 0
    pg = G.read_params(f)
    for k, _ in G.Params._fields_:
        a, b = getattr(po, k), getattr(pg, k)
        if hasattr(a, "__len__"):
            assert list(a) == list(b), k
        else:
            assert a == b, k
    with pytest.raises(G.MsqgError) as e:
        G.read_params(tmp_path / "missing.in")
    assert e.value.code == G.ERR_FILE
    # make_params applies the same derived values as the file path
    kw = base_kw(256, 3)
    assert G.make_params(**kw).DT == O.make_params(**kw).DT == 0.025


def test_bas_files_match_oracle(tmp_path):
    from oracle import oracle as O
    import msom_b200.qg as bas
    L = bas._L()
    N, nl = 32, 3
    v = synth_psi(N, nl)
    fo, fg = str(tmp_path / "o.bas"), str(tmp_path / "g.bas")
    assert O.lib().orc_write_bas(fo.encode(), nl, N, 80., v) == 0
    assert L.qg_write_bas(fg.encode(), nl, N, 80., v) == 0
    assert open(fo, "rb").read() == open(fg, "rb").read()
    a, b = np.zeros_like(v), np.zeros_like(v)
    assert O.lib().orc_read_bas(fo.encode(), nl, N, 80., a) == 0
    assert L.qg_read_bas(fo.encode(), nl, N, 80., b) == 0
    assert np.array_equal(a, b)
    # reading a coarser file onto a finer grid (input_matrixl's nearest-cell lookup)
    a2, b2 = np.zeros((nl, 64, 64)), np.zeros((nl, 64, 64))
    assert O.lib().orc_read_bas(fo.encode(), nl, 64, 80., a2) == 0
    assert L.qg_read_bas(fo.encode(), nl, 64, 80., b2) == 0
    assert np.array_equal(a2, b2)
    assert L.qg_read_bas(b"/nonexistent.bas", nl, N, 80., b) != 0


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu(tmp_path):
    from msom_b200 import capi as G
    with pytest.raises(G.MsqgError) as e:
        G.Model(G.make_params(**base_kw(64, 2)))
    assert e.value.code == G.ERR_CUDA and "no CPU path" in str(e.value)
    f = tmp_path / "params.in"
    f.write_text("N = 64\nnl = 2\nL0 = 80\nRom = 0.025\nFr = [0.0045]\ndh = [0.2,0.8]\n")
    out = subprocess.run([os.path.join(LIBDIR, "qg.e")], cwd=str(tmp_path), capture_output=True, text=True, timeout=120)
    assert out.returncode != 0 and "no CUDA device" in out.stdout


def test_qg_exe_missing_params(tmp_path):
    """qg.c:36-41 + qg.h:735-738: message on stdout, exit(0)."""
    out = subprocess.run([os.path.join(LIBDIR, "qg.e"), "nope.in"], cwd=str(tmp_path), capture_output=True, text=True,
                         timeout=120)
    assert out.returncode == 0 and "file nope.in not found" in out.stdout


def test_argument_validation():
    from msom_b200 import capi as G
    for kw, frag in ((dict(nl=1), "nl must be"), (dict(N=100), "power of two"), (dict(sbc=-2.0), "sbc"),
                     (dict(nptr=1, stochastic=1), "tracers"), (dict(nptr=99), "nptr")):
        k = base_kw(64, 2); k.update(kw)
        if "nl" in kw:
            k.update(dh=[1.0], Fr=[])
        with pytest.raises(G.MsqgError) as e:
            G.Model(G.make_params(**k))
        assert e.value.code == G.ERR_ARG and frag in str(e.value)
