"""GPU parity through the reference-facing surfaces: the `qg` Python module
(msqg/qg.i, qg_bfn.i entry points) and the qg.e process (msqg/qg.c) with its
outdir_NNNN/*.bas outputs, against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from common import base_kw, synth_psi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_params(path, N, nl, tend=0.2, dtout=0.1, extra=""):
    kw = base_kw(N, nl)
    with open(path, "w") as f:
        f.write("#!sh\nN  = %d\nnl = %d\nL0 = %g\nRom = %g\nEkb = %g\ntau0 = %g\nRe4 = %r\nbeta = %g\n"
                "Fr = [%s]\ndh = [%s]\nDT = %g\ntend = %g\ndtout = %g\nCFL = %g\n%s" % (
                    N, nl, kw["L0"], kw["Rom"], kw["Ekb"], kw["tau0"], kw["Re4"], kw["beta"],
                    ",".join(repr(float(x)) for x in kw["Fr"]), ",".join(repr(float(x)) for x in kw["dh"]),
                    kw["DT"], tend, dtout, kw["CFL"], extra))


def test_python_module_matches_oracle(gpu, tmp_path, monkeypatch):
    """read_params -> init_grid -> set_vars -> set_vars_bfn -> set_const -> pyp2q -> pystep_bfn -> pyq2p
    (the call order of msqg/qg_bfn.py) against the oracle's restatement of qg_bfn.h."""
    from oracle import oracle as O
    import msom_b200.qg as bas
    N, nl = 64, 3
    monkeypatch.chdir(tmp_path)
    _write_params("params.in", N, nl)
    bas.read_params("params.in"); bas.init_grid(N); bas.set_vars(); bas.set_vars_bfn(); bas.set_const()
    po = O.Params(); O.lib().orc_default_params(po); O.lib().orc_read_params(b"params.in", po)
    mo = O.Model(po); mo.set_const()
    p = synth_psi(N, nl)
    q = np.zeros_like(p); q_ref = np.zeros_like(p)
    bas.pyp2q(p, q); mo.L.orc_pyp2q(mo.h, p, q_ref)
    assert np.array_equal(q, q_ref)
    for direction in (1.0, -1.0):
        F = np.zeros_like(p); F_ref = np.zeros_like(p)
        bas.pystep_bfn(q, F, direction, 1); mo.L.orc_pystep_bfn(mo.h, q, F_ref, direction, 1)
        assert np.array_equal(F, F_ref)
    p2 = np.zeros_like(p); p2_ref = np.zeros_like(p)
    bas.pyq2p(p2, q); mo.L.orc_pyq2p(mo.h, p2_ref, q)
    assert np.array_equal(p2, p2_ref)
    bas.trash_vars(); bas.trash_vars_bfn()


def test_qg_exe_outputs_match_oracle_run(gpu, tmp_path):
    """./qg.e with p0.bas: same stdout cadence and bit-identical po/qo .bas files as the oracle's run()."""
    from oracle import oracle as O
    N, nl = 64, 2
    wd = tmp_path / "gpu"; wd.mkdir()
    wo = tmp_path / "orc"; wo.mkdir()
    _write_params(str(wd / "params.in"), N, nl, tend=0.1, dtout=0.05)
    psi = synth_psi(N, nl)
    O.lib().orc_write_bas(str(wd / "p0.bas").encode(), nl, N, 80., psi)
    exe = os.path.join(ROOT, "msom_b200", "lib", "qg.e")
    out = subprocess.run([exe], cwd=str(wd), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Config: N = 64, nl = 2, L0 = 80" in out.stdout and "write file" in out.stdout
    # oracle: same init path (psi read back through float32 p0.bas, mean removed)
    po = O.Params(); O.lib().orc_default_params(po); O.lib().orc_read_params(str(wd / "params.in").encode(), po)
    mo = O.Model(po)
    p0 = np.zeros_like(psi); O.lib().orc_read_bas(str(wd / "p0.bas").encode(), nl, N, 80., p0)
    mo.set(O.PSI, p0); mo.L.orc_remove_mean_psi(mo.h); mo.set_const()
    nsteps = mo.run(outdir=str(wo))
    gdir = wd / "outdir_0001"
    names = sorted(f for f in os.listdir(wo) if f.endswith(".bas"))
    assert len(names) == 6 and nsteps > 0          # po/qo at t = 0, 0.05, 0.1
    for f in names:
        assert (gdir / f).read_bytes() == (wo / f).read_bytes(), f
    for f in ("params.in", "sig_filt.bas", "rdpg_2l_N64.bas", "psipg_2l_N64.bas", "frpg_2l_N64.bas",
              "qforc_2l_N64.bas", "dh_2l.bin"):
        assert (gdir / f).exists(), f
    # the readers of msqg/scripts/read_data.py:22-46 see the field
    a = np.fromfile(str(gdir / names[0]), "f4").reshape(nl, N + 1, N + 1).transpose(0, 2, 1)[:, 1:, 1:]
    assert a.shape == (nl, N, N) and np.isfinite(a).all()


def test_qg_exe_periodic_domain(gpu, tmp_path):
    """sbc = -1 in params.in (qg.h:80,711,842-846): ./qg.e runs the doubly periodic domain (msqg_create makes the 1 x 1
    periodic group by itself); po/qo files byte for byte those of the oracle's run() with the same sweep ordering."""
    from oracle import oracle as O
    from test_gpu_periodic import periodic_psi
    N, nl = 64, 3
    wd = tmp_path / "gpu"; wd.mkdir()
    wo = tmp_path / "orc"; wo.mkdir()
    _write_params(str(wd / "params.in"), N, nl, tend=0.1, dtout=0.05, extra="sbc = -1\n")
    psi = periodic_psi(N, nl)
    O.lib().orc_write_bas(str(wd / "p0.bas").encode(), nl, N, 80., psi)
    exe = os.path.join(ROOT, "msom_b200", "lib", "qg.e")
    out = subprocess.run([exe], cwd=str(wd), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    po = O.Params(); O.lib().orc_default_params(po); O.lib().orc_read_params(str(wd / "params.in").encode(), po)
    assert po.sbc == -1
    mo = O.Model(po); mo.set_smoother("rb")
    p0 = np.zeros_like(psi); O.lib().orc_read_bas(str(wd / "p0.bas").encode(), nl, N, 80., p0)
    mo.set(O.PSI, p0); mo.L.orc_remove_mean_psi(mo.h); mo.set_const()
    assert mo.run(outdir=str(wo)) > 0
    gdir = wd / "outdir_0001"
    names = sorted(f for f in os.listdir(wo) if f.endswith(".bas"))
    assert len(names) == 6
    for f in names:
        assert (gdir / f).read_bytes() == (wo / f).read_bytes(), f


def test_qg_exe_energy_diagnostics_files(gpu, tmp_path):
    """ediag = 0 (msqg/qg_energy.h): ./qg.e writes de_{bf,vd,j1,j2,j3,ft}%09d.bas at every output, scaled by
    1/dtout and reset (qg.c:139-166); bit-identical to the oracle's run()."""
    from oracle import oracle as O
    N, nl = 64, 3
    wd = tmp_path / "gpu"; wd.mkdir()
    wo = tmp_path / "orc"; wo.mkdir()
    _write_params(str(wd / "params.in"), N, nl, tend=0.1, dtout=0.05, extra="ediag = 0\n")
    psi = synth_psi(N, nl)
    O.lib().orc_write_bas(str(wd / "p0.bas").encode(), nl, N, 80., psi)
    exe = os.path.join(ROOT, "msom_b200", "lib", "qg.e")
    out = subprocess.run([exe], cwd=str(wd), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    po = O.Params(); O.lib().orc_default_params(po); O.lib().orc_read_params(str(wd / "params.in").encode(), po)
    assert po.ediag == 0
    mo = O.Model(po)
    p0 = np.zeros_like(psi); O.lib().orc_read_bas(str(wd / "p0.bas").encode(), nl, N, 80., p0)
    mo.set(O.PSI, p0); mo.L.orc_remove_mean_psi(mo.h); mo.set_const()
    assert mo.run(outdir=str(wo)) > 0
    gdir = wd / "outdir_0001"
    names = sorted(f for f in os.listdir(wo) if f.endswith(".bas"))
    assert len(names) == 3 * 8 and sum(f.startswith("de_") for f in names) == 18
    for f in names:
        assert (gdir / f).read_bytes() == (wo / f).read_bytes(), f
    # the budget terms are not trivially zero after the first window
    a = np.fromfile(str(gdir / [f for f in names if f.startswith("de_j1")][-1]), "f4")
    assert np.isfinite(a).all() and np.abs(a).max() > 0


def test_qg_exe_passive_tracers_files(gpu, tmp_path):
    """nptr = 2 in params.in: ./qg.e reads ptr0.bas / ptr_relax.bas (qg.c:75-90) and writes ptr%09d.bas at every
    output (qg.c:168-171); bit-identical to the oracle's run()."""
    from oracle import oracle as O
    N, nl, nptr = 64, 2, 2
    wd = tmp_path / "gpu"; wd.mkdir()
    wo = tmp_path / "orc"; wo.mkdir()
    _write_params(str(wd / "params.in"), N, nl, tend=0.1, dtout=0.05, extra="nptr = 2\nPe = [40.,0.]\nptr_r = [0.,2.]\n")
    psi = synth_psi(N, nl)
    rng = np.random.default_rng(4)
    tr = rng.standard_normal((nl * nptr, N, N)); rl = rng.standard_normal((nl * nptr, N, N))
    O.lib().orc_write_bas(str(wd / "p0.bas").encode(), nl, N, 80., psi)
    O.lib().orc_write_bas(str(wd / "ptr0.bas").encode(), nl * nptr, N, 80., tr)
    O.lib().orc_write_bas(str(wd / "ptr_relax.bas").encode(), nl * nptr, N, 80., rl)
    exe = os.path.join(ROOT, "msom_b200", "lib", "qg.e")
    out = subprocess.run([exe], cwd=str(wd), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    po = O.Params(); O.lib().orc_default_params(po); O.lib().orc_read_params(str(wd / "params.in").encode(), po)
    assert po.nptr == 2 and po.iPe[0] == 1 / 40. and po.ptr_ir[1] == 0.5
    mo = O.Model(po)
    p0 = np.zeros_like(psi); O.lib().orc_read_bas(str(wd / "p0.bas").encode(), nl, N, 80., p0)
    t0 = np.zeros_like(tr); O.lib().orc_read_bas(str(wd / "ptr0.bas").encode(), nl * nptr, N, 80., t0)
    r0 = np.zeros_like(rl); O.lib().orc_read_bas(str(wd / "ptr_relax.bas").encode(), nl * nptr, N, 80., r0)
    mo.set(O.PSI, p0); mo.L.orc_remove_mean_psi(mo.h)
    mo.set(O.PTR, t0); mo.set(O.PTR_RELAX, r0); mo.set_const()
    assert mo.run(outdir=str(wo)) > 0
    gdir = wd / "outdir_0001"
    names = sorted(f for f in os.listdir(wo) if f.endswith(".bas"))
    assert sum(f.startswith("ptr") for f in names) == 3
    for f in names:
        assert (gdir / f).read_bytes() == (wo / f).read_bytes(), f


def test_qg_exe_filter_event_files(gpu, tmp_path):
    """dtflt > 0: the `filter` event (qg.h:655-658) fires every dtflt and enters dtnext(); the output event writes
    pf%09d.bas = invertq(tmpl, qofl) (qg.c:124-129).  stdout cadence and every file bit-identical to the oracle."""
    from oracle import oracle as O
    N, nl = 64, 2
    wd = tmp_path / "gpu"; wd.mkdir()
    wo = tmp_path / "orc"; wo.mkdir()
    _write_params(str(wd / "params.in"), N, nl, tend=0.1, dtout=0.05, extra="dtflt = 0.03\nafilt = 4.\nediag = 0\n")
    psi = synth_psi(N, nl)
    O.lib().orc_write_bas(str(wd / "p0.bas").encode(), nl, N, 80., psi)
    exe = os.path.join(ROOT, "msom_b200", "lib", "qg.e")
    out = subprocess.run([exe], cwd=str(wd), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("Filter solution") == 3                    # t = 0.03, 0.06, 0.09
    po = O.Params(); O.lib().orc_default_params(po); O.lib().orc_read_params(str(wd / "params.in").encode(), po)
    mo = O.Model(po)
    p0 = np.zeros_like(psi); O.lib().orc_read_bas(str(wd / "p0.bas").encode(), nl, N, 80., p0)
    mo.set(O.PSI, p0); mo.L.orc_remove_mean_psi(mo.h); mo.set_const()
    assert mo.run(outdir=str(wo)) > 0
    gdir = wd / "outdir_0001"
    names = sorted(f for f in os.listdir(wo) if f.endswith(".bas"))
    assert sum(f.startswith("pf") for f in names) == 3 and sum(f.startswith("de_ft") for f in names) == 3
    for f in names:
        assert (gdir / f).read_bytes() == (wo / f).read_bytes(), f


def test_qg_exe_planetary_geostrophic_input_files(gpu, tmp_path):
    """The multiple-scale coupling inputs of set_const (msqg/qg.h:950-969): psipg_%dl_N%d.bas (large-scale stream
    function), frpg_%dl_N%d.bas (Froude number varying in x and y -> per-cell stretching) and rdpg_%dl_N%d.bas
    (deformation radius of the filter scale) are read from the working directory, backed up into the output directory
    (qg.h:806-820) and every output is bit-identical to the oracle's run() fed the same float32 files."""
    from oracle import oracle as O
    N, nl = 64, 3
    wd = tmp_path / "gpu"; wd.mkdir()
    wo = tmp_path / "orc"; wo.mkdir()
    _write_params(str(wd / "params.in"), N, nl, tend=0.1, dtout=0.05, extra="flsrv = 1\n")
    psi = synth_psi(N, nl)
    y, x = np.meshgrid((np.arange(N) + 0.5) / N, (np.arange(N) + 0.5) / N, indexing="ij")
    kw = base_kw(N, nl)
    fr = np.zeros_like(psi); ppg = np.zeros_like(psi)
    for l in range(nl):
        if l < nl - 1:
            fr[l] = kw["Fr"][l] * (1 + 0.25 * np.sin(2 * np.pi * x) * np.sin(np.pi * y))
        ppg[l] = 0.2 / (l + 1) * np.cos(np.pi * x) * np.sin(np.pi * y)
    rd = (1.0 + 0.5 * x * y)[None]
    names_in = {"p0.bas": (nl, psi), "frpg_%dl_N%d.bas" % (nl, N): (nl, fr), "psipg_%dl_N%d.bas" % (nl, N): (nl, ppg),
                "rdpg_%dl_N%d.bas" % (nl, N): (1, np.ascontiguousarray(rd))}
    for name, (nf, a) in names_in.items():
        O.lib().orc_write_bas(str(wd / name).encode(), nf, N, 80., np.ascontiguousarray(a))
    exe = os.path.join(ROOT, "msom_b200", "lib", "qg.e")
    out = subprocess.run([exe], cwd=str(wd), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    for name in names_in:
        if name != "p0.bas":
            assert "%s .. ok" % name in out.stdout
    po = O.Params(); O.lib().orc_default_params(po); O.lib().orc_read_params(str(wd / "params.in").encode(), po)
    mo = O.Model(po)

    def rd_file(name, nf):
        a = np.zeros((nf, N, N)); O.lib().orc_read_bas(str(wd / name).encode(), nf, N, 80., a); return a
    mo.set(O.PSI, rd_file("p0.bas", nl)); mo.L.orc_remove_mean_psi(mo.h)
    mo.set(O.FR, rd_file("frpg_%dl_N%d.bas" % (nl, N), nl))
    mo.set(O.PSIPG, rd_file("psipg_%dl_N%d.bas" % (nl, N), nl))
    mo.set(O.RD, rd_file("rdpg_%dl_N%d.bas" % (nl, N), 1))
    mo.set_const()
    assert mo.run(outdir=str(wo)) > 0
    gdir = wd / "outdir_0001"
    names = sorted(f for f in os.listdir(wo) if f.endswith(".bas"))
    assert len(names) >= 6
    for f in names:
        assert (gdir / f).read_bytes() == (wo / f).read_bytes(), f
    # backup_config: the float32 inputs come back byte for byte; sig_filt = min(afilt*Rd, Lfmax) follows the Rd file
    for name in names_in:
        if name != "p0.bas":
            assert (gdir / name).read_bytes() == (wd / name).read_bytes(), name
    sig = np.minimum(po.afilt * rd_file("rdpg_%dl_N%d.bas" % (nl, N), 1), po.Lfmax)      # qg.h:1062-1063
    O.lib().orc_write_bas(str(wo / "sig_filt.bas").encode(), 1, N, 80., np.ascontiguousarray(sig))
    assert (gdir / "sig_filt.bas").read_bytes() == (wo / "sig_filt.bas").read_bytes()
