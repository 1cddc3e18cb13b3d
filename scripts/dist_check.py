"""torchrun --nproc-per-node N scripts/dist_check.py : the NCCL back-end of the domain decomposition
against the CPU oracle's emulation of the same decomposition, bit for bit, on every rank."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch, torch.distributed as dist
from common import base_kw, synth_psi
from oracle import oracle as O
from msom_b200 import capi as G
from msom_b200.dist import nccl_group, grid_for

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, nl, agg_n = 256, 3, 64
sm = sys.argv[1] if len(sys.argv) > 1 else "lex"
if sm == "rb":
    agg_n = 128 if world > 2 else 64
kw = base_kw(N, nl)
periodic = sm == "rbper"   # doubly periodic domain (sbc = -1) on the process grid: neighbours wrap around
if periodic:
    sm = "rb"
    kw = base_kw(N, nl, sbc=-1.)
px, py = grid_for(world)
g = nccl_group(G.make_params(**kw), agg_n, local, smoother=sm)
psi = synth_psi(N, nl)
if periodic:
    from test_gpu_periodic import periodic_psi
    psi = periodic_psi(N, nl)
g.set_global(G.PSI, psi); g.set_const()
mo = O.Model(O.make_params(**kw)); mo.L.orc_set_decomp(mo.h, px, py, agg_n); mo.set_smoother(sm)
mo.set(O.PSI, psi); mo.set_const()
ok = True
for s in range(3):
    ok &= (g.step() == mo.step())
_, _, x0, y0, nx, ny = g.boxes[0]
for fg, fo in ((G.PSI, O.PSI), (G.Q, O.Q)):
    ok &= bool(np.array_equal(g.get_tile(0, fg), mo.get(fo)[:, y0:y0 + ny, x0:x0 + nx]))
ok &= (g.total_cycles == mo.L.orc_total_cycles(mo.h))
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST_CHECK", "PASS" if int(t.item()) == 1 else "FAIL", sm + ("-periodic" if periodic else ""), g.transport, "world", world, "grid %dx%d" % (px, py), "exchanges", g.exchanges)
g.close()
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
