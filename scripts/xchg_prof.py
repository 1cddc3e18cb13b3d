"""torchrun script: per-exchange timing of the rb group (profiled, no graphs): count, ms, by kind (aux encodes the kind)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch, torch.distributed as dist
from common import base_kw, synth_psi
from msom_b200 import capi as G
from msom_b200.dist import nccl_group
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, nl = 4096, 4
agg = int(sys.argv[1]) if len(sys.argv) > 1 else 512
g = nccl_group(G.make_params(**base_kw(N, nl)), agg, local, smoother="rb")
g.set_global(G.PSI, synth_psi(N, nl)); g.set_const()
for _ in range(4):
    g.step()
for mode in ("graph", "profiled"):
    torch.cuda.synchronize(); dist.barrier()
    if mode == "profiled":
        g.profile(True)
    c0, x0, l0 = g.total_cycles, g.exchanges, g.launches
    g.timer_start()
    for _ in range(5):
        g.step()
    ms = g.timer_stop() / 5
    if rank == 0:
        print("%s transport=%s world=%d agg=%d: %.2f ms/step, cycles/step %.1f, exchanges/step %.1f, launches/step %.0f" % (
            mode, g.transport, world, agg, ms, (g.total_cycles - c0) / 5, (g.exchanges - x0) / 5, (g.launches - l0) / 5))
    if mode == "profiled":
        prof = g.profile_read(); g.profile(False)
        if rank == 0:
            for k, v in prof.items():
                print("   %-13s %8.3f ms/step  %6.1f scopes/step  aux/step %.0f" % (k, v["ms"] / 5, v["count"] / 5, v["aux"] / 5))
g.close()
dist.destroy_process_group()
