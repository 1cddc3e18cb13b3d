"""How much does a bulk PCIe transfer on another stream slow the msqg step down?  (diagnostic for the pipelined field I/O)
python scripts/io_overlap.py [N] [nl]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from msom_b200 import capi as G
import bench as B

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 4
torch.cuda.set_device(0)
m = G.Model(G.make_params(**B.workload_kw(N, nl)), 0)
m.set_smoother("rb")
st = torch.cuda.Stream()
m.set_stream(st.cuda_stream)
m.set(G.PSI, B.workload_psi(N, nl)); m.set_const()
for _ in range(5):
    m.step()
host = torch.empty((nl, N, N), dtype=torch.float64).pin_memory()
host2 = torch.empty((nl, N, N), dtype=torch.float64).pin_memory()
dev = torch.empty((nl, N, N), dtype=torch.float64, device="cuda")
dev2 = torch.empty((nl, N, N), dtype=torch.float64, device="cuda")
up, down = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, steps=6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        if h2d:
            with torch.cuda.stream(up):
                dev.copy_(host, non_blocking=True)
        if d2h:
            with torch.cuda.stream(down):
                host2.copy_(dev2, non_blocking=True)
        m.step()
    t1 = time.perf_counter()          # the steps are done (every step ends with a stream synchronisation)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) / steps * 1e3, (t2 - t0) / steps * 1e3


for name, a, b in (("step alone", 0, 0), ("+ H2D stream", 1, 0), ("+ D2H stream", 0, 1), ("+ both", 1, 1), ("step alone", 0, 0)):
    s, tot = run(a, b)
    print("%-14s step %.2f ms   step+transfers %.2f ms" % (name, s, tot), flush=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.stream(up):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize(); t1 = time.perf_counter()
with torch.cuda.stream(down):
    host2.copy_(dev2, non_blocking=True)
torch.cuda.synchronize(); t2 = time.perf_counter()
with torch.cuda.stream(up):
    dev.copy_(host, non_blocking=True)
with torch.cuda.stream(down):
    host2.copy_(dev2, non_blocking=True)
torch.cuda.synchronize(); t3 = time.perf_counter()
gb = host.numel() * 8 / 1e9
print("H2D alone %.1f GB/s, D2H alone %.1f GB/s, both at once %.1f + %.1f GB/s" % (gb / (t1 - t0), gb / (t2 - t1), gb / (t3 - t2), gb / (t3 - t2)))
