"""Host-side logic of the multi-GPU path on CPU: process grids, tile boxes, and the world_size-2
rendezvous that carries the NCCL unique id (gloo backend; the halo traffic itself needs GPUs and is
checked by scripts/dist_check.py and tests/test_gpu_dist.py)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_grids_and_boxes_tile_the_domain():
    from msom_b200.dist import grid_for, tile_box
    for world in (1, 2, 4, 8):
        px, py = grid_for(world)
        assert px * py == world
        N = 64
        cover = np.zeros((N, N), dtype=int)
        for r in range(world):
            x0, y0, nx, ny = tile_box(N, px, py, r)
            assert (nx, ny) == (N // px, N // py)
            cover[y0:y0 + ny, x0:x0 + nx] += 1
        assert (cover == 1).all()
    assert grid_for(2) == (2, 1) and grid_for(4) == (2, 2) and grid_for(8) == (4, 2)


_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch.distributed as dist
from msom_b200.dist import broadcast_bytes, grid_for, tile_box
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
payload = bytes(range(128)) if rank == 0 else None
got = broadcast_bytes(payload, 128)
assert got == bytes(range(128)), (rank, got[:8])
px, py = grid_for(world)
box = tile_box(256, px, py, rank)
boxes = [None] * world
dist.all_gather_object(boxes, box)
assert sorted(boxes) == sorted(tile_box(256, px, py, r) for r in range(world))
dist.barrier()
if rank == 0:
    print("GLOO_OK", world, boxes)
dist.destroy_process_group()
"""


def test_unique_id_broadcast_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "GLOO_OK 2" in out.stdout


def test_reference_arm_runs_on_rank0_only(tmp_path):
    """bench.py --impl reference under torchrun: rank 0 prints the line, the other rank exits 0 without work."""
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--N", "4096", "--nl", "2"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    import json
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0
