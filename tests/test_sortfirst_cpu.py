"""CPU tests of the sort-first host logic: the ownership rule, the composite restated on the host, and a
world_size-2 gloo run in which each rank renders with the CPU oracle, keeps only the rows it owns and the frame is
gathered and merged on rank 0 (the N>1 code path of the harness without a GPU)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from softlight_b200 import sortfirst as SF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("height,world,strip", [(1, 2, 8), (77, 2, 8), (1080, 8, 8), (2160, 4, 1), (4320, 8, 16), (150, 3, 5)])
def test_ownership_is_an_exact_partition(height, world, strip):
    masks = [SF.owned_rows(height, r, world, strip) for r in range(world)]
    total = np.sum(masks, axis=0)
    assert np.all(total == 1), "every row has exactly one owner"
    # interleave: consecutive strips go to consecutive ranks (mirrors the reference's scan-line interleave)
    for y in range(height):
        assert masks[(y // strip) % world][y]


def test_merge_owned_rows_rebuilds_the_frame():
    rng = np.random.default_rng(3)
    full = rng.integers(0, 255, size=(77, 33, 4), dtype=np.uint8)
    frames = []
    for r in range(3):
        f = np.full_like(full, 0xEE)
        m = SF.owned_rows(77, r, 3, 8)
        f[m] = full[m]
        frames.append(f)
    assert np.array_equal(SF.merge_owned_rows(frames, 8), full)


_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["SL_ROOT"]); sys.path.insert(0, os.path.join(os.environ["SL_ROOT"], "tests"))
import cases
from softlight_b200 import sortfirst as SF

dist.init_process_group(backend="gloo")
rank, world = dist.get_rank(), dist.get_world_size()
frames = np.load(os.path.join(os.environ["SL_ROOT"], "tests", "golden", "golden_frames.npz"))
ok = True
for name in ("c1_small", "rand_clip_heavy"):
    scene = cases.CASES[name]()
    color, depth, _ = cases.render("port", scene, rcp_table=frames["rcp_table"])
    mine = SF.owned_rows(scene.height, rank, world, 8)
    kept = np.where(mine[:, None, None], color, np.uint8(0xEE))        # only owned rows are valid on a rank
    t = torch.from_numpy(kept.copy())
    gathered = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, gathered, dst=0)
    if rank == 0:
        merged = SF.merge_owned_rows([g.numpy() for g in gathered], 8)
        ok = ok and np.array_equal(merged, frames[name + ".color"])
flag = torch.tensor([1 if ok else 0])
dist.broadcast(flag, src=0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
"""


def test_two_rank_gloo_composite_matches_golden(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, SL_ROOT=ROOT, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
