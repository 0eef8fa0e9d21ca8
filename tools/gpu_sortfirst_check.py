"""N-GPU check of the sort-first path (run under torchrun on a multi-GPU box):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 tools/gpu_sortfirst_check.py
Each rank renders its row strips on its own GPU, the owned rows are stored into rank 0's textures over IPC/NVLink peer
mappings, and rank 0 compares the composited frame with the golden hashes (rendered by the unmodified reference)."""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
import cases
from softlight_b200 import capi as A, scenes as S
from softlight_b200.sortfirst import SortFirst


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    os.environ.setdefault("SLCU_GEOM_SHARE_MIN_PRIMS", "0")  # the check shares every batch, however small
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl")
    frames = np.load(os.path.join(ROOT, "tests", "golden", "golden_frames.npz"))
    hashes = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_hashes.json")))["cases"]
    bad = 0
    for name, strip in (("c1_small", 8), ("c2_small", 8), ("rand_clip_heavy", 1), ("rand_overflow_wide", 2), ("c1_sphere100k", 8), ("c2_instancing10k", 8), ("c3_indoor2m", 8), ("c4_skinned100", 8),
                        ("points_dense", 1), ("points_sparse_1080p", 8), ("wire_alpha", 1), ("wire_overflow_premult", 2), ("wire_1080p", 8)):
        scene = (cases.CASES.get(name) or cases.BIG_CASES[name])()
        ctx = A.Context("cuda", device=local)
        ctx.set_rcp_table(frames["rcp_table"])
        sf = SortFirst(ctx, dist, rank, world, strip_rows=strip)
        if os.environ.get("SLCU_GEOM_SHARE", "1") != "0":  # the ranks split the vertex stage and route the records
            sf.share_geometry(int(scene.tris_per_frame * 1.3) * 246 + (64 << 20))
        inst = S.SceneInstance(scene, ctx)
        fc = sf.make_target(scene.color_type, scene.width, scene.height)
        fd = sf.make_target(scene.depth_type, scene.width, scene.height)
        for _ in range(2):  # twice: the second frame overwrites the first one's rows in place
            inst.render()
            sf.composite(inst.state["color"], fc)
            sf.composite(inst.state["depth"], fd)
            sf.frame_barrier()
        if rank == 0:
            color = ctx.texture_read(fc, scene.color_type, scene.width, scene.height)
            depth = ctx.texture_read(fd, scene.depth_type, scene.width, scene.height)[..., 0]
            ok = sha(color) == hashes[name]["color_sha256"] and sha(depth) == hashes[name]["depth_sha256"]
            print(f"{'ok ' if ok else 'BAD'} {name} world={world} strip={strip} binned={ctx.get_stats()['trisBinned']}", flush=True)
            bad += 0 if ok else 1
        dist.barrier()
        ctx.close()
    t = torch.tensor([bad], device=f"cuda:{local}")
    dist.broadcast(t, src=0)
    dist.destroy_process_group()
    if rank == 0:
        print("SORT-FIRST", "ALL MATCH" if int(t.item()) == 0 else "MISMATCH", flush=True)
    return int(t.item())


if __name__ == "__main__":
    sys.exit(main())
