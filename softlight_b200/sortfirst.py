"""Sort-first plumbing above the C ABI (one process per GPU, torch.distributed for rendezvous only).

The data path is `slcu_composite`: each rank stores the rows it owns straight into the root's texture over an NVLink
peer mapping.  torch.distributed carries the 64-byte IPC handles once at set-up and provides the frame barrier.
`owned_rows` / `merge_owned_rows` restate the ownership rule on the host for tests and for the gloo (CPU) path.
"""
from __future__ import annotations

import numpy as np

DEFAULT_STRIP_ROWS = 8


def owned_rows(height: int, rank: int, world: int, strip_rows: int = DEFAULT_STRIP_ROWS) -> np.ndarray:
    """Boolean mask of the rows rank `rank` owns: (y // strip_rows) % world == rank (slcu_set_partition)."""
    y = np.arange(height)
    return (y // strip_rows) % world == rank


def merge_owned_rows(frames, strip_rows: int = DEFAULT_STRIP_ROWS) -> np.ndarray:
    """Host restatement of the composite: frames[r] is rank r's full-size attachment; take each row from its owner."""
    world = len(frames)
    out = np.empty_like(frames[0])
    for r, f in enumerate(frames):
        m = owned_rows(f.shape[0], r, world, strip_rows)
        out[m] = f[m]
    return out


class SortFirst:
    """Binds one context to its rank: partition, composite targets on the root, handle exchange."""

    def __init__(self, ctx, dist, rank: int, world: int, strip_rows: int = DEFAULT_STRIP_ROWS, root: int = 0, group=None):
        self.ctx, self.dist, self.rank, self.world, self.root, self.group = ctx, dist, rank, world, root, group
        self.strip_rows = strip_rows
        ctx.set_partition(rank, world, strip_rows)

    def make_target(self, color_type: int, w: int, h: int) -> int:
        """Creates the composite destination (same id on every rank; only the root's copy receives rows)."""
        tex = self.ctx.create_texture(color_type, w, h, 1)
        handle = self.ctx.texture_ipc_export(tex) if self.rank == self.root else None
        box = [handle]
        self.dist.broadcast_object_list(box, src=self.root, group=self.group)
        if self.rank != self.root:
            self.ctx.texture_ipc_import(tex, self.root, box[0])
        return tex

    def share_geometry(self, arena_bytes: int):
        """The ranks split the vertex stage and route triangle records to the owners of the rows they touch
        (slcu_geometry_share_*): every rank creates an arena, the handles travel once through torch.distributed."""
        self.ctx.geometry_share_create(arena_bytes)
        handles = [None] * self.world
        self.dist.all_gather_object(handles, self.ctx.geometry_share_export(), group=self.group)
        for r, h in enumerate(handles):
            if r != self.rank:
                self.ctx.geometry_share_import(r, h)
        self.dist.barrier(group=self.group)

    def composite(self, src_tex: int, dst_tex: int):
        self.ctx.composite(src_tex, dst_tex, self.root)

    def frame_barrier(self):
        """The frame is complete on the root once every rank's composite stores have retired."""
        self.ctx.sync()
        self.dist.barrier(group=self.group)
