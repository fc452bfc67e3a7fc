"""Multi-GPU plumbing (one process per GPU, torch.distributed): scene replication and the tile layout.

The path shards by pixels: the frame is cut into 16x16 tiles, tile t belongs to rank
t % world, every rank renders its tiles from a replica of the flat scene.  The only exchanges are (1) the
broadcast of the flat scene buffers from the rank that built them and (2) the gather of the tile-major
frame pieces; both go through torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .flatten import FlatScene

TILE_W = TILE_H = 16
TILE_PIXELS = TILE_W * TILE_H


def tiles_per_rank(width: int, height: int, world: int) -> int:
    n = ((width + TILE_W - 1) // TILE_W) * ((height + TILE_H - 1) // TILE_H)
    return (n + world - 1) // world


def pack_flat(flat: FlatScene):
    """-> (meta dict, uint8 payload) with every array 16-byte aligned inside the payload."""
    names = sorted(flat.arrays)
    arrs = [np.ascontiguousarray(flat.arrays[n]) for n in names]
    sizes = [a.nbytes for a in arrs]
    offs = np.concatenate([[0], np.cumsum([(s + 15) // 16 * 16 for s in sizes])]).astype(np.int64)
    payload = np.zeros(int(offs[-1]), np.uint8)
    for a, o, s in zip(arrs, offs, sizes):
        payload[o:o + s] = a.view(np.uint8).reshape(-1)
    meta = {"names": names, "dtypes": [str(a.dtype) for a in arrs], "shapes": [a.shape for a in arrs],
            "offsets": offs[:-1].tolist(), "sizes": sizes, "nbytes": int(offs[-1])}
    return meta, payload


def unpack_flat(meta: Dict, payload: np.ndarray) -> FlatScene:
    flat = FlatScene()
    for n, d, shp, o, s in zip(meta["names"], meta["dtypes"], meta["shapes"], meta["offsets"], meta["sizes"]):
        flat.arrays[n] = payload[o:o + s].view(np.dtype(d)).reshape(shp).copy()
    return flat


def broadcast_flat_scene(flat: Optional[FlatScene], src: int, device, extra: Optional[dict] = None):
    """Replicate the flat scene built on rank `src` to every rank.  Returns (FlatScene, extra, nbytes).
    The payload travels as one uint8 tensor on `device` (NCCL broadcast over NVLink when it is a GPU)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    box = [None]
    payload_np = None
    if rank == src:
        meta, payload_np = pack_flat(flat)
        box = [{"meta": meta, "extra": extra}]
    dist.broadcast_object_list(box, src=src)
    meta, extra = box[0]["meta"], box[0]["extra"]
    t = torch.empty(meta["nbytes"], dtype=torch.uint8, device=device)
    if rank == src:
        t.copy_(torch.from_numpy(payload_np))
    dist.broadcast(t, src=src)
    if rank != src:
        flat = unpack_flat(meta, t.cpu().numpy())
    return flat, extra, meta["nbytes"]


def untile_numpy(gathered: np.ndarray, width: int, height: int, world: int, channels: int = 3) -> np.ndarray:
    """Host reference of rt_untile_device: [world][tiles_per_rank][256][channels] -> [height][width][channels]."""
    tpr = tiles_per_rank(width, height, world)
    g = gathered.reshape(world, tpr, TILE_H, TILE_W, channels)
    tiles_x = (width + TILE_W - 1) // TILE_W
    tiles_y = (height + TILE_H - 1) // TILE_H
    out = np.zeros((tiles_y * TILE_H, tiles_x * TILE_W, channels), gathered.dtype)
    for t in range(tiles_x * tiles_y):
        ty, tx = divmod(t, tiles_x)
        out[ty * TILE_H:(ty + 1) * TILE_H, tx * TILE_W:(tx + 1) * TILE_W] = g[t % world, t // world]
    return out[:height, :width]


class _DevArray:
    """Wraps a raw device pointer so that torch.as_tensor(...) can view it (CUDA array interface)."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 3}


class PeerFrame:
    """The multi-GPU frame without a gather step: ONE frame buffer lives on rank `dst`; every rank maps it
    through CUDA IPC (rt_peer_open) and its render kernels store their tiles into it across NVLink as they
    are produced (rt_render_shard_device).  `barrier()` closes the frame (rt_peer_barrier: system-scope
    flags in peer memory, one tiny kernel per rank on the ctx stream).  Handles travel through
    torch.distributed (all_gather_object); one process per GPU."""

    def __init__(self, lib, ctx, rank: int, world: int, n_floats: int, dst: int = 0):
        import ctypes as C

        import torch.distributed as dist

        from . import _native as N
        self.lib, self.ctx, self.rank, self.world, self.dst, self.n_floats = lib, ctx, rank, world, dst, n_floats
        self.epoch = 0
        self._own, self._mapped = [], []
        h = C.create_string_buffer(64)
        flags_ptr = C.c_void_p()
        N.check(ctx, lib.rt_peer_alloc(ctx, 4 * max(world, 1), C.byref(flags_ptr), h))
        self._own.append(flags_ptr)
        mine = {"flags": bytes(h.raw)}
        frame_ptr = None
        if rank == dst:
            frame_ptr = C.c_void_p()
            N.check(ctx, lib.rt_peer_alloc(ctx, 4 * n_floats, C.byref(frame_ptr), h))
            self._own.append(frame_ptr)
            mine["frame"] = bytes(h.raw)
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        self.flag_ptrs = (C.c_void_p * world)()
        for r in range(world):
            if r == rank:
                self.flag_ptrs[r] = flags_ptr.value
            else:
                p = C.c_void_p()
                N.check(ctx, lib.rt_peer_open(ctx, everyone[r]["flags"], C.byref(p)))
                self._mapped.append(p)
                self.flag_ptrs[r] = p.value
        if rank == dst:
            self.frame_ptr = frame_ptr.value
        else:
            p = C.c_void_p()
            N.check(ctx, lib.rt_peer_open(ctx, everyone[dst]["frame"], C.byref(p)))
            self._mapped.append(p)
            self.frame_ptr = p.value
        dist.barrier()

    def barrier(self) -> None:
        from . import _native as N
        self.epoch += 1
        N.check(self.ctx, self.lib.rt_peer_barrier(self.ctx, self.rank, self.world, self.flag_ptrs, self.epoch))

    def tensor(self):
        """torch view of the frame (only meaningful on rank dst, where the memory is local)."""
        import torch
        return torch.as_tensor(_DevArray(self.frame_ptr, self.n_floats), device="cuda")

    def close(self) -> None:
        import torch.distributed as dist

        from . import _native as N
        N.check(self.ctx, self.lib.rt_synchronize(self.ctx))
        dist.barrier()  # nobody unmaps while a peer may still be storing
        for p in self._mapped:
            N.check(self.ctx, self.lib.rt_peer_close(self.ctx, p))
        dist.barrier()
        for p in self._own:
            N.check(self.ctx, self.lib.rt_peer_free(self.ctx, p))
        self._mapped, self._own = [], []
