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
