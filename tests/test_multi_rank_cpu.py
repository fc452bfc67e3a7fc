"""The N > 1 path on CPU: world_size-2 `gloo` processes replicate the flat scene (broadcast), each renders
its interleaved tiles with the test-only host build of the kernel body, the tile buffers are all-gathered
and de-interleaved, and the result must equal the single-rank frame.  (On GPUs the same plumbing runs over
NCCL/NVLink with rt_render_tiles_device / rt_untile_device; tests/test_gpu_parity.py covers those.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from raytracer_js_b200 import parallel, scenes

from util import flat_of, hostsim_render, make_params

W, H = 100, 72  # ragged against the 16x16 tiles on purpose


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        flat = extra = None
        if rank == 0:
            b = scenes.random_spheres(600, 0.02, 0.1, seed=4.0, mix="mirrors", box_fraction=0.2)
            flat = flat_of(b)
            prm0 = make_params(flat, b)
            extra = {"sky": prm0.sky_texture, "sub": prm0.default_substance, "refmax": prm0.refmax}
        flat, extra, nbytes = parallel.broadcast_flat_scene(flat, 0, torch.device("cpu"), extra)
        assert nbytes > 0
        from raytracer_js_b200 import _native as N
        prm = N.Params()
        prm.refmax, prm.sky_texture, prm.default_substance = extra["refmax"], extra["sky"], extra["sub"]
        prm.distance_attenuation_factor, prm.n_frames, prm.rng_seed = 1.0, 2, 1.0
        cam = scenes.bench_camera(W, H)
        tiles, _, _ = hostsim_render(flat, cam, prm, n_threads=2, tile_rank=rank, tile_world=world)
        mine = torch.from_numpy(tiles.reshape(-1))
        gathered = torch.empty(world * mine.numel(), dtype=torch.float32)
        dist.all_gather_into_tensor(gathered, mine)
        frame = parallel.untile_numpy(gathered.numpy(), W, H, world)
        if rank == 0:
            full, _, _ = hostsim_render(flat, cam, prm, n_threads=2)
            np.save(os.path.join(out_dir, "ok.npy"), np.array([np.array_equal(frame, full), float(np.abs(full).sum() > 0)]))
    finally:
        dist.destroy_process_group()


def test_two_ranks_tiles_equal_single_rank(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(tmp_path / "ok.npy")
    assert ok[0] == 1.0 and ok[1] == 1.0


def test_pack_unpack_roundtrip():
    b = scenes.random_spheres(200, 0.02, 0.1, seed=4.0, mix="mirrors", textures=[scenes.checker_texture(32, 16)])
    flat = flat_of(b)
    meta, payload = parallel.pack_flat(flat)
    assert all(o % 16 == 0 for o in meta["offsets"])
    back = parallel.unpack_flat(meta, payload)
    assert sorted(back.arrays) == sorted(flat.arrays)
    for k, v in flat.arrays.items():
        np.testing.assert_array_equal(back.arrays[k], v)
        assert back.arrays[k].dtype == v.dtype


@pytest.mark.parametrize("w,h,world", [(1920, 1080, 8), (100, 72, 2), (16, 16, 4), (33, 17, 3)])
def test_tile_layout(w, h, world):
    tpr = parallel.tiles_per_rank(w, h, world)
    tiles_x, tiles_y = (w + 15) // 16, (h + 15) // 16
    assert tpr * world >= tiles_x * tiles_y > (tpr - 1) * world
    # a frame whose pixels hold their own index survives tile -> gather -> untile
    frame = np.arange(h * w * 3, dtype=np.float32).reshape(h, w, 3)
    g = np.zeros((world, tpr, 16, 16, 3), np.float32)
    for t in range(tiles_x * tiles_y):
        ty, tx = divmod(t, tiles_x)
        blk = frame[ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16]
        g[t % world, t // world, :blk.shape[0], :blk.shape[1]] = blk
    np.testing.assert_array_equal(parallel.untile_numpy(g.reshape(-1), w, h, world), frame)
