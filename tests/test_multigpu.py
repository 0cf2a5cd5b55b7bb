"""N>1 path on CPU: world_size-2 gloo run of the sample-pass sharding used by
bench.py --gpus N (rtcuda_b200/multigpu.py).  Each rank renders its shard with
the host build of the kernels (tests/emu), the accumulation buffers are summed
with torch.distributed all_reduce, and the result must equal a single-rank
render of all samples."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rtcuda_b200 import capi, multigpu
from conftest import build_emu, mean_rel_err

W, H, SPP, DEPTH = 48, 40, 6, 6


def test_shard_samples_partition():
    for total in (1, 5, 64, 1024):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                first, count = multigpu.shard_samples(total, r, world)
                got += list(range(first, first + count))
            assert got == list(range(total))
            counts = [multigpu.shard_samples(total, r, world)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1


def _scene(L, instanced):
    """the default scene, flat or as two instances (bunny placed by its transform + static shell: two-level BVH)"""
    if instanced:
        hs = L.host_scene_instanced(capi.RTB_SCENE_S1, *L.load_mesh())
        return hs, L.context(0).scene(hs.idesc)
    hs = L.host_scene(capi.RTB_SCENE_S1, *L.load_mesh())
    return hs, L.context(0).scene(hs.desc)


def _worker(rank, world, port, lib_path, out_path, instanced):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = capi.Lib(lib_path)
    hs, sc = _scene(L, instanced)
    cam = hs.camera(W / H)
    p = capi.render_params(L, width=W, height=H, spp=SPP, max_bounces=DEPTH)
    accum = torch.zeros(3 * W * H, dtype=torch.float32)
    st, first, count = multigpu.render_sharded(sc, cam, p, rank, world, accum.data_ptr(),
                                               all_reduce=lambda: dist.all_reduce(accum), total_spp=SPP)
    rays = torch.tensor([float(st.extend_rays + st.shadow_rays)], dtype=torch.float64)
    dist.all_reduce(rays)
    if rank == 0:
        out = torch.empty_like(accum)
        L.context(0).tonemap_device(accum.data_ptr(), accum.numel(), SPP, out.data_ptr())
        np.savez(out_path, img=out.numpy().reshape(H, W, 3), rays=rays.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,instanced", [(2, False), (3, False), (2, True)])
def test_two_rank_gloo_render_equals_single_rank(tmp_path, world, instanced):
    lib_path = build_emu()
    out_path = str(tmp_path / "img.npz")
    port = 29500 + (os.getpid() % 2000) + world + (10 if instanced else 0)
    mp.spawn(_worker, args=(world, port, lib_path, out_path, instanced), nprocs=world, join=True)
    got = np.load(out_path)
    L = capi.Lib(lib_path)
    hs, sc = _scene(L, instanced)
    ref, st = sc.render(hs.camera(W / H), capi.render_params(L, width=W, height=H, spp=SPP, max_bounces=DEPTH))
    assert got["rays"][0] == st.extend_rays + st.shadow_rays
    assert mean_rel_err(got["img"], ref) <= 1e-5
