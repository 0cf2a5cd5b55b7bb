"""Several GPUs behind the C ABI (rtb_multi_*, include/rtb.h): the sample passes of one render() call
(render.cuh:366-367) split over the GPUs of a box, per-GPU accumulation, one sum-reduction, tonemap on the first GPU.

Host build (tests/emu): the members are emulated devices in one address space, which checks the sharding, the
replication of a built scene, the masks, the reduction plumbing and — with RTB_RENDER_DETERMINISTIC — that the image is
bit-identical for ANY number of GPUs.  `-m gpu`: the same through NCCL on however many B200s the box has.
"""
import ctypes as C

import numpy as np
import pytest

from rtcuda_b200 import capi
from conftest import mean_rel_err


@pytest.fixture(scope="module", params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def L(request):
    return request.getfixturevalue(request.param)


def device_count(L, want):
    """emulated devices are free; on a GPU box use what is there (one GPU still runs the whole path, without a peer)"""
    if "emu" in L.lib._name:
        return want
    import torch
    return max(1, min(want, torch.cuda.device_count()))


@pytest.fixture(scope="module")
def s1(L, bunny):
    return L.host_scene(capi.RTB_SCENE_S1, *bunny)


def test_split_render_equals_the_single_gpu_render(L, s1):
    n = device_count(L, 3)
    cam = s1.camera(1.5)
    p = capi.render_params(L, width=96, height=64, spp=7, max_bounces=6)  # 7 samples over 3 GPUs: shares 3 / 2 / 2
    single, st1 = L.context(0).scene(s1.desc).render(cam, p)
    m = capi.Multi(L, list(range(n)))
    ms = m.scene(s1.desc)
    img, st = ms.render(cam, p)
    assert st.paths == st1.paths == 96 * 64 * 7
    assert abs(int(st.extend_rays) - int(st1.extend_rays)) <= 2 and abs(int(st.shadow_rays) - int(st1.shadow_rays)) <= 2
    assert mean_rel_err(img, single) <= 1e-5  # float sums in another order
    # fewer samples than GPUs: some members render nothing
    p1 = capi.render_params(L, width=96, height=64, spp=1, max_bounces=6)
    a, _ = ms.render(cam, p1)
    b, _ = L.context(0).scene(s1.desc).render(cam, p1)
    assert mean_rel_err(a, b) <= 1e-6
    ms.close(); m.close()


def test_deterministic_mode_is_bit_identical_for_any_number_of_gpus(L, s1):
    cam = s1.camera(1.0)
    p = capi.render_params(L, width=80, height=80, spp=6, max_bounces=8, flags=capi.RTB_RENDER_DETERMINISTIC)
    ref, _ = L.context(0).scene(s1.desc).render(cam, p)
    plain, _ = L.context(0).scene(s1.desc).render(cam, capi.render_params(L, width=80, height=80, spp=6, max_bounces=8))
    assert mean_rel_err(ref, plain) <= 1e-6  # 2^-28 fixed point against float sums
    again, _ = L.context(0).scene(s1.desc).render(cam, p)
    assert (ref.view(np.uint32) == again.view(np.uint32)).all()  # run to run
    for n in sorted({1, device_count(L, 2), device_count(L, 3), device_count(L, 4)}):
        m = capi.Multi(L, list(range(n)))
        ms = m.scene(s1.desc)
        img, _ = ms.render(cam, p)
        assert (img.view(np.uint32) == ref.view(np.uint32)).all(), f"{n} GPUs: image differs from the single-GPU image"
        ms.close(); m.close()
    # and for any split into calls: two sample passes into one fixed-point buffer
    # (covered for the float path by test_parity.py::test_sample_pass_sharding_is_additive)


def test_replicated_scene_and_device_masks(L, s1):
    n = device_count(L, 3)
    m = capi.Multi(L, list(range(n)))
    primary = m.context(0).scene(s1.desc)
    ms = m.replicate(primary)
    cam = s1.camera(1.0)
    p = capi.render_params(L, width=64, height=64, spp=4, max_bounces=5, flags=capi.RTB_RENDER_DETERMINISTIC)
    ref, _ = primary.render(cam, p)
    img, _ = ms.render(cam, p)
    assert (img.view(np.uint32) == ref.view(np.uint32)).all()  # every GPU traverses a copy of the same tree
    for mask in range(1, 1 << n):
        q = capi.render_params(L, width=64, height=64, spp=4, max_bounces=5, flags=capi.RTB_RENDER_DETERMINISTIC, device_mask=mask)
        im, st = ms.render(cam, q)
        assert (im.view(np.uint32) == ref.view(np.uint32)).all(), mask
        assert st.paths == 64 * 64 * 4
    bad = capi.render_params(L, width=64, height=64, spp=4, device_mask=1 << n)
    out = np.zeros((64, 64, 3), np.float32)
    assert L.lib.rtb_multi_render(ms.h, C.byref(cam), C.byref(bad), out.ctypes.data_as(C.c_void_p), None) == -1
    ms.close(); primary.close(); m.close()


def test_one_call_entry_and_bad_arguments(L, s1):
    n = device_count(L, 2)
    cam = s1.camera(1.0)
    p = capi.render_params(L, width=48, height=48, spp=4, max_bounces=4)
    out = np.zeros((48, 48, 3), np.float32)
    devs = (C.c_int32 * n)(*range(n))
    st = capi.RenderStats()
    L.check(L.lib.rtb_render_multi(devs, n, C.byref(s1.desc), None, C.byref(cam), C.byref(p), out.ctypes.data_as(C.c_void_p), C.byref(st)))
    ref, _ = L.context(0).scene(s1.desc).render(cam, p)
    assert mean_rel_err(out, ref) <= 1e-5 and st.paths == 48 * 48 * 4
    h = C.c_void_p()
    assert L.lib.rtb_multi_create(None, 2, C.byref(h)) == -1
    assert L.lib.rtb_multi_create((C.c_int32 * 2)(0, 0), 2, C.byref(h)) == -1  # the same device twice
    assert L.lib.rtb_multi_create((C.c_int32 * 1)(0), 0, C.byref(h)) == -1


def test_instanced_scene_on_several_gpus(L, bunny):
    n = device_count(L, 2)
    hs = L.host_scene_instanced(capi.RTB_SCENE_S2, *bunny, grid=2)
    cam = hs.camera(1.0)
    p = capi.render_params(L, width=64, height=64, spp=4, max_bounces=4, flags=capi.RTB_RENDER_DETERMINISTIC)
    ref, _ = L.context(0).scene(hs.idesc).render(cam, p)
    m = capi.Multi(L, list(range(n)))
    ms = m.scene(hs.idesc)
    img, _ = ms.render(cam, p)
    assert (img.view(np.uint32) == ref.view(np.uint32)).all()
    ms.close(); m.close()


@pytest.mark.gpu
def test_library_owned_communicator_single_rank(gpu):
    """rtb_comm_* with a world of one (the N > 1 path runs under torchrun in bench.py --gpus N)"""
    import torch
    ctx = gpu.context(0)
    comm = capi.Comm(ctx, capi.Comm.unique_id(gpu), 0, 1)
    x = torch.arange(1000, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    comm.allreduce_f32(x.data_ptr(), x.numel())
    assert (x.cpu() == torch.arange(1000, dtype=torch.float32)).all()
    y = torch.arange(1000, dtype=torch.int64, device="cuda") << 40
    torch.cuda.synchronize()
    comm.allreduce_i64(y.data_ptr(), y.numel())
    assert (y.cpu() == torch.arange(1000, dtype=torch.int64) << 40).all()
    comm.close()
