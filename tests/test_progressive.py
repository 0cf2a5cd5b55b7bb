"""Progressive rendering and checkpoints (rtb_accum_*, SURVEY 5 "checkpoint / resume", 8f-4): sample passes added
call by call, the image resolved after any pass, the state saved to a file and resumed.  The reference renders once
and writes once (main.cu:173-192); what it computes for N samples is what N samples added in any number of passes
must give."""
import ctypes as C

import numpy as np
import pytest

from rtcuda_b200 import capi
from conftest import mean_rel_err


@pytest.fixture(scope="module", params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def L(request):
    return request.getfixturevalue(request.param)


@pytest.fixture(scope="module")
def setup(L, bunny):
    hs = L.host_scene(capi.RTB_SCENE_S1, *bunny)
    ctx = L.context(0)
    return hs, ctx, ctx.scene(hs.desc), hs.camera(1.25)


@pytest.mark.parametrize("deterministic", [False, True])
def test_passes_add_up_and_resume_from_a_checkpoint(L, setup, tmp_path, deterministic):
    hs, ctx, sc, cam = setup
    W, H = 80, 64
    flags = capi.RTB_RENDER_DETERMINISTIC if deterministic else 0
    whole, _ = sc.render(cam, capi.render_params(L, width=W, height=H, spp=9, max_bounces=6, flags=flags))
    acc = capi.Accum(ctx, W, H, deterministic)
    assert acc.samples == 0
    out = np.zeros((H, W, 3), np.float32)
    assert L.lib.rtb_accum_resolve(acc.h, out.ctypes.data_as(C.c_void_p)) == -1  # nothing to resolve yet
    images = []
    for spp in (2, 3):
        acc.add(sc, cam, capi.render_params(L, width=W, height=H, spp=spp, max_bounces=6))
        images.append(acc.resolve(W, H))
    assert acc.samples == 5
    first5, _ = sc.render(cam, capi.render_params(L, width=W, height=H, spp=5, max_bounces=6, flags=flags))
    assert mean_rel_err(images[-1], first5) <= 1e-6  # progressive output: the image so far IS the 5-sample render
    ck = str(tmp_path / "render.rtba")
    acc.save(ck)
    acc.close()
    # "another day": a new context, a new scene object, the buffer from the file
    ctx2 = L.context(0)
    sc2 = ctx2.scene(hs.desc)
    acc2 = capi.Accum(ctx2, path=ck)
    assert acc2.samples == 5
    acc2.add(sc2, cam, capi.render_params(L, width=W, height=H, spp=4, max_bounces=6))
    final = acc2.resolve(W, H)
    assert acc2.samples == 9
    if deterministic:
        assert (final.view(np.uint32) == whole.view(np.uint32)).all()  # bit-identical to the uninterrupted render
    else:
        assert mean_rel_err(final, whole) <= 1e-6
    acc2.close(); sc2.close()


def test_checkpoint_files_are_validated(L, setup, tmp_path):
    hs, ctx, sc, cam = setup
    bad = tmp_path / "bad.rtba"
    h = C.c_void_p()
    bad.write_bytes(b"RTBA" + bytes(20))
    assert L.lib.rtb_accum_load(ctx.h, str(bad).encode(), C.byref(h)) == -4
    import struct
    bad.write_bytes(struct.pack("<IIiiii", 0x41425452, 1, 1 << 20, 1 << 20, 3, 0))  # 2^40 pixels in a 24-byte file
    assert L.lib.rtb_accum_load(ctx.h, str(bad).encode(), C.byref(h)) == -4
    assert L.lib.rtb_accum_load(ctx.h, b"/nonexistent/x.rtba", C.byref(h)) == -4
    acc = capi.Accum(ctx, 32, 32)
    p = capi.render_params(L, width=64, height=32, spp=1)
    assert L.lib.rtb_accum_add_samples(acc.h, sc.h, C.byref(cam), C.byref(p), None) == -1  # size mismatch
    assert L.lib.rtb_accum_create(ctx.h, 0, 5, 0, C.byref(h)) == -1
