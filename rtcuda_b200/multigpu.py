"""Sample-pass sharding of one image over the GPUs of a box (plumbing).

The reference is single-GPU (no cudaSetDevice / NCCL anywhere).
Every (pixel, sample) path is independent under the counter-based RNG, so rank
r renders samples [first, first+count) of every pixel into its own fp32
accumulation buffer (scene and BVH replicated); one sum all-reduce of the
3*W*H floats (NCCL over NVLink on GPUs, gloo in the CPU tests) followed by the
tonemap gives the image a single GPU would have produced with all samples.
"""
import ctypes as C


def shard_samples(total_spp, rank, world):
    """contiguous, balanced split of sample indices [0,total_spp) -> (first_sample, count)"""
    base, rem = divmod(total_spp, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def render_sharded(scene, cam, params, rank, world, accum_ptr, all_reduce=None, total_spp=None):
    """Render this rank's shard of `total_spp` samples into the buffer at accum_ptr (device memory
    of the scene's context, zeroed by the caller); `all_reduce()` is called after the local render.
    Returns (stats or None, first_sample, count)."""
    from . import capi
    total_spp = total_spp or params.spp
    first, count = shard_samples(total_spp, rank, world)
    st = None
    if count > 0:
        p = capi.RenderParams()
        C.memmove(C.byref(p), C.byref(params), C.sizeof(p))
        p.spp, p.first_sample, p.total_spp = count, first, total_spp
        st = scene.render_accumulate(cam, p, accum_ptr)
    if all_reduce is not None:
        all_reduce()
    return st, first, count
