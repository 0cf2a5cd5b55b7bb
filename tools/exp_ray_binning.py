"""Experiment (profiles/r2/r2_exp_ray_binning_session21.log): how much faster do the incoherent bounce rays of the
10 M-triangle scene trace when they are binned by origin cell (Morton code of the origin) and direction octant first?
Primary rays of the 3840x2160 camera, then cosine-weighted bounces from the hit points (torch on the GPU, only to make
the rays); every set is traced by rtb_trace_closest_device in its natural order, randomly permuted and sorted by seven keys."""
import ctypes as C, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtcuda_b200 import capi
L = capi.Lib()
grid = int(os.environ.get("GRID", "12"))
hs = L.host_scene(3, *L.load_mesh(), grid=grid)
w, h = 3840, 2160
cam = hs.camera(w / h)
ctx = L.context(0)
sc = ctx.scene(hs.desc)
dev = torch.device("cuda:0")
V = torch.from_numpy(hs.arrays()["vertices"].copy()).to(dev)  # (n, 9)
N = torch.linalg.cross(V[:, 3:6] - V[:, 0:3], V[:, 6:9] - V[:, 0:3])
N = N / N.norm(dim=1, keepdim=True).clamp_min(1e-30)
rays_h = L.primary_rays(cam, w, h)
sel = np.arange(0, w * h, 2)
rays = torch.from_numpy(rays_h[sel].view(np.float32).reshape(-1, 7).copy()).to(dev)
n = rays.shape[0]
def trace(r, reps=3):
    r = r.contiguous()
    hits = torch.empty((r.shape[0], 4), dtype=torch.float32, device=dev)
    best = 1e9
    for _ in range(reps):
        ms = C.c_float()
        L.check(L.lib.rtb_trace_closest_device(sc.h, C.c_void_p(r.data_ptr()), C.c_int64(r.shape[0]), C.c_void_p(hits.data_ptr()), C.byref(ms)))
        best = min(best, ms.value)
    return hits, best
def morton(p, lo, hi, bits=10):
    q = ((p - lo) / (hi - lo) * ((1 << bits) - 1)).clamp(0, (1 << bits) - 1).to(torch.int64)
    def spread(x):
        x = (x | (x << 16)) & 0x030000FF
        x = (x | (x << 8)) & 0x0300F00F
        x = (x | (x << 4)) & 0x030C30C3
        x = (x | (x << 2)) & 0x09249249
        return x
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
g = torch.Generator(device=dev); g.manual_seed(1)
lo = V.view(-1, 3).min(0).values; hi = V.view(-1, 3).max(0).values
cur = rays
for bounce in range(0, 5):
    hits, ms_nat = trace(cur)
    m = cur.shape[0]
    print(f"bounce {bounce}: {m} rays, natural order {ms_nat:.3f} ms = {m / ms_nat / 1e3:.0f} Mrays/s", flush=True)
    if bounce > 0:
        perm = torch.randperm(m, device=dev, generator=g)
        _, ms_rand = trace(cur[perm])
        key = morton(cur[:, 0:3], lo, hi)
        octant = ((cur[:, 3] > 0).to(torch.int64) | ((cur[:, 4] > 0).to(torch.int64) << 1) | ((cur[:, 5] > 0).to(torch.int64) << 2))
        out = [f"random {ms_rand:.3f}"]
        for name, k in (("morton30", key), ("cell8bit", key >> 22), ("cell11bit", key >> 19), ("cell14bit", key >> 16), ("cell8+oct", ((key >> 22) << 3) | octant),
                        ("cell11+oct", ((key >> 19) << 3) | octant), ("oct+morton30", (octant << 30) | key)):
            order = torch.sort(k, stable=True).indices
            _, ms_s = trace(cur[order])
            out.append(f"{name} {ms_s:.3f}")
        print("   " + "  ".join(out), flush=True)
    # next bounce: cosine-weighted direction about the normal facing the ray, from the hit point
    prim = hits[:, 3].view(torch.int32).to(torch.int64)
    ok = prim >= 0
    cur, hits, prim = cur[ok], hits[ok], prim[ok]
    P = cur[:, 0:3] + hits[:, 0:1] * cur[:, 3:6]
    nn = N[prim]
    nn = torch.where((nn * cur[:, 3:6]).sum(1, keepdim=True) > 0, -nn, nn)
    s = torch.randn((cur.shape[0], 3), device=dev, generator=g)
    s = s / s.norm(dim=1, keepdim=True)
    d = nn + s
    d = d / d.norm(dim=1, keepdim=True).clamp_min(1e-20)
    o = P + 1e-4 * nn
    cur = torch.cat([o, d, torch.full((cur.shape[0], 1), 3.0e38, device=dev)], dim=1)
