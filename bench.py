#!/usr/bin/env python
"""bench.py — the driver's benchmark contract for the render hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

One "step" = one full render of the workload (generate .. extend/shadow .. shade
.. accumulate .. tonemap) on synthetic input: the reference's default scene
(Stanford bunny in a Cornell box, main.cu:41-166) at BASELINE.json configs[1]
"bunny 1920x1080, 64 spp, max depth 8, diffuse + area light, fixed RNG seed".
N > 1 (torchrun, one rank per GPU): weak scaling by sample pass — every rank
renders its own 64 samples of every pixel ([64r, 64r+64) of 64N), scene
replicated, per-GPU accumulation buffers summed with one NCCL all-reduce inside
the timed step, then tonemapped.

Prints ONE JSON line (rank 0).  `value` = Mrays/s (extend + shadow rays
actually traversed by all ranks) / device time; `e2e` = same metric through the
C ABI from pinned HOST buffers: rtb_scene_create (H2D + GPU BVH build) +
rtb_render (device->host framebuffer) per step.

--impl reference times the reference itself: its own CUDA build
(oracle/_ref/ref_harness, compiled from /root/reference where it lies) on one
B200 — the reference is CUDA-only and has no CPU path (BASELINE.json) — and
falls back to the scalar host-C++ oracle port on the host cores when that
binary is absent.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene kind, grid, width, height, spp per GPU, max depth, description)
    "c1": (1, 0, 600, 600, 10, 10, "C1 main.cu default: bunny+Cornell box 69,463 tris, 600x600, 10 spp, depth 10"),
    "c2": (1, 0, 1920, 1080, 64, 8, "C2 bunny+Cornell box 69,463 tris, 1920x1080, 64 spp, depth 8, matte + 2 area-light tris, seed 1"),
    "c3": (3, 12, 3840, 2160, 16, 8, "C3 144-bunny field 10,000,956 tris, 3840x2160, 16 spp, depth 8"),
    "c4": (2, 0, 1920, 1080, 64, 16, "C4 bunny+Cornell box, matte/mirror/glass round-robin, 1920x1080, 64 spp, depth 16 + RR"),
    # C5 is STRONG scaling: 1024 spp in total, split evenly over the ranks (the spp entry is the total)
    "c5": (3, 12, 3840, 2160, 1024, 8, "C5 144-bunny field 10,000,956 tris, 3840x2160, 1024 spp in total (sample passes split over the GPUs), depth 8"),
}
# the same scenes given as INSTANCES (rtb_scene_create_instanced, two-level BVH): identical flattened triangles,
# 4 MB of nodes + triangles instead of 600 MB; results equal the flat scene's up to the rounding of the ray transform
WORKLOADS["c3i"] = WORKLOADS["c3"][:6] + ("C3 as instances: 144 placements of one 69,451-triangle bunny + Cornell shell (two-level BVH; flattens to the 10,000,956 triangles of C3), 3840x2160, 16 spp, depth 8",)
WORKLOADS["c5i"] = WORKLOADS["c5"][:6] + ("C5 as instances (two-level BVH; flattens to the 10,000,956 triangles of C5), 3840x2160, 1024 spp in total, depth 8",)
INSTANCED = {"c3i", "c5i"}
STRONG = {"c5", "c5i"}
METRIC = "Mrays/s (extend+shadow)"


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def load_scene(L, capi, kind, grid, instanced=False):
    verts, faces = L.load_mesh()
    if instanced:
        return L.host_scene_instanced(kind, verts, faces, grid=grid)
    return L.host_scene(kind, verts, faces, grid=grid)


def cpu_baseline_port(capi, L, hs, cam, params, budget_s=12.0, threads=None):
    """the oracle (scalar host-C++ restatement of the reference) on the host cores, bounded sample"""
    from oracle import binding
    threads = threads or os.cpu_count() or 1
    orc = binding.Oracle().scene(hs.desc)
    npix = params.width * params.height
    # sample = evenly spaced blocks of pixels (same spp / depth), sized from a calibration run
    calib = min(npix, 1024)
    start = (npix // 2) - calib // 2
    t0 = time.perf_counter()
    _, _, st = orc.render(cam, params, start, start + calib, threads=threads)
    dt = max(time.perf_counter() - t0, 1e-6)
    n = int(min(npix, max(calib, calib * budget_s / dt)))
    start = (npix - n) // 2
    t0 = time.perf_counter()
    _, _, st = orc.render(cam, params, start, start + n, threads=threads)
    dt = time.perf_counter() - t0
    rays = st[1] + st[2]
    return {"value": rays / dt * 1e-6, "unit": "Mrays/s", "cores": threads, "kind": "port",
            "sample": f"{n} of {npix} pixels (centre block) x {params.spp} spp, depth {params.max_bounces}: "
                      f"{rays} rays in {dt:.2f} s", "seconds": dt, "rays": int(rays)}


def run_reference(args, rank):
    if rank != 0:
        return 0
    from rtcuda_b200 import capi
    from oracle import binding
    kind, grid, W, H, spp, depth, desc_txt = WORKLOADS[args.workload]
    if args.workload in STRONG:
        # the reference's camera_ray_end_id is an int (render.cuh:371): 4K x 1024 spp overflows it; C5 is extrapolated
        # from the reference's ms/spp on C3
        print(json.dumps({"impl": "reference", "unavailable": "the reference cannot run 3840x2160x1024 spp in one call (int overflow, render.cuh:371); use --workload c3"}))
        return 0
    emu_or_cuda = capi.DEFAULT_LIB
    L = capi.Lib(emu_or_cuda)  # host-side scene code only; no GPU call is made on this arm
    hs = load_scene(L, capi, kind, grid)
    cam = hs.camera(W / H)
    params = capi.render_params(L, width=W, height=H, spp=spp, max_bounces=depth)
    line = {"impl": "reference", "metric": METRIC, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": desc_txt}}
    harness = binding.REF_HARNESS
    done = False
    if os.path.exists(harness) and not args.reference_cpu:
        try:
            with tempfile.TemporaryDirectory() as td:
                scene_file = os.path.join(td, "scene.rtbs")
                hs.save(scene_file)
                passes = args.warmup + args.steps

                def run(cmd):
                    out = subprocess.run([harness, scene_file] + [str(c) for c in cmd], capture_output=True, text=True, timeout=3000)
                    if out.returncode != 0:
                        raise RuntimeError(out.stderr[-500:])
                    return [json.loads(l[5:]) for l in out.stdout.splitlines() if l.startswith("JSON ")]
                js = run(["loop", W, H, spp, depth, passes, 1])
                scene_js, loop_js = js[0], js[-1]
                ms = loop_js["pass_ms"][args.warmup:]
                rays = loop_js["pass_rays"][args.warmup:]
                value = sum(rays) / sum(ms) * 1e-3
                js2 = run(["render", W, H, spp, depth, args.steps, args.warmup])
                render_js = js2[-1]
                e2e_ms = render_js["ms_per_call"] + js2[0]["bvh_build_ms"]
                line.update({"value": value, "ms_per_step": sum(ms) / len(ms),
                             "reference_kind": "reference CUDA build (unmodified kernels, nvcc -O3 sm_100a) on 1 B200; "
                                               "loop time excludes its cudaMallocs and RNG init",
                             "rays_per_step": rays[0], "ms_per_spp": sum(ms) / len(ms) / spp,
                             "reference_iterations": loop_js["iterations"] // passes,
                             "reference_bvh_build_ms_host": scene_js["bvh_build_ms"],
                             "e2e": {"value": rays[0] / e2e_ms * 1e-3, "unit": "Mrays/s", "h2d_bytes_per_step": 0,
                                     "d2h_bytes_per_step": 0,
                                     "note": "Bvh::Bvh host build + unmodified render() call incl. its mallocs and framebuffer copy"}})
                done = True
        except Exception as e:  # harness unusable on this box: fall through to the CPU port
            line["reference_harness_error"] = str(e)[-300:]
    cb = cpu_baseline_port(capi, L, hs, cam, params, budget_s=15.0)
    line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if not done:
        line.update({"value": cb["value"], "ms_per_step": cb["seconds"] * 1e3,
                     "reference_kind": "oracle port (scalar host C++ restatement) on host cores",
                     "e2e": {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--reference-cpu", action="store_true", help="reference arm: force the CPU oracle port")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--flags", type=int, default=0)
    args = ap.parse_args()
    # W >= 3 (timing contract); the multi-second C5 steps may run with fewer to fit a GPU lease
    args.warmup = max(args.warmup, 3) if args.impl == "ours" and args.workload not in STRONG else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist
    from rtcuda_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference --reference-cpu for the oracle)")
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    kind, grid, W, H, spp, depth, desc_txt = WORKLOADS[args.workload]
    strong = args.workload in STRONG
    L = capi.Lib()
    ctx = L.context(local_rank)
    instanced = args.workload in INSTANCED
    hs = load_scene(L, capi, kind, grid, instanced)
    cam = hs.camera(W / H)
    if strong:
        if spp % world:
            raise SystemExit(f"bench.py: {spp} spp do not split evenly over {world} ranks")
        total_spp, spp = spp, spp // world
    else:
        total_spp = spp * world
    p = capi.render_params(L, width=W, height=H, spp=spp, max_bounces=depth, first_sample=rank * spp,
                           total_spp=total_spp, pool_size=args.pool, flags=args.flags)
    # pinned host copies of the scene arrays: the e2e leg uploads from these every step
    arr = hs.arrays()
    pin = {k: torch.from_numpy(np.ascontiguousarray(arr[k])).pin_memory() for k in ("vertices", "material_ids", "light_ids")}
    pdesc = capi.SceneDesc()
    C.memmove(C.byref(pdesc), C.byref(hs.desc), C.sizeof(pdesc))
    pdesc.vertices = pin["vertices"].data_ptr(); pdesc.material_ids = pin["material_ids"].data_ptr(); pdesc.light_ids = pin["light_ids"].data_ptr()
    h2d = sum(t.numel() * t.element_size() for t in pin.values()) + hs.desc.num_materials * 20 + hs.desc.num_lights * 40
    if instanced:  # the same pinned geometry, placed by the instance table
        idesc = capi.InstancedSceneDesc()
        C.memmove(C.byref(idesc), C.byref(hs.idesc), C.sizeof(idesc))
        idesc.geometry = pdesc
        h2d += hs.idesc.num_instances * C.sizeof(capi.Instance) + (hs.idesc.num_meshes + 1) * 8
        pdesc = idesc
    nfl = 3 * W * H
    host_img = torch.empty(nfl, dtype=torch.float32).pin_memory()

    scene = ctx.scene(pdesc)
    bst = scene.stats()
    accum = torch.zeros(nfl, dtype=torch.float32, device="cuda")
    out = torch.empty_like(accum)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    seg_events = []  # (after render, after all-reduce) of the timed steps: where a multi-GPU step spends its time

    def step(timed=False):
        accum.zero_()
        st = scene.render_accumulate(cam, p, accum.data_ptr())
        if timed and world > 1:
            e1 = torch.cuda.Event(enable_timing=True); e1.record()
        if world > 1:
            dist.all_reduce(accum)
        if timed and world > 1:
            e2 = torch.cuda.Event(enable_timing=True); e2.record()
            seg_events.append((e1, e2))
        ctx.tonemap_device(accum.data_ptr(), nfl, total_spp, out.data_ptr())
        return st

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # nvidia-smi needs ~0.1-0.3 s before its first sample and a timed region can be that short: keep the GPU under the
    # same load (untimed steps, every rank alike) until the sampler is running, so that the samples are taken under load
    t_spin = time.perf_counter()
    step()
    torch.cuda.synchronize()
    n_extra = torch.tensor([max(0, min(32, int(0.6 / max(time.perf_counter() - t_spin, 1e-3))))], device="cuda")
    if world > 1:
        dist.broadcast(n_extra, src=0)  # every rank runs the same number of steps (each holds an all-reduce)
    for _ in range(int(n_extra.item())):
        step()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stats = []
    for k in range(args.steps):
        flush.zero_()  # evict L2 between timed steps
        torch.cuda.synchronize()
        ev[k][0].record()
        stats.append(step(timed=True))
        ev[k][1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    allreduce_ms = sum(a.elapsed_time(b) for a, b in seg_events) / max(len(seg_events), 1) if seg_events else 0.0
    render_ms = sum(ev[k][0].elapsed_time(seg_events[k][0]) for k in range(len(seg_events))) / max(len(seg_events), 1) if seg_events else 0.0
    per_rank = None
    if world > 1:  # every rank's own render / all-reduce time (the all-reduce time includes waiting for the slowest rank)
        mine = torch.tensor([render_ms, allreduce_ms], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"render_ms": [round(float(t[0]), 2) for t in allr], "allreduce_ms": [round(float(t[1]), 2) for t in allr]}
    t_total = torch.tensor([sum(ms_steps)], dtype=torch.float64, device="cuda")
    rays_total = torch.tensor([float(sum(s.extend_rays + s.shadow_rays for s in stats))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays_total, op=dist.ReduceOp.SUM)
    ms_per_step = t_total.item() / args.steps
    value = rays_total.item() / t_total.item() * 1e-3  # rays / ms -> Mrays/s
    launches = int(sum(s.kernel_launches for s in stats) + args.steps)  # + tonemap per step

    # ---- end to end through the C ABI from pinned host buffers ----
    def e2e_step():
        sc2 = ctx.scene(pdesc)  # rtb_scene_create: H2D of the triangle soup + GPU BVH build
        if world == 1:
            st = capi.RenderStats()
            L.check(L.lib.rtb_render(sc2.h, C.byref(cam), C.byref(p), C.c_void_p(host_img.data_ptr()), C.byref(st)))
        else:
            accum.zero_()
            st = sc2.render_accumulate(cam, p, accum.data_ptr())
            dist.all_reduce(accum)
            ctx.tonemap_device(accum.data_ptr(), nfl, total_spp, out.data_ptr())
            host_img.copy_(out)
            torch.cuda.synchronize()
        sc2.close()
        return st
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_rays = 0
    n_e2e = max(1, min(args.steps, 5))
    for _ in range(n_e2e):
        s = e2e_step()
        e2e_rays += s.extend_rays + s.shadow_rays
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    e2e_r = torch.tensor([float(e2e_rays)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_r, op=dist.ReduceOp.SUM)

    if rank == 0:
        # ---- roofline of the dominant kernel, measured live ----
        # k_trace<3>: the extend AND shadow rays of one iteration in one persistent launch (stats.fused_trace);
        # its duration is ms_extend (CUDA events on the render stream inside rtb_render_accumulate).
        # The timed steps run four (or two) concurrent wavefronts (streams), whose kernels overlap; the per-kernel durations
        # come from extra steps with RTB_RENDER_SINGLE_PIPELINE (one stream, the kernel timed alone).
        ps = capi.render_params(L, width=W, height=H, spp=spp, max_bounces=depth, first_sample=rank * spp,
                                total_spp=total_spp, pool_size=args.pool, flags=args.flags | capi.RTB_RENDER_SINGLE_PIPELINE)
        s0 = []
        for _ in range(max(1, min(args.steps, 3))):
            flush.zero_()
            accum.zero_()
            s0.append(scene.render_accumulate(cam, ps, accum.data_ptr()))
        fused = all(s.fused_trace for s in s0)
        tr_ms = sum(s.ms_extend + s.ms_shadow for s in s0); sh_ms = sum(s.ms_shade for s in s0); tot_ms = sum(s.ms_total for s in s0)
        tr_launches = sum(s.extend_launches for s in s0)
        ext_rays = sum(s.extend_rays for s in s0); shd_rays = sum(s.shadow_rays for s in s0)
        pc = capi.render_params(L, width=W, height=H, spp=max(1, min(2, spp)), max_bounces=depth, first_sample=rank * spp,
                                total_spp=total_spp, flags=capi.RTB_RENDER_COUNT_WORK)
        accum.zero_()
        cst = scene.render_accumulate(cam, pc, accum.data_ptr())
        e_nodes = cst.extend_nodes / max(cst.extend_rays, 1); e_tris = cst.extend_tris / max(cst.extend_rays, 1)
        s_nodes = cst.shadow_nodes / max(cst.shadow_rays, 1); s_tris = cst.shadow_tris / max(cst.shadow_rays, 1)
        hit_frac = cst.hits / max(cst.extend_rays, 1)
        # algorithmic bytes: 80 B per node fetched, 48 B per triangle tested; an extend ray reads its
        # 32 B record (origin|pixel, dir|sample) and, when it hits, 16 B (beta) and writes the 48 B hit record; a
        # shadow ray reads 32 B (origin|tmax, dir|excluded) and, unoccluded, 16 B (radiance|pixel) + 12 B splat
        bytes_extend = e_nodes * 80 + e_tris * 48 + 32 + 64 * hit_frac
        bytes_shadow = s_nodes * 80 + s_tris * 48 + 32 + 28
        bytes_per_launch = (bytes_extend * ext_rays + bytes_shadow * shd_rays) / max(tr_launches, 1)
        avg_launch_ms = tr_ms / max(tr_launches, 1)
        achieved = bytes_per_launch / (avg_launch_ms * 1e-3) * 1e-9 if avg_launch_ms > 0 else 0.0
        peak, peak_src = hbm_peak()
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                per_ray = json.load(f).get(args.workload, {}).get("k_trace_dram_bytes_per_ray")
                if per_ray is not None:  # ncu DRAM bytes per ray of this kernel x the rays one launch of this run processes
                    traffic = per_ray * (ext_rays + shd_rays) / max(tr_launches, 1)
        except Exception:
            pass
        # second roofline for an L2-resident scene: ncu L2 bytes per ray of this kernel against the measured L2 read bandwidth
        l2 = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                l2_per_ray = json.load(f).get(args.workload, {}).get("k_trace_l2_bytes_per_ray")
            with open(os.path.join(ROOT, "profiles", "l2_bandwidth.json")) as f:
                l2_peak = float(json.load(f)["l2_read_gbs"])
            if l2_per_ray is not None and avg_launch_ms > 0:
                l2_ach = l2_per_ray * (ext_rays + shd_rays) / max(tr_launches, 1) / (avg_launch_ms * 1e-3) * 1e-9
                l2 = {"achieved": l2_ach, "peak": l2_peak, "unit": "GB/s", "frac": l2_ach / l2_peak,
                      "note": "ncu lts__t_bytes per ray x rays per launch / launch duration; peak = tools/l2_bandwidth.cu on this pool's B200 (profiles/l2_bandwidth.json)"}
        except Exception:
            pass
        # SURVEY 8d floor model: one root-to-leaf descent of a BVH8 with <= 4 triangles per leaf, 80 B nodes, 48 B
        # triangles, 48 B of ray I/O: ceil(log8(n / 4)) * 80 + 4 * 48 + 48 bytes per ray, all of it from HBM
        floor = None
        try:
            import math
            n_flat = int(bst.num_flat_triangles) or int(bst.num_triangles)
            floor_bytes = math.ceil(math.log(max(n_flat / 4.0, 8.0), 8)) * 80 + 4 * 48 + 48
            rays_per_s = (ext_rays + shd_rays) / (tr_ms * 1e-3) if tr_ms > 0 else 0.0
            floor_rate = peak * 1e9 / floor_bytes
            floor = {"bytes_per_ray": floor_bytes, "rays_per_s_at_peak": floor_rate, "kernel_rays_per_s": rays_per_s,
                     "frac": rays_per_s / floor_rate,
                     "note": "rays per second of the trace kernel while it runs (single-pipeline steps) against the HBM peak divided by the "
                             "floor-model bytes per ray; above 1 when the scene is served by L1 / L2 and a ray needs fewer bytes than the model"}
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": "k_trace<3> (extend + shadow rays, one persistent launch per iteration)" if fused else "k_trace<1> + k_trace<2>",
                    "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_extend_ray": bytes_extend, "algorithmic_bytes_per_shadow_ray": bytes_shadow,
                    "extend_nodes_per_ray": e_nodes, "extend_tris_per_ray": e_tris, "shadow_nodes_per_ray": s_nodes,
                    "shadow_tris_per_ray": s_tris, "hit_fraction": hit_frac,
                    "avg_launch_ms": avg_launch_ms, "launches": int(tr_launches), "timed_with": "single pipeline (kernel alone on one stream), %d steps" % len(s0),
                    "single_pipeline_ms_per_step": tot_ms / len(s0),
                    "kernel_share_of_step": tr_ms / tot_ms if tot_ms else None,
                    "shade_share_of_step": sh_ms / tot_ms if tot_ms else None, "l2": l2, "floor_model": floor,
                    "note": "scene (%.1f MB nodes+triangles) %s; node/triangle counts from the counting kernel variant"
                            % ((bst.node_bytes + bst.triangle_bytes) / 1e6,
                               "is L2-resident: the kernel is bound by instruction issue, not by HBM (see profiles/)" if bst.node_bytes + bst.triangle_bytes < 100e6
                               else "exceeds L2")}
        pool_used = min(int(p.pool_size) or int(os.environ.get("RTB_POOL", 1 << 25)), int(stats[0].paths))
        line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": desc_txt, "spp_per_gpu": spp, "total_spp": total_spp,
                           "sharding": "sample pass per rank, scene replicated, NCCL all-reduce of the accumulation buffer" if world > 1 else "single GPU",
                           "pool_size": pool_used,
                           "l2": "256 MB device memset between timed steps (L2 flush); the ray / hit queues (%.1f GB) are streamed every iteration"
                                 % (pool_used * 288 / 1e9)},
                "ms_per_spp": ms_per_step / total_spp * (1 if strong else world), "paths_per_step": int(stats[0].paths) * world,
                "rays_per_step": rays_total.item() / args.steps,
                "iterations_per_step": int(stats[0].iterations), "pipelines": int(stats[0].pipelines),
                "per_rank": per_rank, "bvh_build_ms": bst.build_ms, "bvh_nodes": int(bst.num_nodes),
                "bvh_sah": bst.sah_cost,
                "e2e": {"value": e2e_r.item() / e2e_t.item() * 1e-6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(nfl * 4), "steps": n_e2e, "ms_per_step": e2e_t.item() / n_e2e * 1e3,
                        "note": "rtb_scene_create (pinned H2D + GPU BVH build) + render + device->host framebuffer, every step"},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline}
        if instanced:
            line["config"]["instances"] = int(bst.num_instances)
            line["config"]["stored_triangles"] = int(bst.num_triangles)
            line["config"]["flattened_triangles"] = int(bst.num_flat_triangles)
        if world == 1 and not args.no_cpu_baseline and not instanced:  # (the oracle would need the 10 M flattened triangles)
            cb = cpu_baseline_port(capi, L, hs, cam, capi.render_params(L, width=W, height=H, spp=spp, max_bounces=depth))
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            try:  # the same port on ONE host thread as well (smaller sample)
                c1 = cpu_baseline_port(capi, L, hs, cam, capi.render_params(L, width=W, height=H, spp=spp, max_bounces=depth), budget_s=4.0, threads=1)
                line["cpu_baseline"]["one_thread"] = {"value": c1["value"], "unit": c1["unit"], "cores": 1, "sample": c1["sample"]}
            except Exception as e:  # (a reported extra, never a reason to lose the line)
                line["cpu_baseline"]["one_thread"] = {"error": str(e)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
