"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline and
--impl reference legs) import this module; the product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_HARNESS = os.path.join(HERE, "_ref", "ref_harness")


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


class Oracle:
    def __init__(self):
        if not os.path.exists(LIB):
            build()
        self.lib = C.CDLL(LIB)
        self.lib.orc_scene_create.restype = C.c_void_p

    def scene(self, desc):
        return OracleScene(self, desc)

    def tri_intersect(self, verts9, ray):
        v = (C.c_float * 9)(*verts9)
        out = (C.c_float * 3)()
        r = np.ascontiguousarray(ray)
        hit = self.lib.orc_tri_intersect(v, r.ctypes.data_as(C.c_void_p), out)
        return bool(hit), tuple(out)

    def offset_ray_origin(self, p, n):
        a = (C.c_float * 3)(*p); b = (C.c_float * 3)(*n); o = (C.c_float * 3)()
        self.lib.orc_offset_ray_origin(a, b, o)
        return np.array(list(o), np.float32)

    def rand4(self, seed, pixel, sample, block):
        o = (C.c_float * 4)()
        self.lib.orc_rand4(C.c_uint32(seed), C.c_uint32(pixel), C.c_uint32(sample), C.c_uint32(block), o)
        return list(o)

    def sample_f(self, material, wo, n, u1, u2):
        f = (C.c_float * 3)(); no = (C.c_float * 3)(); wi = (C.c_float * 3)(); pdf = C.c_float()
        self.lib.orc_sample_f(C.byref(material), (C.c_float * 3)(*wo), (C.c_float * 3)(*n), C.c_float(u1), C.c_float(u2),
                              f, no, wi, C.byref(pdf))
        return np.array(list(f)), np.array(list(no)), np.array(list(wi)), pdf.value

    def sample_li_area(self, verts9, p, u1, u2):
        wi = (C.c_float * 3)(); t = C.c_float(); pdf = C.c_float()
        self.lib.orc_sample_li_area((C.c_float * 9)(*verts9), (C.c_float * 3)(*p), C.c_float(u1), C.c_float(u2), wi, C.byref(t), C.byref(pdf))
        return np.array(list(wi), np.float32), t.value, pdf.value

    def camera_look_at(self, cam_struct_type, lookfrom, lookat, up, vfov, aspect):
        cam = cam_struct_type()
        self.lib.orc_camera_look_at((C.c_float * 3)(*lookfrom), (C.c_float * 3)(*lookat), (C.c_float * 3)(*up),
                                    C.c_float(vfov), C.c_float(aspect), C.byref(cam))
        return cam


class OracleScene:
    def __init__(self, orc, desc):
        self.lib = orc.lib
        self.h = C.c_void_p(self.lib.orc_scene_create(C.byref(desc)))

    def bvh_stats(self):
        a, b = C.c_int(), C.c_int()
        self.lib.orc_bvh_stats(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def trace_closest(self, rays, hit_dtype, threads=0):
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(len(rays), dtype=hit_dtype)
        self.lib.orc_trace_closest(self.h, rays.ctypes.data_as(C.c_void_p), C.c_int64(len(rays)),
                                   hits.ctypes.data_as(C.c_void_p), threads or os.cpu_count())
        return hits

    def trace_any(self, rays, excluded=None, threads=0):
        rays = np.ascontiguousarray(rays)
        occ = np.zeros(len(rays), dtype=np.uint8)
        ex = None
        if excluded is not None:
            excluded = np.ascontiguousarray(excluded, dtype=np.int32)
            ex = excluded.ctypes.data_as(C.c_void_p)
        self.lib.orc_trace_any(self.h, rays.ctypes.data_as(C.c_void_p), ex, C.c_int64(len(rays)),
                               occ.ctypes.data_as(C.c_void_p), threads or os.cpu_count())
        return occ

    def trace_counts(self, rays):
        rays = np.ascontiguousarray(rays)
        a, b = C.c_double(), C.c_double()
        self.lib.orc_trace_counts(self.h, rays.ctypes.data_as(C.c_void_p), C.c_int64(len(rays)), C.byref(a), C.byref(b))
        return a.value, b.value

    def render(self, cam, params, pixel_begin=0, pixel_end=-1, threads=0, want_accum=False):
        """returns (rgb[h,w,3] gamma-encoded, accum or None, (paths, extend_rays, shadow_rays))"""
        h, w = params.height, params.width
        rgb = np.zeros((h, w, 3), np.float32)
        acc = np.zeros((h, w, 3), np.float32) if want_accum else None
        st = (C.c_uint64 * 3)()
        self.lib.orc_render(self.h, C.byref(cam), C.byref(params), C.c_int64(pixel_begin), C.c_int64(pixel_end),
                            rgb.ctypes.data_as(C.c_void_p), acc.ctypes.data_as(C.c_void_p) if want_accum else None,
                            st, threads or os.cpu_count())
        return rgb, acc, tuple(st)

    def __del__(self):
        try:
            self.lib.orc_scene_destroy(self.h)
        except Exception:
            pass
