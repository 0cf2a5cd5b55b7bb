"""Convert the Stanford bunny (ASCII PLY, the reference's only fixture:
/root/reference/bun_zipper.ply, read by main.cu:60-62) into the compact
binary mesh rtcuda_b200/data/bunny.rtbm ("RTBM", float32 vertices, int32
faces) using the library's own PLY reader, so the mesh travels to the GPU box
where /root/reference does not exist.

    python tools/make_bunny_fixture.py [/root/reference/bun_zipper.ply]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtcuda_b200 import capi  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/bun_zipper.ply"
libpath = capi.DEFAULT_LIB if os.path.exists(capi.DEFAULT_LIB) else os.path.join(ROOT, "tests/emu/librtb_emu.so")
L = capi.Lib(libpath)
verts, faces = L.load_mesh(src)
print(len(verts), "vertices", len(faces), "faces")
v = np.ascontiguousarray(verts, np.float32)
f = np.ascontiguousarray(faces, np.int32)
L.check(L.lib.rtb_mesh_save_bin(capi.BUNNY_BIN.encode(), v.ctypes.data_as(C.c_void_p), C.c_int64(len(v)),
                                f.ctypes.data_as(C.c_void_p), C.c_int64(len(f))))
v2, f2 = L.load_mesh(capi.BUNNY_BIN)
assert (v2 == v).all() and (f2 == f).all()
print("wrote", capi.BUNNY_BIN, os.path.getsize(capi.BUNNY_BIN), "bytes")
