"""rtcuda_b200 — B200-native wavefront path tracer behind lashhw/rtcuda's
scene / primitive / material / camera / light interfaces.

The product is the CUDA library `librtb.so` (C ABI: include/rtb.h) built from
rtcuda_b200/csrc.  This package only binds it for tests and bench.py.
"""
from . import capi  # noqa: F401
