"""Decode submit on the bench workload, per-family device times (profile pass). No validation (for timing experiments)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from jpeg_detection_resnet_ssd_b200 import _lib
_lib.LIB_PATH = os.environ.get('SSDC_LIB_AB', _lib.LIB_PATH)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
bias = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
y, cands, enc = bench.make_workload(B, 64, bias, seed=1234, pinned=False)
ctx = _lib.get_context(); lib = ctx.lib
d_y = ctx.dev_alloc(y.nbytes)
_lib.check(lib.ssdc_memcpy_h2d(ctx.handle, 0, d_y, _lib.ptr(y), y.nbytes))
p = bench.decode_params(_lib)
A = y.shape[1]
for _ in range(3):
    _lib.check(lib.ssdc_decode_submit(ctx.handle, d_y, _lib.F32, 1, B, A, 21, _lib.C.byref(p)))
ctx.synchronize()
steps = 20
ctx.timer_start()
for _ in range(steps):
    _lib.check(lib.ssdc_decode_submit(ctx.handle, d_y, _lib.F32, 1, B, A, 21, _lib.C.byref(p)))
ms = ctx.timer_stop()
ctx.profile_enable(True)
for _ in range(5):
    _lib.check(lib.ssdc_decode_submit(ctx.handle, d_y, _lib.F32, 1, B, A, 21, _lib.C.byref(p)))
prof = ctx.profile_read(); ctx.profile_enable(False)
d1 = prof['decode_filter'][0] / prof['decode_filter'][1]
print('step %.4f ms  D1 %.4f ms  = %.0f GB/s (%.1f%% of 6544.7)  cands/img %.0f' % (ms / steps, d1, y.nbytes / d1 / 1e6, y.nbytes / d1 / 1e6 / 65.447, cands),
      {k: round(v[0] / 5, 4) for k, v in prof.items() if v[1]})
