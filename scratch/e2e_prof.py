"""Where the end-to-end decode time goes: pinned H2D alone vs the public call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from jpeg_detection_resnet_ssd_b200 import _lib
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_output_decoder import decode_detections
B = 1024
y, cands, enc = bench.make_workload(B, 64, 8.0, seed=1234, pinned=True)
ctx = _lib.get_context(); lib = ctx.lib
d_y = ctx.dev_alloc(y.nbytes)
for _ in range(2): _lib.check(lib.ssdc_memcpy_h2d(ctx.handle, 0, d_y, _lib.ptr(y), y.nbytes))
t0 = time.perf_counter()
for _ in range(5): _lib.check(lib.ssdc_memcpy_h2d(ctx.handle, 0, d_y, _lib.ptr(y), y.nbytes))
t_h2d = (time.perf_counter() - t0) / 5
for _ in range(2): out = decode_detections(y, 0.01, 0.45, 200, 'centroids', True, 300, 300)
t0 = time.perf_counter()
for _ in range(5): out = decode_detections(y, 0.01, 0.45, 200, 'centroids', True, 300, 300)
t_all = (time.perf_counter() - t0) / 5
p = bench.decode_params(_lib)
t0 = time.perf_counter()
for _ in range(5):
    rows, counts, idx = _lib._decode_once(ctx, y, _lib.F32, p)
t_once = (time.perf_counter() - t0) / 5
print('H2D alone %.2f ms (%.1f GB/s); _decode_once %.2f ms; decode_detections %.2f ms (%.0f img/s)' % (t_h2d * 1e3, y.nbytes / t_h2d / 1e9, t_once * 1e3, t_all * 1e3, B / t_all))
