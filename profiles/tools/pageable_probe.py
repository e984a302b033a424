"""Wall clock of decode_detections on pageable / pinned host batches of several sizes (diagnosis of the staging path)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import synth
from jpeg_detection_resnet_ssd_b200 import _lib
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_output_decoder import decode_detections

enc = synth.make_encoder(SSDInputEncoder, 'ssd300')
base = synth.synth_y_pred(synth.anchors_of(enc), enc.variances, 21, 8, 5, bg_bias=8.0, hot=40)
print('cpus', len(os.sched_getaffinity(0)), 'SSDC_STAGE_THREADS', os.environ.get('SSDC_STAGE_THREADS'))
for B in (8, 32, 128):
    y = np.ascontiguousarray(np.tile(base, (B // 8, 1, 1)))
    yp = _lib.pinned_empty(y.shape, y.dtype); yp[...] = y
    for name, arr in (('pageable', y), ('pinned', yp)):
        for _ in range(3):
            decode_detections(arr, 0.01, 0.45, 200, 'centroids', True, 300, 300)
        n = 30
        t0 = time.perf_counter()
        for _ in range(n):
            decode_detections(arr, 0.01, 0.45, 200, 'centroids', True, 300, 300)
        dt = (time.perf_counter() - t0) / n
        print('B=%d %s: %.3f ms per call, %.1f GB/s' % (B, name, dt * 1e3, y.nbytes / dt / 1e9))
    t0 = time.perf_counter()
    for _ in range(20):
        z = y.copy()
    print('   numpy copy of the batch: %.3f ms' % ((time.perf_counter() - t0) / 20 * 1e3))
