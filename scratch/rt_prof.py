"""Round trip encode -> decode_detections_fast on the device (bench extra), per-family kernel times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from jpeg_detection_resnet_ssd_b200 import _lib, synth
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
enc = synth.make_encoder(SSDInputEncoder, 'ssd300')
ctx, h = enc._encoder(); lib = ctx.lib
gt = synth.synth_ground_truth(300, 300, 20, B, seed=78)
flat, offs = synth.flatten_ground_truth(gt)
A = 8732
d_enc = ctx.dev_alloc(B * A * 33 * 8)
pf = _lib.DecodeParams()
pf.mode, pf.input_coords, pf.normalize, pf.border_pixels = _lib.MODE_FAST, 0, 1, 0
pf.top_k, pf.nms_cap, pf.log_wh, pf.do_nms = 0, 0, 1, 1
pf.conf_thresh, pf.iou_thresh, pf.img_h, pf.img_w = 0.5, 0.45, 300.0, 300.0
def rt():
    _lib.check(lib.ssdc_encode(h, _lib.ptr(flat), _lib.ptr(offs), B, 1, d_enc, None, None))
    _lib.check(lib.ssdc_decode_submit(ctx.handle, d_enc, _lib.F64, 1, B, A, 21, _lib.C.byref(pf)))
for _ in range(3): rt()
ctx.synchronize()
ctx.timer_start()
for _ in range(10): rt()
ms = ctx.timer_stop() / 10
ctx.profile_enable(True)
for _ in range(3): rt()
prof = ctx.profile_read(); ctx.profile_enable(False)
print('round trip B=%d: %.4f ms' % (B, ms), {k: round(v[0] / 3, 4) for k, v in prof.items() if v[1]})
