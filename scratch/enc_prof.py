"""Encode B=1024 (device-resident output) a few times; prints the device-timed ms per call."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from jpeg_detection_resnet_ssd_b200 import _lib, synth
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
enc = synth.make_encoder(SSDInputEncoder, 'ssd300')
ctx, h = enc._encoder()
lib = ctx.lib
gt = synth.synth_ground_truth(300, 300, 20, B, seed=77)
flat, offs = synth.flatten_ground_truth(gt)
d_out = ctx.dev_alloc(B * 8732 * 33 * 8)
for _ in range(3):
    _lib.check(lib.ssdc_encode(h, _lib.ptr(flat), _lib.ptr(offs), B, 1, d_out, None, None))
ctx.synchronize()
ctx.timer_start()
for _ in range(steps):
    _lib.check(lib.ssdc_encode(h, _lib.ptr(flat), _lib.ptr(offs), B, 1, d_out, None, None))
ms = ctx.timer_stop()
print('encode B=%d: %.4f ms/call' % (B, ms / steps))
