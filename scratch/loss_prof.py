import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, numpy as np
from jpeg_detection_resnet_ssd_b200 import _lib, synth
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
enc = synth.make_encoder(SSDInputEncoder, 'ssd300')
ctx = _lib.get_context()
print(bench.loss_numbers(ctx, _lib, enc, 6544.7))
