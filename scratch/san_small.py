"""Small encode + decode + VOC run for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from jpeg_detection_resnet_ssd_b200 import synth
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_output_decoder import decode_detections, decode_detections_fast
for layout, B in (('tiny', 5), ('ssd300', 6)):
    kw = synth.layout_kwargs(layout)
    enc = SSDInputEncoder(**kw)
    gt = synth.synth_ground_truth(kw['img_height'], kw['img_width'], kw['n_classes'], B, seed=3, max_boxes=12)
    gt[1] = np.concatenate([gt[1], gt[1][:1]])          # duplicate row: bipartite collision
    y, mi = enc(gt, return_matches=True)
    y2 = enc(gt, diagnostics=True)
    out = decode_detections_fast(y, confidence_thresh=0.5, iou_threshold=0.45, top_k='all', img_height=kw['img_height'], img_width=kw['img_width'])
    yp = synth.synth_y_pred(synth.anchors_of(enc), enc.variances, kw['n_classes'] + 1, B, 11, bg_bias=6.0, hot=20)
    det = decode_detections(yp, 0.01, 0.45, 50, img_height=kw['img_height'], img_width=kw['img_width'])
    print(layout, y.shape, sum(len(o) for o in out), sum(len(o) for o in det))
print('done')
