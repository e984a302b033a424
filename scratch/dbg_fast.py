import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from oracle import cases
from jpeg_detection_resnet_ssd_b200 import _lib, synth
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
from helpers import load_golden, product_rows7
case = [c for c in cases.DECODE_CASES if c['name']=='f_ssd300_fast'][0]
g = load_golden(case['name'])
y = cases.build_decode_input(case, SSDInputEncoder)
kw = case['kwargs']
rows, counts, idx = _lib.run_decode(y, _lib.MODE_FAST, kw['confidence_thresh'], kw['iou_threshold'], kw['top_k'], 'centroids', True, 300, 300, 'half')
got, gc = product_rows7(rows, counts, idx)
want, wc = g['rows_cr'], g['counts_cr']
print('counts', gc, wc)
n = min(len(got), len(want))
d = np.nonzero((got[:n,:3] != want[:n,:3]).any(axis=1))[0]
print('first diffs', d[:10], len(d))
if len(d):
    i = d[0]
    print(got[i-1:i+3]); print(want[i-1:i+3])
# set differences image 0
a = set(map(int, got[:gc[0],0])); b = set(map(int, want[:wc[0],0]))
print('only got', sorted(a-b)[:20], 'only want', sorted(b-a)[:20])
