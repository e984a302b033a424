"""Compare encoder pipelines (sparse / general) on the round-trip bench input, oracle on the images that differ."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from jpeg_detection_resnet_ssd_b200 import synth
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
from oracle import ssd_codec_oracle as orc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 78
kw = synth.layout_kwargs('ssd300')
enc = SSDInputEncoder(**kw)
gt = synth.synth_ground_truth(300, 300, 20, B, seed=seed)
os.environ.pop('SSDC_ENC_GENERAL', None)
y1, m1 = enc(gt, return_matches=True)
os.environ['SSDC_ENC_GENERAL'] = '1'
y2, m2 = enc(gt, return_matches=True)
os.environ['SSDC_ENC_NO_OVERLAP'] = '1'
y3, m3 = enc(gt, return_matches=True)
d12 = np.nonzero(np.any(m1 != m2, axis=1))[0]
d23 = np.nonzero(np.any(m2 != m3, axis=1))[0]
print('sparse vs general: images differing', d12.tolist(), '; general vs serial', d23.tolist(), '; y equal', np.array_equal(y1, y2), np.array_equal(y2, y3))
oenc = orc.SSDInputEncoder(**kw)
check = sorted(set(d12.tolist()) | set(d23.tolist()) | set(range(0, B, max(1, B // 16))))
bad = 0
for i in check:
    yo, mo = oenc([gt[i]], return_matches=True)
    for name, m in (('sparse', m1), ('general', m2), ('serial', m3)):
        if not np.array_equal(m[i], mo[0]):
            bad += 1
            w = np.nonzero(m[i] != mo[0])[0]
            print('image', i, name, 'differs from oracle at anchors', w[:8], 'got', m[i][w[:8]], 'want', mo[0][w[:8]])
print('checked', len(check), 'images against the oracle, mismatching (image, pipeline) pairs:', bad)
if len(sys.argv) > 3:
    bad = 0
    for i in range(B):
        yo, mo = oenc([gt[i]], return_matches=True)
        if not np.array_equal(m1[i], mo[0]):
            bad += 1; print('image', i, 'differs')
    print('full oracle comparison over', B, 'images: mismatches', bad)
