"""GPU tests of the one-process / many-devices mode: the batch is split into
contiguous shards across the context's devices, results are gathered in image order and must equal
the single-device results bit for bit.  Skipped on a single-GPU box."""
import numpy as np
import pytest

import synth
from jpeg_detection_resnet_ssd_b200 import _lib, Context, set_context
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder

pytestmark = pytest.mark.gpu


def test_multi_device_context_matches_single(ctx):
    n = _lib.load_library().ssdc_device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    enc = synth.make_encoder(SSDInputEncoder, 'ssd300')
    y = synth.synth_y_pred(synth.anchors_of(enc), enc.variances, 21, 11, 5, bg_bias=7.5, hot=40)   # 11: ragged shards
    ref = _lib.run_decode(y, _lib.MODE_PER_CLASS, 0.01, 0.45, 200, 'centroids', True, 300, 300, 'half', ctx=ctx)
    gt = synth.synth_ground_truth(300, 300, 20, 11, 6)
    y_ref, m_ref = enc(gt, return_matches=True)
    multi = Context(list(range(min(n, 8))))
    old = set_context(multi)
    try:
        got = _lib.run_decode(y, _lib.MODE_PER_CLASS, 0.01, 0.45, 200, 'centroids', True, 300, 300, 'half', ctx=multi)
        for a, b in zip(ref, got):
            assert np.array_equal(a, b)
        enc2 = synth.make_encoder(SSDInputEncoder, 'ssd300')
        y_got, m_got = enc2(gt, return_matches=True)
        assert np.array_equal(m_ref, m_got) and np.array_equal(y_ref, y_got, equal_nan=True)
        # fewer images than devices
        got1 = _lib.run_decode(y[:1], _lib.MODE_FAST, 0.2, 0.45, 'all', 'centroids', True, 300, 300, 'half', ctx=multi)
        ref1 = _lib.run_decode(y[:1], _lib.MODE_FAST, 0.2, 0.45, 'all', 'centroids', True, 300, 300, 'half', ctx=ctx)
        for a, b in zip(ref1, got1):
            assert np.array_equal(a, b)
    finally:
        set_context(old)
        multi.close()
