"""GPU parity tests of the standalone ops (iou, convert_coordinates, matching, greedy_nms) and of
the Keras-layer contract (parity unpinned: checked against the CPU restatement only)."""
import numpy as np
import pytest

from oracle import ssd_codec_oracle as orc
import synth
from jpeg_detection_resnet_ssd_b200.bounding_box_utils import bounding_box_utils as bbu
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder import matching_utils as mu
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder import ssd_output_decoder as dec
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
from jpeg_detection_resnet_ssd_b200.keras_layers.keras_layer_DecodeDetections import DecodeDetections
from jpeg_detection_resnet_ssd_b200.keras_layers.keras_layer_DecodeDetectionsFast import DecodeDetectionsFast

from helpers import load_golden

pytestmark = pytest.mark.gpu


def test_iou_and_intersection(ctx):
    g = load_golden('thin_ops')
    for fmt in ('corners', 'minmax', 'centroids'):
        for border in ('half', 'include', 'exclude'):
            o = bbu.iou(g['b1_' + fmt], g['b2_' + fmt], coords=fmt, mode='outer_product', border_pixels=border)
            assert np.array_equal(o, g['iou_outer_%s_%s' % (fmt, border)])
            o = bbu.iou(g['b1_' + fmt], g['b3_' + fmt], coords=fmt, mode='element-wise', border_pixels=border)
            assert np.array_equal(o, g['iou_elem_%s_%s' % (fmt, border)])
            o = bbu.iou(g['b1_' + fmt], g['b2_' + fmt][0], coords=fmt, mode='element-wise', border_pixels=border)
            assert np.array_equal(o, g['iou_bcast_%s_%s' % (fmt, border)])
            if fmt != 'centroids':
                o = bbu.intersection_area(g['b1_' + fmt], g['b2_' + fmt], coords=fmt, mode='outer_product', border_pixels=border)
                assert np.array_equal(o, g['inter_outer_%s_%s' % (fmt, border)])
    with pytest.raises(ValueError):
        bbu.iou(np.zeros((2, 3)), np.zeros((2, 4)))
    with pytest.raises(ValueError):
        bbu.iou(np.zeros((2, 4)), np.zeros((2, 4)), mode='inner')


def test_convert_coordinates(ctx):
    g = load_golden('thin_ops')
    for conv in ('minmax2centroids', 'centroids2minmax', 'corners2centroids', 'centroids2corners',
                 'minmax2corners', 'corners2minmax'):
        for border in ('half', 'include', 'exclude'):
            for key, start in (('conv32', 3), ('conv64', -5)):
                o = bbu.convert_coordinates(g[key], start_index=start, conversion=conv, border_pixels=border)
                w = g['conv_%s_%s_%s' % (key, conv, border)]
                assert o.dtype == np.float64 and np.array_equal(o, w)
    with pytest.raises(ValueError):
        bbu.convert_coordinates(g['conv64'], 0, 'corners2polar')


def test_matching(ctx):
    g = load_golden('thin_ops')
    for key in ('w_small', 'w_quirk'):
        assert np.array_equal(mu.match_bipartite_greedy(g[key]), g['bip_' + key])
        gt, an = mu.match_multi(g[key], 0.5)
        assert np.array_equal(gt, g['multi_gt_' + key]) and np.array_equal(an, g['multi_anchor_' + key])
    rng = np.random.default_rng(5)
    w = rng.uniform(0, 1, (37, 2000))
    w[rng.uniform(size=w.shape) < 0.7] = 0
    assert np.array_equal(mu.match_bipartite_greedy(w), orc.match_bipartite_greedy(w))
    a, b = mu.match_multi(w, 0.8)
    c, d = orc.match_multi(w, 0.8)
    assert np.array_equal(a, c) and np.array_equal(b, d)


def test_greedy_nms(ctx):
    g = load_golden('thin_ops')
    rows = g['nms_rows']
    o = dec.greedy_nms([rows, rows[:40]], iou_threshold=0.3, coords='corners', border_pixels='half')
    assert np.array_equal(o[0], g['nms_full']) and np.array_equal(o[1], g['nms_40'])
    assert np.array_equal(dec._greedy_nms(rows[:, 1:], iou_threshold=0.45, coords='corners', border_pixels='include'), g['nms1'])
    assert np.array_equal(dec._greedy_nms2(rows, iou_threshold=0.45, coords='corners', border_pixels='half'), g['nms2'])
    rng = np.random.default_rng(9)
    n = 5000
    x0 = rng.uniform(0, 280, n); y0 = rng.uniform(0, 280, n)
    big = np.stack([np.ones(n), np.round(rng.uniform(0, 1, n), 3), x0, y0, x0 + rng.uniform(5, 60, n), y0 + rng.uniform(5, 60, n)], 1)
    assert np.array_equal(dec._greedy_nms2(big, 0.45), orc.greedy_nms_rows(big, 1, 2, 0.45))


@pytest.mark.parametrize('layer_cls,oracle_fn', [(DecodeDetections, orc.decode_layer), (DecodeDetectionsFast, orc.decode_layer_fast)])
def test_keras_layer_contract(layer_cls, oracle_fn, ctx):
    """PARITY UNPINNED (no TensorFlow here): compared with the CPU restatement of the layer."""
    enc = synth.make_encoder(SSDInputEncoder, 'tiny')
    y = synth.synth_y_pred(synth.anchors_of(enc), enc.variances, 4, 3, 95, bg_bias=1.0, hot=10)
    layer = layer_cls(confidence_thresh=0.05, iou_threshold=0.45, top_k=30, nms_max_output_size=12,
                      img_height=96, img_width=128)
    out = layer(y)
    want = oracle_fn(y, 0.05, 0.45, 30, 12, True, 96, 128, exp_mode='cr')
    assert out.shape == (3, 30, 6) and out.dtype == np.float32
    assert np.array_equal(out, want)
    assert layer.compute_output_shape((3, 100, 16)) == (3, 30, 6)
    assert layer.get_config()['top_k'] == 30
    with pytest.raises(ValueError):
        layer_cls(coords='corners', img_height=1, img_width=1)
