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


def _layer_input_from_tf_boxes(boxes_yxyx, scores, fast):
    """(1, n, 14) float32 y_pred whose layer decode reproduces the given boxes: zero offsets, anchor = the box in
    centroid form (negative sizes for flipped corners), class 1 confidence = the score."""
    b = np.asarray(boxes_yxyx, np.float32).reshape(-1, 4)
    n = b.shape[0]
    y = np.zeros((1, n, 14), np.float32)
    y[0, :, 1] = scores
    y[0, :, 0] = np.asarray(scores, np.float32) - np.float32(1.0) if fast else 0.0
    y[0, :, 6] = (b[:, 1] + b[:, 3]) / np.float32(2)       # cx
    y[0, :, 7] = (b[:, 0] + b[:, 2]) / np.float32(2)       # cy
    y[0, :, 8] = b[:, 3] - b[:, 1]                         # w (x2 - x1)
    y[0, :, 9] = b[:, 2] - b[:, 0]                         # h
    y[0, :, 10:] = [0.1, 0.1, 0.2, 0.2]
    return y


@pytest.mark.parametrize('mode', ['layer', 'layer_fast'])
def test_device_layer_modes_reproduce_tensorflows_published_nms_answers(mode, ctx):
    """The device layer modes on TensorFlow's own NMS unit-test cases (tests/golden/tf_nms_published.json): the
    selected boxes, in selection order, are the published answers.  max_output_size below top_k takes the general
    pipeline, at or above it the image sweep."""
    import json
    import os
    from jpeg_detection_resnet_ssd_b200 import _lib
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'tf_nms_published.json')) as fh:
        vec = json.load(fh)
    m = _lib.MODE_LAYER if mode == 'layer' else _lib.MODE_LAYER_FAST
    for c in vec['cases']:
        if not c['boxes']:
            # (A = 0 is not a valid tensor shape at the C ABI: one box below the confidence threshold instead)
            y = _layer_input_from_tf_boxes([[0, 0, 1, 1]], [-5.0], mode == 'layer_fast')
        else:
            y = _layer_input_from_tf_boxes(c['boxes'], c['scores'], mode == 'layer_fast')
        for top_k in (4, 40):
            rows, counts, idx = _lib.run_decode(y, m, -1.5, c['iou_threshold'], top_k, 'centroids', False, None, None, 'half',
                                                log_wh=True, nms_cap=c['max_output_size'], ctx=ctx)
            assert idx.tolist() == c['selected'][:top_k], (c['name'], top_k)
            assert counts.tolist() == [len(c['selected'][:top_k])]
            if c['selected']:
                assert np.array_equal(rows[:, 1].astype(np.float32), np.asarray(c['scores'], np.float32)[c['selected'][:top_k]])


@pytest.mark.parametrize('layer_cls,oracle_fn', [(DecodeDetections, orc.decode_layer), (DecodeDetectionsFast, orc.decode_layer_fast)])
@pytest.mark.parametrize('cap,top_k', [(400, 200), (8, 200), (3, 50), (400, 20)])
def test_keras_layer_contract_ssd300(layer_cls, oracle_fn, cap, top_k, ctx):
    """SSD300-sized layer runs against the CPU restatement: the reference's defaults (cap 400, top_k 200), a
    per-class cap that binds (8 and 3 boxes per class: `nms_max_output_size < top_k` also switches the kernel path),
    and a small top_k."""
    enc = synth.make_encoder(SSDInputEncoder, 'ssd300')
    y = synth.synth_y_pred(synth.anchors_of(enc), enc.variances, 21, 2, 96, bg_bias=7.0, hot=60)
    layer = layer_cls(confidence_thresh=0.01, iou_threshold=0.45, top_k=top_k, nms_max_output_size=cap, img_height=300, img_width=300)
    out = layer(y)
    want = oracle_fn(y, 0.01, 0.45, top_k, cap, True, 300, 300, exp_mode='cr')
    assert out.shape == (2, top_k, 6) and out.dtype == np.float32
    assert np.array_equal(out, want)
    if cap < top_k and layer_cls is DecodeDetections:
        per_class = np.bincount(out[0, :, 0].astype(int), minlength=21)[1:]
        assert per_class.max() == cap         # the cap really binds
