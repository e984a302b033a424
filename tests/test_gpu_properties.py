"""Property-based GPU parity tests (hypothesis): small adversarial inputs - ties, touching and
degenerate boxes, duplicated rows, all-zero weight matrices - against the oracle."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st, HealthCheck

from oracle import ssd_codec_oracle as orc
from jpeg_detection_resnet_ssd_b200.bounding_box_utils import bounding_box_utils as bbu
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder import matching_utils as mu
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder import ssd_output_decoder as dec

pytestmark = pytest.mark.gpu
COMMON = dict(deadline=None, max_examples=40, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])

# coordinates on a coarse grid => many exact ties, touching edges and zero-area boxes
grid = st.integers(min_value=0, max_value=12).map(lambda v: v * 2.5)


@st.composite
def corner_boxes(draw, min_n=1, max_n=24, allow_degenerate=True):
    n = draw(st.integers(min_n, max_n))
    rows = []
    for _ in range(n):
        x0, y0 = draw(grid), draw(grid)
        w = draw(st.integers(0 if allow_degenerate else 1, 6)) * 2.5
        h = draw(st.integers(0 if allow_degenerate else 1, 6)) * 2.5
        rows.append([x0, y0, x0 + w, y0 + h])
    return np.array(rows, dtype=np.float64)


@settings(**COMMON)
@given(b1=corner_boxes(), b2=corner_boxes(), border=st.sampled_from(['half', 'include', 'exclude']))
def test_iou_outer_matches_oracle(ctx, b1, b2, border):
    with np.errstate(all='ignore'):
        want = orc.iou(b1, b2, coords='corners', mode='outer_product', border_pixels=border)
    got = bbu.iou(b1, b2, coords='corners', mode='outer_product', border_pixels=border)
    assert np.array_equal(got, want, equal_nan=True)


@settings(**COMMON)
@given(boxes=corner_boxes(max_n=40), data=st.data(), thr=st.sampled_from([0.0, 0.2, 0.45, 0.5, 1.0]),
       border=st.sampled_from(['half', 'include']))
def test_greedy_nms_matches_oracle(ctx, boxes, data, thr, border):
    n = boxes.shape[0]
    scores = np.array(data.draw(st.lists(st.integers(0, 5), min_size=n, max_size=n)), dtype=np.float64) / 5.0
    rows = np.concatenate([np.ones((n, 1)), scores[:, None], boxes], axis=1)
    with np.errstate(all='ignore'):
        want = orc.greedy_nms_rows(rows, 1, 2, thr, border)
    got = dec._greedy_nms2(rows, iou_threshold=thr, coords='corners', border_pixels=border)
    assert np.array_equal(np.asarray(got).reshape(-1, 6), np.asarray(want).reshape(-1, 6), equal_nan=True)


@settings(**COMMON)
@given(m=st.integers(1, 9), n=st.integers(9, 60), data=st.data())
def test_matching_matches_oracle(ctx, m, n, data):
    vals = data.draw(st.lists(st.integers(0, 4), min_size=m * n, max_size=m * n))
    w = (np.array(vals, dtype=np.float64) / 4.0).reshape(m, n)          # many ties and zero rows
    assert np.array_equal(mu.match_bipartite_greedy(w), orc.match_bipartite_greedy(w))
    for thr in (0.0, 0.5, 1.0):
        a, b = mu.match_multi(w, thr)
        c, d = orc.match_multi(w, thr)
        assert np.array_equal(a, c) and np.array_equal(b, d)


@settings(**COMMON)
@given(data=st.data(), thr=st.sampled_from([0.3, 0.45]), top_k=st.sampled_from([200, 'all']))
def test_decode_with_ties_matches_oracle(ctx, data, thr, top_k):
    """Tiny random layouts with confidences and boxes on coarse grids: every NMS pick is a tie broken by
    the anchor index; exp-free (w/h offsets 0) so the whole chain is IEEE-exact.  top_k=200 exercises the
    image-sweep path (no truncation), 'all' the per-class pipeline."""
    from jpeg_detection_resnet_ssd_b200 import _lib
    from helpers import product_rows7, to_rows7
    A = data.draw(st.integers(1, 48))
    C = 4
    B = 2
    y = np.zeros((B, A, C + 12), np.float32)
    conf = np.array(data.draw(st.lists(st.integers(0, 4), min_size=B * A * C, max_size=B * A * C)), np.float32).reshape(B, A, C) / 4
    y[:, :, :C] = conf
    cx = np.array(data.draw(st.lists(st.integers(1, 6), min_size=A, max_size=A)), np.float32) / 8
    cy = np.array(data.draw(st.lists(st.integers(1, 6), min_size=A, max_size=A)), np.float32) / 8
    wh = np.array(data.draw(st.lists(st.integers(1, 4), min_size=A, max_size=A)), np.float32) / 8
    y[:, :, C + 4] = cx; y[:, :, C + 5] = cy; y[:, :, C + 6] = wh; y[:, :, C + 7] = wh
    y[:, :, C + 8:] = np.float32([0.1, 0.1, 0.2, 0.2])
    kw = dict(confidence_thresh=0.2, iou_threshold=thr, top_k=top_k, input_coords='centroids',
              normalize_coords=True, img_height=64, img_width=64)
    want, wc = to_rows7(orc.decode_detections(y, exp_mode='cr', with_anchor_index=True, **kw))
    got, gc = product_rows7(*_lib.run_decode(y, _lib.MODE_PER_CLASS, 0.2, thr, top_k, 'centroids', True, 64, 64, 'half', ctx=ctx))
    assert np.array_equal(gc, wc) and np.array_equal(got, want)
