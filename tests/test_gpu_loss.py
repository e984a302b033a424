"""GPU parity of the SSD multibox loss (`keras_loss_function.keras_ssd_loss.SSDLoss`, `ssdc_ssd_loss`) with the
numpy oracle.  Floating point: 1e-5 relative (float32 arithmetic per box, sums in another order)."""
import numpy as np
import pytest

from oracle import ssd_loss_oracle as lo
from oracle import ssd_codec_oracle as orc
import synth
from jpeg_detection_resnet_ssd_b200.keras_loss_function.keras_ssd_loss import SSDLoss

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.fixture(autouse=True, params=['tma', 'ldg'])
def loss_path(request, ctx):
    """Both loaders of the box kernel: TMA ring (default) and the plain tile copy (unaligned inputs)."""
    ctx.set_option('loss_no_tma', 1 if request.param == 'ldg' else 0)
    yield request.param
    ctx.set_option('loss_no_tma', 0)


def make_batch(layout, B, seed, bg_bias=3.0):
    kw = synth.layout_kwargs(layout)
    enc = orc.SSDInputEncoder(**kw)
    gt = synth.synth_ground_truth(kw['img_height'], kw['img_width'], kw['n_classes'], B, seed, max_boxes=8)
    y_true = enc(gt)
    anchors = y_true[0, :, -8:-4]
    y_pred = synth.synth_y_pred(anchors, kw['variances'], kw['n_classes'] + 1, B, seed + 1, bg_bias=bg_bias, hot=20)
    return y_true, y_pred


def close(a, b):
    return np.allclose(a, b, rtol=RTOL, atol=1e-6)


@pytest.mark.parametrize('layout,B', [('tiny', 5), ('ssd300', 6)])
@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_loss_matches_oracle(ctx, layout, B, dtype):
    y_true, y_pred = make_batch(layout, B, 11)
    got = SSDLoss().compute_loss(y_true.astype(dtype), y_pred)
    want = lo.compute_loss(y_true, y_pred)
    assert got.dtype == np.float32 and got.shape == (B,)
    assert close(got, want), (got, want)


@pytest.mark.parametrize('kw', [dict(neg_pos_ratio=1), dict(neg_pos_ratio=3, n_neg_min=500), dict(alpha=0.25), dict(neg_pos_ratio=0)])
def test_loss_parameters(ctx, kw):
    y_true, y_pred = make_batch('tiny', 7, 23)
    assert close(SSDLoss(**kw).compute_loss(y_true, y_pred), lo.compute_loss(y_true, y_pred, **kw))


def test_loss_edge_cases(ctx):
    y_true, y_pred = make_batch('tiny', 4, 31)
    # no positives at all: every box background
    yt = y_true.copy()
    yt[:, :, :-12] = 0.0
    yt[:, :, 0] = 1.0
    assert close(SSDLoss().compute_loss(yt, y_pred), lo.compute_loss(yt, y_pred))
    assert close(SSDLoss(n_neg_min=10).compute_loss(yt, y_pred), lo.compute_loss(yt, y_pred, n_neg_min=10))
    # neutral boxes (all-zero class vectors) are ignored by both parts
    yt = y_true.copy()
    yt[:, ::3, :-12] = 0.0
    assert close(SSDLoss().compute_loss(yt, y_pred), lo.compute_loss(yt, y_pred))
    # all negative losses zero: background predicted with confidence exactly 1
    yp = y_pred.copy()
    yp[:, :, :-12] = 0.0
    yp[:, :, 0] = 1.0
    assert close(SSDLoss().compute_loss(y_true, yp), lo.compute_loss(y_true, yp))
    # zeros in y_pred are clamped to 1e-15
    yp = y_pred.copy()
    yp[:, ::5, 1] = 0.0
    assert close(SSDLoss().compute_loss(y_true, yp), lo.compute_loss(y_true, yp))
    with pytest.raises(ValueError):
        SSDLoss().compute_loss(y_true, y_pred[:, :-1])


def test_loss_ties_take_lower_indices(ctx):
    """Many boxes share the k-th largest negative loss (identical predictions): tf.nn.top_k admits them in flat index
    order, which decides WHICH image's sum they enter."""
    y_true, y_pred = make_batch('tiny', 6, 41)
    yp = y_pred.copy()
    yp[:, :, :-12] = yp[0, 0, :-12]                      # every box predicts the same class distribution
    for ratio in (1, 2, 3):
        got = SSDLoss(neg_pos_ratio=ratio).compute_loss(y_true, yp)
        want = lo.compute_loss(y_true, yp, neg_pos_ratio=ratio)
        assert close(got, want), (ratio, got, want)
    # a few distinct values, each shared by many boxes
    rng = np.random.default_rng(3)
    pick = rng.integers(0, 4, size=yp.shape[:2])
    for v in range(4):
        yp[pick == v, :-12] = y_pred[v, v, :-12]
    assert close(SSDLoss().compute_loss(y_true, yp), lo.compute_loss(y_true, yp))
