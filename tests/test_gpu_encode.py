"""GPU parity tests of SSDInputEncoder (matching + offset encoding) through the C ABI."""
import numpy as np
import pytest

from oracle import cases
from oracle import ssd_codec_oracle as orc
import synth
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder import ssd_input_encoder as enc_mod
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder import ssd_input_encoder_no_log as enc_nolog

from helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu

OFFSET_RTOL = 1e-5    # north_star: encoded offsets within 1e-5 relative


@pytest.fixture(autouse=True, params=['sparse', 'sparse_dense_patch', 'general', 'serial'])
def enc_path(request, ctx):
    """Every test runs through the encoder pipelines: shape-class sparse path (default; also with the patch
    kernel that scans the dense decision array instead of the position list), the general kernels beside
    the template stream, and the fully serial general kernels."""
    opts = {'sparse': {}, 'sparse_dense_patch': {'enc_dense_patch': 1}, 'general': {'enc_general': 1},
            'serial': {'enc_general': 1, 'enc_no_overlap': 1}}[request.param]
    for k, v in opts.items():
        ctx.set_option(k, v)
    yield request.param
    for k in opts:
        ctx.set_option(k, 0)


def make(case):
    kw = synth.layout_kwargs(case['layout'], **case.get('overrides', {}))
    mod = enc_mod if case.get('log_wh', True) else enc_nolog
    return mod.SSDInputEncoder(**kw), kw


@pytest.mark.parametrize('case', cases.ENCODE_CASES, ids=lambda c: c['name'])
def test_encode_matches_golden(case, ctx):
    g = load_golden(case['name'])
    gt = cases.build_encode_input(case)
    assert (cases.sha256_of(*gt) if len(gt) else '') == str(g['input_sha'])
    enc, kw = make(case)
    out = enc(gt, diagnostics=case.get('diagnostics', False), return_matches=True)
    y, mi = out[0], out[-1]
    B, A, W = y.shape
    assert [B, A, W] == list(g['shape'])
    flat = mi.reshape(-1)
    nz = np.nonzero(flat != -1)[0]
    # matched-anchor assignments and neutral set: bit-exact
    assert np.array_equal(nz, g['nz_idx'])
    assert np.array_equal(flat[nz], g['nz_match'])
    got = y.reshape(B * A, W)[nz]
    C = W - 12
    assert np.array_equal(got[:, :C], g['nz_rows'][:, :C])                    # class vectors
    assert np.array_equal(got[:, C + 4:], g['nz_rows'][:, C + 4:])            # anchors + variances
    assert rel_err(got[:, C:C + 4], g['nz_rows'][:, C:C + 4]).max(initial=0.0) <= OFFSET_RTOL
    # every other row is the plain background row of the template
    rest = np.ones(B * A, dtype=bool)
    rest[nz] = False
    bg = y.reshape(B * A, W)[rest]
    expect = np.zeros(C)
    expect[kw.get('background_id', 0)] = 1
    assert np.array_equal(bg[:, :C], np.broadcast_to(expect, bg[:, :C].shape))
    if case.get('log_wh', True):
        assert np.all(bg[:, C:C + 4] == 0)
    else:    # *_no_log twin: an unmatched row encodes (w_a / w_a) / variance
        v = np.asarray(kw['variances'], dtype=float)
        assert np.all(bg[:, C:C + 2] == 0) and np.array_equal(bg[:, C + 2:C + 4], np.broadcast_to(1.0 / v[2:], bg[:, C + 2:C + 4].shape))
    anchors = np.tile(synth.anchors_of(enc), (B, 1))[rest]
    assert np.array_equal(bg[:, C + 4:C + 8], anchors)
    assert np.array_equal(bg[:, C + 8:], np.broadcast_to(np.asarray(kw['variances'], dtype=float), bg[:, C + 8:].shape))
    if case.get('diagnostics'):
        y2 = out[1]
        assert np.all(y2[:, :, C:C + 4] == 0)
        assert np.array_equal(np.delete(y2, np.s_[C:C + 4], axis=2), np.delete(y, np.s_[C:C + 4], axis=2))


@pytest.mark.parametrize('seed', [81, 82, 83])
def test_encode_vs_live_oracle(seed, ctx):
    kw = synth.layout_kwargs('ssd300', neg_iou_limit=0.3)
    enc = enc_mod.SSDInputEncoder(**kw)
    oenc = orc.SSDInputEncoder(**kw)
    gt = synth.synth_ground_truth(300, 300, 20, 8, seed, max_boxes=25)
    y, mi = enc(gt, return_matches=True)
    yo, mo = oenc(gt, return_matches=True)
    assert np.array_equal(mi, mo)
    assert rel_err(y, yo).max() <= OFFSET_RTOL
    C = 21
    assert np.array_equal(y[:, :, :C], yo[:, :, :C]) and np.array_equal(y[:, :, C + 4:], yo[:, :, C + 4:])


def test_template_and_errors(ctx):
    kw = synth.layout_kwargs('tiny')
    enc = enc_mod.SSDInputEncoder(**kw)
    oenc = orc.SSDInputEncoder(**kw)
    assert np.array_equal(enc.generate_encoding_template(3), oenc.generate_encoding_template(3))
    t, centers, wh, steps, offs = enc.generate_encoding_template(1, diagnostics=True)
    assert len(centers) == 3 and len(wh) == 3
    bad = [np.array([[1, 10, 10, 10, 30]], dtype=float)]
    with pytest.raises(enc_mod.DegenerateBoxError):
        enc([np.array([[1, 5, 5, 50, 50]], dtype=float)] + bad)
    with pytest.raises(IndexError):
        enc([np.array([[9, 5, 5, 50, 50]], dtype=float)])
    # batch of empty images: pure background
    y = enc([np.zeros((0, 5)), np.zeros((0, 5))])
    assert np.all(y[:, :, 0] == 1) and np.all(y[:, :, 1:8] == 0)


def test_encode_many_boxes_and_large_batch(ctx):
    """more ground-truth boxes than any VOC image has, and a batch large enough to need several
    waves: compared with the oracle on a sample of images, checked for batch independence."""
    kw = synth.layout_kwargs('ssd300')
    enc = enc_mod.SSDInputEncoder(**kw)
    oenc = orc.SSDInputEncoder(**kw)
    gt = synth.synth_ground_truth(300, 300, 20, 64, 91, max_boxes=60, min_boxes=30)
    gt = gt + gt
    y, mi = enc(gt, return_matches=True)
    assert np.array_equal(y[:64], y[64:]) and np.array_equal(mi[:64], mi[64:])
    yo, mo = oenc(gt[:4], return_matches=True)
    assert np.array_equal(mi[:4], mo)
    assert rel_err(y[:4], yo).max() <= OFFSET_RTOL


def test_encode_runner_up_in_other_class(ctx):
    """Two identical ground-truth boxes: the second one loses its best anchor in the bipartite rounds and
    must fall back to the true runner-up - which lies in another shape class / anchor chunk than the best
    (50 px box: best is a 42 px anchor of the first predictor layer, runner-up the 60 px anchor of the
    second layer that contains it).  Both pipelines prune pairs against the row's best IoU, so this checks
    that nothing pruned is needed later."""
    kw = synth.layout_kwargs('ssd300')
    enc = enc_mod.SSDInputEncoder(**kw)
    oenc = orc.SSDInputEncoder(**kw)
    gt = []
    rng = np.random.default_rng(5)
    for i in range(24):
        cx, cy = 4 + 8 * int(rng.integers(4, 30)), 4 + 8 * int(rng.integers(4, 30))
        half = 25 + int(rng.integers(-3, 4))
        box = [float(rng.integers(1, 21)), cx - half, cy - half, cx + half, cy + half]
        rows = [box, box]
        if i % 3 == 0:
            rows.append([3.0, cx - half, cy - half, cx + half, cy + half])          # a third claimant, other class
        if i % 4 == 1:
            rows.insert(0, [7.0, 20, 30, 260, 280])
        gt.append(np.array(rows, dtype=float))
    y, mi = enc(gt, return_matches=True)
    yo, mo = oenc(gt, return_matches=True)
    assert np.array_equal(mi, mo)
    assert np.array_equal(y[:, :, :21], yo[:, :, :21])
    assert rel_err(y, yo).max() <= OFFSET_RTOL


def test_encode_more_rows_than_sparse_path_takes(ctx):
    """> 128 ground-truth rows in an image (beyond the sparse path's shared-memory tables)."""
    kw = synth.layout_kwargs('tiny')
    enc = enc_mod.SSDInputEncoder(**kw)
    oenc = orc.SSDInputEncoder(**kw)
    gt = synth.synth_ground_truth(kw['img_height'], kw['img_width'], kw['n_classes'], 3, 17, max_boxes=170, min_boxes=140)
    y, mi = enc(gt, return_matches=True)
    yo, mo = oenc(gt, return_matches=True)
    assert np.array_equal(mi, mo)
    assert rel_err(y, yo).max() <= OFFSET_RTOL


def test_encode_bench_batch_vs_oracle(ctx, enc_path):
    """The bench's round-trip batch (512 images, seed 78), every image against the oracle.  (One of these images
    exposes a runner-up that lies in an anchor chunk / shape class which the first pass pruned against the
    row's original best - the case test_encode_runner_up_in_other_class builds by hand.)"""
    if enc_path == 'serial':
        pytest.skip('same kernels as "general"')
    kw = synth.layout_kwargs('ssd300')
    enc = enc_mod.SSDInputEncoder(**kw)
    oenc = orc.SSDInputEncoder(**kw)
    gt = synth.synth_ground_truth(300, 300, 20, 512, seed=78)
    y, mi = enc(gt, return_matches=True)
    for i in range(len(gt)):
        yo, mo = oenc([gt[i]], return_matches=True)
        assert np.array_equal(mi[i], mo[0]), i
        nz = np.nonzero(mo[0] != -1)[0]
        assert np.array_equal(y[i][nz, :21], yo[0][nz, :21]), i
        assert rel_err(y[i][nz], yo[0][nz]).max(initial=0.0) <= OFFSET_RTOL, i


@pytest.mark.parametrize('layout', ['tiny', 'ssd300'])
def test_encode_low_thresholds_everything_matches(ctx, layout):
    """IoU thresholds close to zero: almost every anchor is a positive match, the per-row candidate lists of the
    sparse path overflow (rows fall back to the cooperative rescan) and the patch list approaches its capacity."""
    kw = synth.layout_kwargs(layout, pos_iou_threshold=0.02, neg_iou_limit=0.01)
    enc = enc_mod.SSDInputEncoder(**kw)
    oenc = orc.SSDInputEncoder(**kw)
    H, W = kw['img_height'], kw['img_width']
    gt = [np.array([[1, 0.05 * W, 0.05 * H, 0.95 * W, 0.95 * H], [2, 0.1 * W, 0.2 * H, 0.6 * W, 0.9 * H]], dtype=float),
          np.array([[3, 0.3 * W, 0.1 * H, 0.9 * W, 0.5 * H]], dtype=float),
          synth.synth_ground_truth(H, W, kw['n_classes'], 1, 5, max_boxes=9, min_boxes=9)[0]]
    y, mi = enc(gt, return_matches=True)
    yo, mo = oenc(gt, return_matches=True)
    assert np.array_equal(mi, mo)
    assert (mi >= 0).mean() > 0.5
    C = kw['n_classes'] + 1
    assert np.array_equal(y[:, :, :C], yo[:, :, :C]) and np.array_equal(y[:, :, C + 4:], yo[:, :, C + 4:])
    assert rel_err(y, yo).max() <= OFFSET_RTOL


def test_back_to_back_device_encodes_keep_their_own_ground_truth(ctx):
    """`ssdc_encode` with device outputs only enqueues work: two calls with different ground truth and no
    synchronisation in between must each be matched against their OWN rows and offsets (the pinned staging of the
    per-call host data is an event-guarded ring), even when the caller reuses its host arrays right away."""
    from jpeg_detection_resnet_ssd_b200 import _lib
    enc = synth.make_encoder(enc_mod.SSDInputEncoder, 'ssd300')
    _, h = enc._encoder()
    lib = ctx.lib
    B, n = 8, 8 * 8732 * 33 * 8
    batches = [synth.synth_ground_truth(300, 300, 20, B, seed=s, min_boxes=1 + 3 * s, max_boxes=3 + 3 * s) for s in range(6)]
    want = [enc(gt) for gt in batches]
    bufs = [ctx.dev_alloc(n) for _ in batches]
    flat0, offs0 = synth.flatten_ground_truth(batches[-1])
    flat = np.zeros((max(synth.flatten_ground_truth(g)[0].shape[0] for g in batches), 5))
    offs = np.zeros(B + 1, np.int64)
    for gt, d in zip(batches, bufs):
        f, o = synth.flatten_ground_truth(gt)
        flat[:f.shape[0]] = f          # the SAME host arrays are overwritten for every call
        offs[:] = o
        _lib.check(lib.ssdc_encode(h, _lib.ptr(flat), _lib.ptr(offs), B, 1, d, None, None))
    flat[:] = 0.0
    ctx.synchronize()
    for w, d in zip(want, bufs):
        got = np.empty_like(w)
        ctx.d2h(got, d)
        assert np.array_equal(got, w)
        ctx.dev_free(d)


@pytest.mark.parametrize('lanes', [1, 2, 3, 6])
def test_device_encodes_on_lanes_are_ordered_where_they_must_be(ctx, lanes):
    """Device-output encodes run on `enc_lanes` lanes (stream pair + scratch each).  Calls that write the same buffer must
    land in call order (the later call wins, never a mix of one call's template and another's patches), a reader enqueued
    behind them (ssdc_memcpy_d2h here) must see the finished result, and a different `match_idx` output per call must
    stay with its call."""
    from jpeg_detection_resnet_ssd_b200 import _lib
    enc = synth.make_encoder(enc_mod.SSDInputEncoder, 'ssd300')
    _, h = enc._encoder()
    lib = ctx.lib
    B, A = 6, 8732
    n = B * A * 33 * 8
    batches = [synth.synth_ground_truth(300, 300, 20, B, seed=40 + s, min_boxes=1 + 2 * s, max_boxes=4 + 2 * s) for s in range(7)]
    want = [enc(gt, return_matches=True) for gt in batches]
    buf = [ctx.dev_alloc(n) for _ in range(2)]
    idx = [ctx.dev_alloc(B * A * 4) for _ in batches]
    target = [0, 0, 1, 0, 1, 1, 0]
    ctx.set_option('enc_lanes', lanes)
    try:
        for _ in range(3):                                    # (repeated: the lanes' scratch is reused while calls are in flight)
            for gt, t, di in zip(batches, target, idx):
                f, o = synth.flatten_ground_truth(gt)
                _lib.check(lib.ssdc_encode(h, _lib.ptr(f), _lib.ptr(o), B, 1, buf[t], None, di))
        got = np.empty((B, A, 33))
        for t, last in ((0, 6), (1, 5)):
            ctx.d2h(got, buf[t])                              # no synchronize in between: the copy itself runs behind the lanes
            assert np.array_equal(got, want[last][0]), (lanes, t)
        mi = np.empty((B, A), np.int32)
        for w, di in zip(want, idx):
            ctx.d2h(mi, di)
            assert np.array_equal(mi, w[1])
    finally:
        ctx.set_option('enc_lanes', 0)
        ctx.synchronize()
        for d in buf + idx:
            ctx.dev_free(d)
