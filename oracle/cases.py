"""Parity cases shared by oracle/make_golden.py and the tests (test infrastructure).

Each case is a dict with a builder for the seeded input and the keyword arguments of the
codec call.  Inputs are regenerated from seeds by `jpeg_detection_resnet_ssd_b200.synth`
(pure numpy) on every host; the golden files store a SHA-256 of the input bytes so that a
host whose RNG / libm produce different inputs is detected instead of silently compared.
"""
from __future__ import division

import hashlib

import numpy as np

import synth


def sha256_of(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


# ----------------------------------------------------------------------------------------
# decode cases
# ----------------------------------------------------------------------------------------
def _ypred(encoder_cls, layout, batch, seed, coords='centroids', dtype=np.float32, quantize=None,
           kill=None, **gen):
    enc = synth.make_encoder(encoder_cls, layout, coords=coords)
    anchors = synth.anchors_of(enc)
    y = synth.synth_y_pred(anchors, enc.variances, enc.n_classes, batch, seed, dtype=dtype, **gen)
    C = enc.n_classes
    if quantize:
        # many exactly equal scores: every NMS pick is a tie broken by the anchor index
        y[:, :, :C] = (np.round(y[:, :, :C] * quantize) / quantize).astype(y.dtype)
    if kill is not None:
        # image `kill` gets no candidate at all
        y[kill, :, 1:C] = 0
        y[kill, :, 0] = 1
    return y


DECODE_CASES = [
    dict(name='d_ssd300_sparse', fn='decode_detections', layout='ssd300', batch=2, seed=101,
         gen=dict(bg_bias=9.0, hot=40),
         kwargs=dict(confidence_thresh=0.01, iou_threshold=0.45, top_k=200, input_coords='centroids',
                     normalize_coords=True, img_height=300, img_width=300)),
    dict(name='d_ssd300_dense', fn='decode_detections', layout='ssd300', batch=1, seed=102,
         gen=dict(bg_bias=6.0, hot=40),
         kwargs=dict(confidence_thresh=0.01, iou_threshold=0.45, top_k=200, input_coords='centroids',
                     normalize_coords=True, img_height=300, img_width=300)),
    dict(name='d_ssd300_expfree', fn='decode_detections', layout='ssd300', batch=2, seed=103,
         gen=dict(bg_bias=7.5, hot=60, exp_free=True),
         kwargs=dict(confidence_thresh=0.01, iou_threshold=0.45, top_k=200, input_coords='centroids',
                     normalize_coords=True, img_height=300, img_width=300)),
    dict(name='d_ssd300_ties', fn='decode_detections', layout='ssd300', batch=1, seed=104,
         gen=dict(bg_bias=5.0, hot=30, exp_free=True), quantize=64.0,
         kwargs=dict(confidence_thresh=0.01, iou_threshold=0.45, top_k='all', input_coords='centroids',
                     normalize_coords=True, img_height=300, img_width=300)),
    dict(name='d_ssd300_include', fn='decode_detections', layout='ssd300', batch=1, seed=105,
         gen=dict(bg_bias=8.0, hot=80),
         kwargs=dict(confidence_thresh=0.01, iou_threshold=0.45, top_k=200, input_coords='centroids',
                     normalize_coords=True, img_height=300, img_width=300, border_pixels='include')),
    dict(name='d_tiny_minmax_all', fn='decode_detections', layout='tiny', batch=3, seed=106, coords='minmax',
         gen=dict(bg_bias=2.0, hot=10, offset_sigma=0.8),
         kwargs=dict(confidence_thresh=0.05, iou_threshold=0.3, top_k='all', input_coords='minmax',
                     normalize_coords=True, img_height=96, img_width=128, border_pixels='half')),
    dict(name='d_tiny_corners_f32', fn='decode_detections', layout='tiny', batch=3, seed=107, coords='corners',
         gen=dict(bg_bias=2.0, hot=10, offset_sigma=0.8),
         kwargs=dict(confidence_thresh=0.05, iou_threshold=0.4, top_k=20, input_coords='corners',
                     normalize_coords=True, img_height=96, img_width=128)),
    dict(name='d_tiny_nonorm_topk5', fn='decode_detections', layout='tiny', batch=4, seed=108,
         gen=dict(bg_bias=1.5, hot=10), kill=2,
         kwargs=dict(confidence_thresh=0.05, iou_threshold=0.45, top_k=5, input_coords='centroids',
                     normalize_coords=False)),
    dict(name='d_tiny_f64', fn='decode_detections', layout='tiny', batch=2, seed=109, dtype='float64',
         gen=dict(bg_bias=2.0, hot=10),
         kwargs=dict(confidence_thresh=0.03, iou_threshold=0.45, top_k=50, input_coords='centroids',
                     normalize_coords=True, img_height=96, img_width=128, border_pixels='exclude')),
    dict(name='d_tiny_nolog', fn='decode_detections', layout='tiny', batch=2, seed=110, log_wh=False,
         gen=dict(bg_bias=2.0, hot=10, offset_sigma=2.0),
         kwargs=dict(confidence_thresh=0.05, iou_threshold=0.45, top_k=200, input_coords='centroids',
                     normalize_coords=True, img_height=96, img_width=128)),
    dict(name='f_ssd300_fast', fn='decode_detections_fast', layout='ssd300', batch=2, seed=111,
         gen=dict(bg_bias=1.0, hot=60),
         kwargs=dict(confidence_thresh=0.2, iou_threshold=0.45, top_k='all', input_coords='centroids',
                     normalize_coords=True, img_height=300, img_width=300)),
    dict(name='f_tiny_fast_nonms_topk', fn='decode_detections_fast', layout='tiny', batch=3, seed=112,
         gen=dict(bg_bias=0.5, hot=10), kill=1,
         kwargs=dict(confidence_thresh=0.15, iou_threshold=None, top_k=7, input_coords='centroids',
                     normalize_coords=True, img_height=96, img_width=128)),
    dict(name='f_tiny_fast_nonms_all', fn='decode_detections_fast', layout='tiny', batch=2, seed=113,
         gen=dict(bg_bias=0.5, hot=10),
         kwargs=dict(confidence_thresh=0.15, iou_threshold=0, top_k='all', input_coords='centroids',
                     normalize_coords=True, img_height=96, img_width=128)),
    dict(name='f_tiny_fast_corners_f32', fn='decode_detections_fast', layout='tiny', batch=2, seed=114, coords='corners',
         gen=dict(bg_bias=0.5, hot=10),
         kwargs=dict(confidence_thresh=0.15, iou_threshold=0.45, top_k=10, input_coords='corners',
                     normalize_coords=True, img_height=96, img_width=128)),
    dict(name='f_roundtrip_f64', fn='decode_detections_fast', layout='ssd300', batch=3, seed=115, roundtrip=True,
         kwargs=dict(confidence_thresh=0.5, iou_threshold=0.45, top_k='all', input_coords='centroids',
                     normalize_coords=True, img_height=300, img_width=300)),
]


def build_decode_input(case, encoder_cls):
    """Returns the y_pred tensor of a decode case.  `encoder_cls` supplies the anchors (any
    implementation: they are bit-identical, which the golden SHA-256 verifies)."""
    if case.get('roundtrip'):
        enc = synth.make_encoder(encoder_cls, case['layout'])
        gt = synth.synth_ground_truth(enc.img_height, enc.img_width, enc.n_classes - 1, case['batch'], case['seed'])
        return np.ascontiguousarray(enc(gt))
    dtype = np.dtype(case.get('dtype', 'float32'))
    return _ypred(encoder_cls, case['layout'], case['batch'], case['seed'], coords=case.get('coords', 'centroids'),
                  dtype=dtype, quantize=case.get('quantize'), kill=case.get('kill'), **case.get('gen', {}))


def canonical_rows(per_image, width=7):
    """list of (k_i, width) arrays -> (rows sorted per image by class asc, score desc, anchor
    asc; counts).  Column layout: [anchor, class, conf, xmin, ymin, xmax, ymax]."""
    counts = np.array([0 if np.size(p) == 0 else p.shape[0] for p in per_image], dtype=np.int64)
    blocks = []
    for p in per_image:
        if np.size(p) == 0:
            continue
        p = np.asarray(p, dtype=np.float64)
        order = np.lexsort((p[:, 0], -p[:, 2], p[:, 1]))
        blocks.append(p[order])
    rows = np.concatenate(blocks, axis=0) if blocks else np.zeros((0, width))
    return rows, counts


# ----------------------------------------------------------------------------------------
# encode cases
# ----------------------------------------------------------------------------------------
def _quirk_gt():
    # tiny layout is 96 (h) x 128 (w)
    far = np.array([[1, 500, 500, 560, 540]], dtype=np.float64)                 # no overlap with any anchor
    return [
        np.array([[2, 10, 10, 60, 50], [2, 10, 10, 60, 50], [1, 70, 20, 120, 90]], dtype=np.float64),  # duplicate GT
        np.zeros((0, 5)),                                                          # empty image
        np.concatenate([np.array([[3, 30, 30, 90, 80]], dtype=np.float64), far, far + [1, 0, 0, 0, 0]]),  # zero rows
        np.array([[1, 0, 0, 128, 96]], dtype=np.float64),                          # whole image, m = 1
        np.concatenate([far, np.array([[2, 5, 5, 40, 40], [3, 6, 6, 41, 41]], dtype=np.float64)]),    # zero row first
        np.array([[1, 2, 2, 7, 7], [2, 3, 3, 8, 8], [3, 100, 60, 127, 95], [1, 64, 48, 66, 50]], dtype=np.float64),
    ]


ENCODE_CASES = [
    dict(name='e_ssd300_b4', layout='ssd300', batch=4, seed=201),
    dict(name='e_ssd300_neg03', layout='ssd300', batch=2, seed=202, overrides=dict(neg_iou_limit=0.3)),
    dict(name='e_ssd512_b1', layout='ssd512', batch=1, seed=203),
    dict(name='e_tiny_diag', layout='tiny', batch=5, seed=204, diagnostics=True),
    dict(name='e_tiny_corners', layout='tiny', batch=4, seed=205, overrides=dict(coords='corners')),
    dict(name='e_tiny_minmax_incl', layout='tiny', batch=4, seed=206, overrides=dict(coords='minmax', border_pixels='include')),
    dict(name='e_tiny_bipartite', layout='tiny', batch=4, seed=207, overrides=dict(matching_type='bipartite')),
    dict(name='e_tiny_nonorm', layout='tiny', batch=4, seed=208, overrides=dict(normalize_coords=False)),
    dict(name='e_tiny_nolog', layout='tiny', batch=4, seed=209, log_wh=False),
    dict(name='e_tiny_clip', layout='tiny', batch=4, seed=210, overrides=dict(clip_boxes=True)),
    dict(name='e_tiny_quirks', layout='tiny', gt='quirks'),
    dict(name='e_tiny_bg1', layout='tiny', batch=3, seed=211, overrides=dict(background_id=1, neg_iou_limit=0.2)),
]


def build_encode_input(case):
    kw = synth.layout_kwargs(case['layout'], **case.get('overrides', {}))
    if case.get('gt') == 'quirks':
        return _quirk_gt()
    gt = synth.synth_ground_truth(kw['img_height'], kw['img_width'], kw['n_classes'], case['batch'], case['seed'],
                                  max_boxes=case.get('max_boxes', 20))
    if kw.get('background_id', 0) != 0:
        # classes must avoid the background id
        for g in gt:
            g[:, 0] = np.where(g[:, 0] == kw['background_id'], 0, g[:, 0])
    return gt


# ----------------------------------------------------------------------------------------
# thin-op cases (inputs are small and stored inline in the golden file)
# ----------------------------------------------------------------------------------------
def thin_inputs(seed=301):
    rng = np.random.default_rng(seed)
    def boxes(n, fmt):
        x0 = rng.uniform(0, 80, n); y0 = rng.uniform(0, 60, n)
        w = rng.uniform(1, 60, n); h = rng.uniform(1, 50, n)
        if fmt == 'corners':
            return np.stack([x0, y0, x0 + w, y0 + h], 1)
        if fmt == 'minmax':
            return np.stack([x0, x0 + w, y0, y0 + h], 1)
        return np.stack([x0 + w / 2, y0 + h / 2, w, h], 1)
    out = {}
    for fmt in ('corners', 'minmax', 'centroids'):
        out['b1_' + fmt] = boxes(7, fmt)
        out['b2_' + fmt] = boxes(11, fmt)
        out['b3_' + fmt] = boxes(7, fmt)
    out['w_small'] = rng.uniform(0, 1, (5, 40))
    w = rng.uniform(0, 1, (6, 50))
    w[:, rng.integers(0, 50, 20)] = 0
    w[2] = 0                              # an all-zero row: the re-match quirk
    w[4, 7] = w[1, 7] = 0.99              # two rows share the best column
    out['w_quirk'] = w
    out['conv32'] = rng.uniform(0, 100, (3, 5, 9)).astype(np.float32)
    out['conv64'] = rng.uniform(0, 100, (4, 6))
    nb = boxes(300, 'corners')
    out['nms_rows'] = np.concatenate([rng.integers(1, 4, (300, 1)).astype(float),
                                      np.round(rng.uniform(0, 1, (300, 1)), 2), nb], axis=1)
    return out


# ----------------------------------------------------------------------------------------
# VOC evaluation cases
# ----------------------------------------------------------------------------------------
VOC_CASES = [
    dict(name='voc_basic', seed=401, n_images=60, n_classes=5, kwargs=dict(matching_iou_threshold=0.5, border_pixels='include', sorting_algorithm='mergesort')),
    dict(name='voc_neutral', seed=402, n_images=40, n_classes=4, neutral=True, kwargs=dict(matching_iou_threshold=0.5, border_pixels='include', sorting_algorithm='mergesort')),
    dict(name='voc_ties_half', seed=403, n_images=30, n_classes=3, quantize=20, kwargs=dict(matching_iou_threshold=0.3, border_pixels='half', sorting_algorithm='mergesort')),
    dict(name='voc_under_area', seed=404, n_images=30, n_classes=3, ignore_under_area=900, kwargs=dict(matching_iou_threshold=0.5, border_pixels='exclude', sorting_algorithm='mergesort')),
    dict(name='voc_quiet_quirk', seed=405, n_images=20, n_classes=3, verbose=False, kwargs=dict(matching_iou_threshold=0.5, border_pixels='include', sorting_algorithm='mergesort')),
    dict(name='voc_large', seed=406, n_images=400, n_classes=20, dets_per_image=60, kwargs=dict(matching_iou_threshold=0.5, border_pixels='include', sorting_algorithm='mergesort')),
]


def build_voc_input(case):
    """Synthetic dataset + detections: integer ground-truth boxes, detections = jittered copies of some of
    them (several per object: duplicates), plus random false positives, random confidences."""
    rng = np.random.default_rng(case['seed'])
    n_images, C = case['n_images'], case['n_classes']
    labels, neutral, image_ids = [], [], []
    for i in range(n_images):
        m = int(rng.integers(0 if i % 7 == 3 else 1, 9))
        w = rng.integers(12, 200, size=m); h = rng.integers(12, 200, size=m)
        x0 = rng.integers(0, 300 - 10, size=m); y0 = rng.integers(0, 300 - 10, size=m)
        cls = rng.integers(1, C + 1, size=m)
        labels.append(np.stack([cls, x0, y0, x0 + w, y0 + h], axis=1).astype(np.int64).reshape(m, 5))
        neutral.append(rng.uniform(size=m) < 0.25)
        image_ids.append('%06d' % (1000 + 3 * i))
    preds = [[] for _ in range(C + 1)]
    per_img = case.get('dets_per_image', 14)
    for i in range(n_images):
        lab = labels[i]
        for _ in range(per_img):
            if lab.shape[0] and rng.uniform() < 0.7:
                g = lab[int(rng.integers(0, lab.shape[0]))]
                jit = rng.normal(0, 6.0, size=4)
                box = g[1:5] + jit
                c = int(g[0]) if rng.uniform() < 0.85 else int(rng.integers(1, C + 1))
            else:
                x0, y0 = rng.uniform(0, 250, size=2)
                box = np.array([x0, y0, x0 + rng.uniform(5, 150), y0 + rng.uniform(5, 150)])
                c = int(rng.integers(1, C + 1))
            conf = float(rng.uniform(0.01, 1.0))
            if case.get('quantize'):
                conf = round(conf * case['quantize']) / case['quantize']
            preds[c].append((image_ids[i], conf, round(float(box[0]), 1), round(float(box[1]), 1),
                             round(float(box[2]), 1), round(float(box[3]), 1)))
    return dict(labels=labels, eval_neutral=neutral if case.get('neutral') else None, image_ids=image_ids,
                prediction_results=preds, n_classes=C)


# ----------------------------------------------------------------------------------------
# decoder -> evaluator glue and BoxFilter: seeded inputs
# ----------------------------------------------------------------------------------------
def build_evalprep_input(seed=11, n_images=9):
    """Decoded predictions of a batch (float64 (k_i, 6), one image empty), per-image inverter specifications
    ('resize', H, W, out_h, out_w) / ('translate', dy, dx) / ('identity',) / None, and label arrays for BoxFilter."""
    rng = np.random.default_rng(seed)
    preds, specs = [], []
    for i in range(n_images):
        k = 0 if i == 3 else int(rng.integers(1, 40))
        p = np.zeros((k, 6))
        p[:, 0] = rng.integers(1, 21, size=k)
        p[:, 1] = rng.uniform(0.01, 1.0, size=k)
        x0 = rng.uniform(-20, 280, size=k); y0 = rng.uniform(-20, 280, size=k)
        p[:, 2], p[:, 3] = x0, y0
        p[:, 4], p[:, 5] = x0 + rng.uniform(1, 150, size=k), y0 + rng.uniform(1, 150, size=k)
        # values that sit on rounding boundaries (x.5 after scaling, x.x5 for the one-decimal rounding)
        if k > 2:
            p[0, 2:6] = [10.25, 20.75, 30.05, 40.15]
            p[1, 2:6] = [0.5, 1.5, 2.5, 3.5]
        preds.append(p)
        H, W = int(rng.integers(200, 700)), int(rng.integers(200, 700))
        kind = i % 4
        if kind == 0:
            specs.append([('resize', H, W, 300, 300)])
        elif kind == 1:
            specs.append([('translate', int(rng.integers(-50, 50)), int(rng.integers(-50, 50))), ('resize', H, W, 300, 300)])
        elif kind == 2:
            specs.append([('identity',), None, ('resize', 2 * H, W, 300, 300), ('translate', 7, -3)])
        else:
            specs.append([])
    labels = []
    for i in range(12):
        m = 0 if i == 5 else int(rng.integers(1, 25))
        l = np.zeros((m, 5))
        l[:, 0] = rng.integers(1, 21, size=m)
        x0 = np.floor(rng.uniform(-60, 320, size=m)); y0 = np.floor(rng.uniform(-60, 320, size=m))
        l[:, 1], l[:, 2] = x0, y0
        l[:, 3] = x0 + np.floor(rng.uniform(-2, 120, size=m))        # some degenerate, some tiny
        l[:, 4] = y0 + np.floor(rng.uniform(-2, 120, size=m))
        labels.append(l)
    return dict(preds=preds, specs=specs, labels=labels)


BOXFILTER_CONFIGS = [
    dict(),
    dict(overlap_criterion='iou', overlap_bounds=(0.1, 0.9)),
    dict(overlap_criterion='area', overlap_bounds=(0.3, 1.0)),
    dict(overlap_criterion='area', overlap_bounds=(0.0, 1.0), border_pixels='include'),
    dict(overlap_criterion='iou', overlap_bounds=(0.0, 0.5), border_pixels='exclude', check_min_area=False),
    dict(check_overlap=False, min_area=400),
    dict(overlap_criterion='center_point', check_degenerate=False, check_min_area=False),
]
