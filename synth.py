"""Seeded synthetic workloads of the named anchor layouts.

Pure numpy, host side.  Shared by bench.py, the tests and the golden-vector generator so
every party sees byte-identical inputs for a given (layout, seed).
"""
from __future__ import division

import numpy as np

LAYOUTS = {
    # training_dct_pascal_j2d_resnet.py:92-111,251-265 (ssd_custom == SSD300 layout, 8732 anchors)
    'ssd300': dict(img_height=300, img_width=300, n_classes=20,
                   predictor_sizes=[(38, 38), (19, 19), (10, 10), (5, 5), (3, 3), (1, 1)],
                   scales=[0.1, 0.2, 0.37, 0.54, 0.71, 0.88, 1.05],
                   aspect_ratios_per_layer=[[1.0, 2.0, 0.5], [1.0, 2.0, 0.5, 3.0, 1.0 / 3.0],
                                            [1.0, 2.0, 0.5, 3.0, 1.0 / 3.0], [1.0, 2.0, 0.5, 3.0, 1.0 / 3.0],
                                            [1.0, 2.0, 0.5], [1.0, 2.0, 0.5]],
                   two_boxes_for_ar1=True, steps=[8, 16, 32, 64, 100, 300], offsets=[0.5] * 6,
                   clip_boxes=False, variances=[0.1, 0.1, 0.2, 0.2], matching_type='multi',
                   pos_iou_threshold=0.5, neg_iou_limit=0.5, normalize_coords=True),
    # not in the reference; synthesised through the same generic encoder (24564 anchors)
    'ssd512': dict(img_height=512, img_width=512, n_classes=20,
                   predictor_sizes=[(64, 64), (32, 32), (16, 16), (8, 8), (4, 4), (2, 2), (1, 1)],
                   scales=[0.07, 0.15, 0.3, 0.45, 0.6, 0.75, 0.9, 1.05],
                   aspect_ratios_per_layer=[[1.0, 2.0, 0.5]] + [[1.0, 2.0, 0.5, 3.0, 1.0 / 3.0]] * 4 + [[1.0, 2.0, 0.5]] * 2,
                   two_boxes_for_ar1=True, steps=[8, 16, 32, 64, 128, 256, 512], offsets=[0.5] * 7,
                   clip_boxes=False, variances=[0.1, 0.1, 0.2, 0.2], matching_type='multi',
                   pos_iou_threshold=0.5, neg_iou_limit=0.5, normalize_coords=True),
    # a small layout for fast oracle runs and edge-case tests (3 classes + background)
    'tiny': dict(img_height=96, img_width=128, n_classes=3,
                 predictor_sizes=[(6, 8), (3, 4), (1, 1)], scales=[0.15, 0.4, 0.7, 1.0],
                 aspect_ratios_per_layer=[[1.0, 2.0, 0.5], [1.0, 2.0, 0.5, 3.0, 1.0 / 3.0], [1.0, 2.0, 0.5]],
                 two_boxes_for_ar1=True, steps=None, offsets=None,
                 clip_boxes=False, variances=[0.1, 0.1, 0.2, 0.2], matching_type='multi',
                 pos_iou_threshold=0.5, neg_iou_limit=0.3, normalize_coords=True),
}


def layout_kwargs(name, **overrides):
    kw = dict(LAYOUTS[name])
    kw.update(overrides)
    return kw


def make_encoder(encoder_cls, name, **overrides):
    """Instantiate any `SSDInputEncoder`-compatible class (product, oracle or reference)."""
    return encoder_cls(**layout_kwargs(name, **overrides))


def anchors_of(encoder):
    """(A, 4) float64 anchors of an encoder instance (any of the three implementations)."""
    return np.concatenate([np.asarray(b).reshape(-1, 4) for b in encoder.boxes_list], axis=0)


def synth_y_pred(anchors, variances, n_classes_incl_bg, batch, seed, bg_bias=9.0, hot=40,
                 offset_sigma=0.5, exp_free=False, dtype=np.float32):
    """Synthetic raw SSD output `(batch, A, C+12)`: softmax of N(0,1) logits with a background
    bias `bg_bias` (controls the candidate density) and `hot` boosted (anchor, class) logits per
    image; N(0, offset_sigma) box offsets; anchors and variances cast like AnchorBoxes does
    (keras_layer_AnchorBoxes.py:253).  `exp_free`: zero w/h offsets so that exp(0) == 1 exactly and
    the whole decode chain is IEEE-exact on every implementation."""
    rng = np.random.default_rng(seed)
    A = anchors.shape[0]
    C = n_classes_incl_bg
    logits = rng.standard_normal((batch, A, C), dtype=np.float32)
    logits[:, :, 0] += np.float32(bg_bias)
    if hot > 0:
        for b in range(batch):
            ai = rng.integers(0, A, size=hot)
            ci = rng.integers(1, C, size=hot)
            logits[b, ai, ci] += rng.uniform(bg_bias - 2.0, bg_bias + 4.0, size=hot).astype(np.float32)
    # softmax in float64, rounded once to float32: float32 np.exp is SIMD dispatched and host
    # dependent at the ulp level, which would make the "same" seeded input differ between hosts
    l64 = logits.astype(np.float64)
    l64 -= l64.max(axis=-1, keepdims=True)
    e = np.exp(l64)
    conf = (e / e.sum(axis=-1, keepdims=True)).astype(np.float32)
    off = (rng.standard_normal((batch, A, 4), dtype=np.float32) * np.float32(offset_sigma)).astype(np.float32)
    if exp_free:
        off[:, :, 2:] = 0
    y = np.empty((batch, A, C + 12), dtype=np.float32)
    y[:, :, :C] = conf
    y[:, :, C:C + 4] = off
    y[:, :, C + 4:C + 8] = anchors.astype(np.float32)[None]
    y[:, :, C + 8:] = np.asarray(variances, dtype=np.float32)[None, None]
    return y.astype(dtype, copy=False)


def synth_ground_truth(img_height, img_width, n_classes, batch, seed, max_boxes=20, min_boxes=1):
    """VOC-like ground truth: per image 1..max_boxes boxes with integer pixel corners
    (object_detection_2d_data_generator_dct_j2d.py:509-512), rows `[class, xmin, ymin, xmax, ymax]`."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(batch):
        m = int(rng.integers(min_boxes, max_boxes + 1))
        w = rng.uniform(15, 0.93 * img_width, size=m)
        h = rng.uniform(15, 0.93 * img_height, size=m)
        x0 = rng.uniform(0, img_width - w)
        y0 = rng.uniform(0, img_height - h)
        xmin = np.floor(x0)
        ymin = np.floor(y0)
        xmax = np.floor(x0 + w) + 1
        ymax = np.floor(y0 + h) + 1
        cls = rng.integers(1, n_classes + 1, size=m)
        out.append(np.stack([cls, xmin, ymin, xmax, ymax], axis=1).astype(np.float64))
    return out


def flatten_ground_truth(gt_list):
    """list of (m_i, 5) -> (sum m_i, 5) float64 and (B+1,) int64 offsets (the C ABI layout)."""
    offs = np.zeros(len(gt_list) + 1, dtype=np.int64)
    parts = []
    for i, g in enumerate(gt_list):
        g = np.asarray(g, dtype=np.float64)
        m = 0 if g.size == 0 else g.shape[0]
        offs[i + 1] = offs[i] + m
        if m:
            parts.append(g[:, :5])
    flat = np.ascontiguousarray(np.concatenate(parts, axis=0)) if parts else np.zeros((0, 5))
    return flat, offs
