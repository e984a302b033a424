"""CPU oracle for the SSD box codec hot path.  TEST INFRASTRUCTURE ONLY.

This module is a from-scratch numpy restatement of the algorithm of the
reference's box codec (Shulk97/JPEG_detection_Resnet_SSD, `localisation_part/`).
It is the *checker* for the CUDA path: only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.  The
product package (`jpeg_detection_resnet_ssd_b200/`) never does; it fails loudly
when the CUDA library is missing.

Parity pin: `oracle/make_golden.py` runs the *real* reference (imported from
`/root/reference/localisation_part` with the `np.float = float; np.int = int`
alias shim) and this restatement on the same seeded inputs and requires
bit-identical outputs (`exp_mode='numpy'`), then commits the vectors under
`tests/golden/`.  `tests/test_oracle_golden.py` re-checks the restatement
against those committed vectors on every run.  The Keras-layer contract
(`decode_layer*`) has no runnable reference here (TensorFlow absent): that part
is **parity unpinned** and only follows the cited source lines.

Every function cites the reference lines it follows (paths relative to
`/root/reference/localisation_part/`).

`exp_mode`:
  * 'numpy' - float32 `np.exp`, exactly what the reference executes (SIMD
    dispatched, not correctly rounded, host dependent at the ulp level).
  * 'cr'    - correctly rounded float32 exp, computed as float32(exp(float64)).
    This is the definition the CUDA kernels implement, so index parity against
    this mode is bit-exact by construction and host independent.
"""
from __future__ import division

import numpy as np

_BORDER_D = {'half': 0, 'include': 1, 'exclude': -1}


# ----------------------------------------------------------------------------
# bounding_box_utils/bounding_box_utils.py
# ----------------------------------------------------------------------------

def convert_coordinates(tensor, start_index, conversion, border_pixels='half'):
    """bounding_box_utils.py:24-87.  Returns a float64 copy; the arithmetic is
    carried out in the *input* dtype and then stored (line 60 + 62-83)."""
    d = _BORDER_D.get(border_pixels)
    s = start_index
    out = np.array(tensor, dtype=np.float64, copy=True)
    t = tensor
    if conversion == 'minmax2centroids':
        out[..., s] = (t[..., s] + t[..., s + 1]) / 2.0
        out[..., s + 1] = (t[..., s + 2] + t[..., s + 3]) / 2.0
        out[..., s + 2] = t[..., s + 1] - t[..., s] + d
        out[..., s + 3] = t[..., s + 3] - t[..., s + 2] + d
    elif conversion == 'centroids2minmax':
        out[..., s] = t[..., s] - t[..., s + 2] / 2.0
        out[..., s + 1] = t[..., s] + t[..., s + 2] / 2.0
        out[..., s + 2] = t[..., s + 1] - t[..., s + 3] / 2.0
        out[..., s + 3] = t[..., s + 1] + t[..., s + 3] / 2.0
    elif conversion == 'corners2centroids':
        out[..., s] = (t[..., s] + t[..., s + 2]) / 2.0
        out[..., s + 1] = (t[..., s + 1] + t[..., s + 3]) / 2.0
        out[..., s + 2] = t[..., s + 2] - t[..., s] + d
        out[..., s + 3] = t[..., s + 3] - t[..., s + 1] + d
    elif conversion == 'centroids2corners':
        out[..., s] = t[..., s] - t[..., s + 2] / 2.0
        out[..., s + 1] = t[..., s + 1] - t[..., s + 3] / 2.0
        out[..., s + 2] = t[..., s] + t[..., s + 2] / 2.0
        out[..., s + 3] = t[..., s + 1] + t[..., s + 3] / 2.0
    elif conversion in ('minmax2corners', 'corners2minmax'):
        out[..., s + 1] = t[..., s + 2]
        out[..., s + 2] = t[..., s + 1]
    else:
        raise ValueError("Unexpected conversion value. Supported values are 'minmax2centroids', 'centroids2minmax', 'corners2centroids', 'centroids2corners', 'minmax2corners', and 'corners2minmax'.")
    return out


def _axis_ids(coords):
    # bounding_box_utils.py:353-362
    if coords == 'corners':
        return 0, 1, 2, 3  # xmin, ymin, xmax, ymax
    return 0, 2, 1, 3      # minmax: xmin at 0, ymin at 2, xmax at 1, ymax at 3


def _intersection(b1, b2, coords, mode):
    """bounding_box_utils.py:226-280 as called from :345 (border_pixels is NOT
    forwarded there, so the intersection always uses d = 0)."""
    ix0, iy0, ix1, iy1 = _axis_ids(coords)
    d = 0
    if mode == 'outer_product':
        lo = np.maximum(b1[:, None, [ix0, iy0]], b2[None, :, [ix0, iy0]])
        hi = np.minimum(b1[:, None, [ix1, iy1]], b2[None, :, [ix1, iy1]])
        side = np.maximum(0, hi - lo + d)
        return side[:, :, 0] * side[:, :, 1]
    lo = np.maximum(b1[:, [ix0, iy0]], b2[:, [ix0, iy0]])
    hi = np.minimum(b1[:, [ix1, iy1]], b2[:, [ix1, iy1]])
    side = np.maximum(0, hi - lo + d)
    return side[:, 0] * side[:, 1]


def iou(boxes1, boxes2, coords='centroids', mode='outer_product', border_pixels='half'):
    """bounding_box_utils.py:283-383."""
    if boxes1.ndim > 2:
        raise ValueError("boxes1 must have rank either 1 or 2, but has rank {}.".format(boxes1.ndim))
    if boxes2.ndim > 2:
        raise ValueError("boxes2 must have rank either 1 or 2, but has rank {}.".format(boxes2.ndim))
    if boxes1.ndim == 1:
        boxes1 = boxes1[None, :]
    if boxes2.ndim == 1:
        boxes2 = boxes2[None, :]
    if not (boxes1.shape[1] == boxes2.shape[1] == 4):
        raise ValueError("All boxes must consist of 4 coordinates, but the boxes in `boxes1` and `boxes2` have {} and {} coordinates, respectively.".format(boxes1.shape[1], boxes2.shape[1]))
    if mode not in ('outer_product', 'element-wise'):
        raise ValueError("`mode` must be one of 'outer_product' and 'element-wise', but got '{}'.".format(mode))
    if coords == 'centroids':
        boxes1 = convert_coordinates(boxes1, 0, 'centroids2corners')
        boxes2 = convert_coordinates(boxes2, 0, 'centroids2corners')
        coords = 'corners'
    elif coords not in ('minmax', 'corners'):
        raise ValueError("Unexpected value for `coords`. Supported values are 'minmax', 'corners' and 'centroids'.")

    inter = _intersection(boxes1, boxes2, coords, mode)
    ix0, iy0, ix1, iy1 = _axis_ids(coords)
    d = _BORDER_D.get(border_pixels)
    a1 = (boxes1[:, ix1] - boxes1[:, ix0] + d) * (boxes1[:, iy1] - boxes1[:, iy0] + d)
    a2 = (boxes2[:, ix1] - boxes2[:, ix0] + d) * (boxes2[:, iy1] - boxes2[:, iy0] + d)
    if mode == 'outer_product':
        union = a1[:, None] + a2[None, :] - inter
    else:
        union = a1 + a2 - inter
    with np.errstate(divide='ignore', invalid='ignore'):
        return inter / union


# ----------------------------------------------------------------------------
# ssd_encoder_decoder/ssd_output_decoder.py
# ----------------------------------------------------------------------------

def greedy_nms_rows(rows, score_col, box_col, iou_threshold, border_pixels='half'):
    """ssd_output_decoder.py:77-92 / :94-109 / :469-486 (one routine; the three
    reference variants differ only in which column holds the score and where
    the four corner coordinates start).  Keeps the reference's cost structure:
    one Python iteration and a handful of numpy calls per kept box."""
    left = np.copy(rows)
    kept = []
    while left.shape[0] > 0:
        top = np.argmax(left[:, score_col])          # first maximum
        best = np.copy(left[top])
        kept.append(best)
        left = np.delete(left, top, axis=0)
        if left.shape[0] == 0:
            break
        sim = iou(left[:, box_col:box_col + 4], best[box_col:box_col + 4],
                  coords='corners', mode='element-wise', border_pixels=border_pixels)
        left = left[sim <= iou_threshold]           # NaN compares False => dropped
    return np.array(kept)


def greedy_nms(y_pred_decoded, iou_threshold=0.45, coords='corners', border_pixels='half'):
    """ssd_output_decoder.py:27-75: rows `[class_id, score, 4 coords]`."""
    out = []
    for item in y_pred_decoded:
        left = np.copy(item)
        kept = []
        while left.shape[0] > 0:
            top = np.argmax(left[:, 1])
            best = np.copy(left[top])
            kept.append(best)
            left = np.delete(left, top, axis=0)
            if left.shape[0] == 0:
                break
            sim = iou(left[:, 2:], best[2:], coords=coords, mode='element-wise', border_pixels=border_pixels)
            left = left[sim <= iou_threshold]
        out.append(np.array(kept))
    return out


def _exp(x, exp_mode):
    if exp_mode == 'numpy' or x.dtype != np.float32:
        return np.exp(x)
    if exp_mode == 'cr':
        return np.exp(x.astype(np.float64)).astype(np.float32)
    raise ValueError(exp_mode)


def _decode_offsets(coordpart, y_pred, input_coords, log_wh, exp_mode):
    """Shared by decode_detections (:174-192) and decode_detections_fast
    (:296-314).  `coordpart` is a writable view of the four offset columns, in
    y_pred's dtype; it is updated in place exactly in the reference's order.
    Returns True when the caller must still run the centroid/minmax -> corner
    conversion (which upcasts the whole tensor to float64)."""
    anc = y_pred[:, :, -8:-4]
    var = y_pred[:, :, -4:]
    if input_coords == 'centroids':
        wh = coordpart[:, :, 2:4] * var[:, :, 2:4]
        if log_wh:
            wh = _exp(wh, exp_mode)                      # *_no_log twin drops the exp (:175)
        coordpart[:, :, 2:4] = wh
        coordpart[:, :, 2:4] *= anc[:, :, 2:4]
        coordpart[:, :, 0:2] *= var[:, :, 0:2] * anc[:, :, 2:4]
        coordpart[:, :, 0:2] += anc[:, :, 0:2]
        return 'centroids2corners'
    if input_coords == 'minmax':
        coordpart *= var
        coordpart[:, :, 0:2] *= (y_pred[:, :, -7] - y_pred[:, :, -8])[..., None]
        coordpart[:, :, 2:4] *= (y_pred[:, :, -5] - y_pred[:, :, -6])[..., None]
        coordpart += anc
        return 'minmax2corners'
    if input_coords == 'corners':
        coordpart *= var
        coordpart[:, :, [0, 2]] *= (y_pred[:, :, -6] - y_pred[:, :, -8])[..., None]
        coordpart[:, :, [1, 3]] *= (y_pred[:, :, -5] - y_pred[:, :, -7])[..., None]
        coordpart += anc
        return None
    return 'bad'


def decode_detections(y_pred, confidence_thresh=0.01, iou_threshold=0.45, top_k=200,
                      input_coords='centroids', normalize_coords=True, img_height=None,
                      img_width=None, border_pixels='half', log_wh=True, exp_mode='numpy',
                      with_anchor_index=False):
    """ssd_output_decoder.py:111-226 (and :342-467 when `with_anchor_index`,
    which prepends the anchor id like `decode_detections_debug`)."""
    if normalize_coords and ((img_height is None) or (img_width is None)):
        raise ValueError("If relative box coordinates are supposed to be converted to absolute coordinates, the decoder needs the image size in order to decode the predictions, but `img_height == {}` and `img_width == {}`".format(img_height, img_width))

    raw = np.copy(y_pred[:, :, :-8])
    conv = _decode_offsets(raw[:, :, -4:], y_pred, input_coords, log_wh, exp_mode)
    if conv == 'bad':
        raise ValueError("Unexpected value for `input_coords`. Supported input coordinate formats are 'minmax', 'corners' and 'centroids'.")
    if conv is not None:
        raw = convert_coordinates(raw, start_index=-4, conversion=conv)
    if normalize_coords:
        raw[:, :, [-4, -2]] *= img_width
        raw[:, :, [-3, -1]] *= img_height

    n_classes = raw.shape[-1] - 4
    n_boxes = raw.shape[1]
    ids = np.arange(n_boxes, dtype=raw.dtype)
    results = []
    for item in raw:
        per_class = []
        for cid in range(1, n_classes):
            cand = item[:, [cid, -4, -3, -2, -1]]
            mask = cand[:, 0] > confidence_thresh
            cand = cand[mask]
            if cand.shape[0] > 0:
                if with_anchor_index:
                    cand = np.concatenate([cand, ids[mask][:, None]], axis=1)
                keep = greedy_nms_rows(cand, 0, 1, iou_threshold, border_pixels)
                rows = np.zeros((keep.shape[0], keep.shape[1] + 1))
                rows[:, 0] = cid
                rows[:, 1:] = keep
                per_class.append(rows)
        if per_class:
            pred = np.concatenate(per_class, axis=0)
            if top_k != 'all' and pred.shape[0] > top_k:
                sel = np.argpartition(pred[:, 1], kth=pred.shape[0] - top_k, axis=0)[pred.shape[0] - top_k:]
                pred = pred[sel]
        else:
            pred = np.array(per_class)
        results.append(pred)
    return results


def decode_detections_fast(y_pred, confidence_thresh=0.5, iou_threshold=0.45, top_k='all',
                           input_coords='centroids', normalize_coords=True, img_height=None,
                           img_width=None, border_pixels='half', log_wh=True, exp_mode='numpy',
                           with_anchor_index=False):
    """ssd_output_decoder.py:228-333."""
    if normalize_coords and ((img_height is None) or (img_width is None)):
        raise ValueError("If relative box coordinates are supposed to be converted to absolute coordinates, the decoder needs the image size in order to decode the predictions, but `img_height == {}` and `img_width == {}`".format(img_height, img_width))

    conv6 = np.copy(y_pred[:, :, -14:-8])
    conv6[:, :, 0] = np.argmax(y_pred[:, :, :-12], axis=-1)
    conv6[:, :, 1] = np.amax(y_pred[:, :, :-12], axis=-1)
    conv = _decode_offsets(conv6[:, :, 2:], y_pred, input_coords, log_wh, exp_mode)
    if conv == 'bad':
        raise ValueError("Unexpected value for `coords`. Supported values are 'minmax', 'corners' and 'centroids'.")
    if conv is not None:
        conv6 = convert_coordinates(conv6, start_index=-4, conversion=conv)
    if normalize_coords:
        conv6[:, :, [2, 4]] *= img_width
        conv6[:, :, [3, 5]] *= img_height

    n_boxes = conv6.shape[1]
    ids = np.arange(n_boxes, dtype=conv6.dtype)
    results = []
    for item in conv6:
        if with_anchor_index:
            item = np.concatenate([item, ids[:, None]], axis=1)
        boxes = item[np.nonzero(item[:, 0])]
        boxes = boxes[boxes[:, 1] >= confidence_thresh]
        if iou_threshold:
            boxes = greedy_nms_rows(boxes, 1, 2, iou_threshold, border_pixels)
        if top_k != 'all' and boxes.shape[0] > top_k:
            sel = np.argpartition(boxes[:, 1], kth=boxes.shape[0] - top_k, axis=0)[boxes.shape[0] - top_k:]
            boxes = boxes[sel]
        results.append(boxes)
    return results


# ----------------------------------------------------------------------------
# keras_layers/keras_layer_DecodeDetections{,Fast}.py  -- PARITY UNPINNED
# ----------------------------------------------------------------------------

def _tf_iou(a, b):
    """IoU as computed by TensorFlow 1.x's NonMaxSuppression CPU kernel
    (third-party dependency `tensorflow-gpu`, pinned 1.8.0 in the Pipfile /
    1.14.0 in Pipfile.lock; source not in /root/reference).  Boxes are
    (ymin, xmin, ymax, xmax) float32; the kernel re-orders each box's corners
    with min/max, returns 0 when either area is <= 0 and otherwise
    inter / (area_a + area_b - inter), all in float32."""
    f = np.float32
    ymin_a, xmin_a = min(a[0], a[2]), min(a[1], a[3])
    ymax_a, xmax_a = max(a[0], a[2]), max(a[1], a[3])
    ymin_b, xmin_b = min(b[0], b[2]), min(b[1], b[3])
    ymax_b, xmax_b = max(b[0], b[2]), max(b[1], b[3])
    area_a = f(f(ymax_a - ymin_a) * f(xmax_a - xmin_a))
    area_b = f(f(ymax_b - ymin_b) * f(xmax_b - xmin_b))
    if area_a <= 0 or area_b <= 0:
        return f(0.0)
    iy0, ix0 = max(ymin_a, ymin_b), max(xmin_a, xmin_b)
    iy1, ix1 = min(ymax_a, ymax_b), min(xmax_a, xmax_b)
    inter = f(max(f(iy1 - iy0), f(0.0)) * max(f(ix1 - ix0), f(0.0)))
    return f(inter / f(f(area_a + area_b) - inter))


def _tf_nms(boxes_yxyx, scores, max_out, iou_threshold):
    """tf.image.non_max_suppression as called at keras_layer_DecodeDetections.py
    :195-199: candidates by descending score (ties: lower index first), a
    candidate is dropped when its IoU with any already selected box is
    > iou_threshold (float32), stop after `max_out` selections."""
    order = np.lexsort((np.arange(scores.shape[0]), -scores.astype(np.float64)))
    thr = np.float32(iou_threshold)
    sel = []
    for i in order:
        if len(sel) >= max_out:
            break
        ok = True
        for j in reversed(sel):
            if _tf_iou(boxes_yxyx[i], boxes_yxyx[j]) > thr:
                ok = False
                break
        if ok:
            sel.append(int(i))
    return np.array(sel, dtype=np.int64)


def _layer_boxes(y_pred, normalize_coords, img_height, img_width, exp_mode):
    """keras_layer_DecodeDetections.py:124-146: float32 throughout; note the
    association `(off * var) * size + centre` differs from the numpy decoder's
    `off * (var * size) + centre`."""
    y = y_pred.astype(np.float32, copy=False)
    f = np.float32
    cx = y[..., -12] * y[..., -4] * y[..., -6] + y[..., -8]
    cy = y[..., -11] * y[..., -3] * y[..., -5] + y[..., -7]
    w = _exp(y[..., -10] * y[..., -2], exp_mode) * y[..., -6]
    h = _exp(y[..., -9] * y[..., -1], exp_mode) * y[..., -5]
    xmin = cx - f(0.5) * w
    ymin = cy - f(0.5) * h
    xmax = cx + f(0.5) * w
    ymax = cy + f(0.5) * h
    if normalize_coords:
        xmin = xmin * f(img_width)
        ymin = ymin * f(img_height)
        xmax = xmax * f(img_width)
        ymax = ymax * f(img_height)
    return xmin, ymin, xmax, ymax


def _layer_topk(rows, top_k):
    """keras_layer_DecodeDetections.py:238-251: pad to top_k with zero rows if
    short, then tf.nn.top_k(sorted=True) on the score column (stable: lower
    index first among equal scores)."""
    if rows.shape[0] < top_k:
        rows = np.concatenate([rows, np.zeros((top_k - rows.shape[0], 6), np.float32)], axis=0)
    order = np.lexsort((np.arange(rows.shape[0]), -rows[:, 1].astype(np.float64)))[:top_k]
    return rows[order]


def decode_layer(y_pred, confidence_thresh=0.01, iou_threshold=0.45, top_k=200,
                 nms_max_output_size=400, normalize_coords=True, img_height=None,
                 img_width=None, exp_mode='cr'):
    """keras_layer_DecodeDetections.py:109-265 restated on the CPU (float32).
    Output `(B, top_k, 6)` float32, zero rows = padding."""
    xmin, ymin, xmax, ymax = _layer_boxes(y_pred, normalize_coords, img_height, img_width, exp_mode)
    y = y_pred.astype(np.float32, copy=False)
    B, A = y.shape[0], y.shape[1]
    n_classes = y.shape[2] - 12
    thr = np.float32(confidence_thresh)
    out = np.zeros((B, top_k, 6), np.float32)
    for b in range(B):
        blocks = []
        for cid in range(1, n_classes):
            conf = y[b, :, cid]
            m = conf > thr
            blk = np.zeros((nms_max_output_size, 6), np.float32)
            if m.any():
                bx = np.stack([ymin[b][m], xmin[b][m], ymax[b][m], xmax[b][m]], axis=1)
                sel = _tf_nms(bx, conf[m], nms_max_output_size, iou_threshold)
                k = sel.shape[0]
                blk[:k, 0] = cid
                blk[:k, 1] = conf[m][sel]
                blk[:k, 2] = xmin[b][m][sel]
                blk[:k, 3] = ymin[b][m][sel]
                blk[:k, 4] = xmax[b][m][sel]
                blk[:k, 5] = ymax[b][m][sel]
            blocks.append(blk)
        out[b] = _layer_topk(np.concatenate(blocks, axis=0), top_k)
    return out


def decode_layer_fast(y_pred, confidence_thresh=0.01, iou_threshold=0.45, top_k=200,
                      nms_max_output_size=400, normalize_coords=True, img_height=None,
                      img_width=None, exp_mode='cr'):
    """keras_layer_DecodeDetectionsFast.py:111-248 restated on the CPU."""
    xmin, ymin, xmax, ymax = _layer_boxes(y_pred, normalize_coords, img_height, img_width, exp_mode)
    y = y_pred.astype(np.float32, copy=False)
    B = y.shape[0]
    cls = np.argmax(y[..., :-12], axis=-1)
    conf = np.amax(y[..., :-12], axis=-1)
    thr = np.float32(confidence_thresh)
    out = np.zeros((B, top_k, 6), np.float32)
    for b in range(B):
        m = (cls[b] != 0)
        if m.any():
            m = m & (conf[b] > thr)
        rows = np.zeros((1, 6), np.float32)
        if m.any():
            bx = np.stack([ymin[b][m], xmin[b][m], ymax[b][m], xmax[b][m]], axis=1)
            sel = _tf_nms(bx, conf[b][m], nms_max_output_size, iou_threshold)
            rows = np.stack([cls[b][m][sel].astype(np.float32), conf[b][m][sel], xmin[b][m][sel],
                             ymin[b][m][sel], xmax[b][m][sel], ymax[b][m][sel]], axis=1)
        out[b] = _layer_topk(rows, top_k)
    return out


# ----------------------------------------------------------------------------
# ssd_encoder_decoder/matching_utils.py
# ----------------------------------------------------------------------------

def match_bipartite_greedy(weight_matrix):
    """matching_utils.py:22-79.  Exactly `m` rounds on a private copy; rows that
    were already matched still take part in later rounds (all-zero row => its
    argmax is column 0 with weight 0)."""
    w = np.copy(weight_matrix)
    m = w.shape[0]
    rows = list(range(m))
    matches = np.zeros(m, dtype=int)
    for _ in range(m):
        col_of_row = np.argmax(w, axis=1)
        best = w[rows, col_of_row]
        g = np.argmax(best)
        a = col_of_row[g]
        matches[g] = a
        w[g] = 0
        w[:, a] = 0
    return matches


def match_multi(weight_matrix, threshold):
    """matching_utils.py:81-116."""
    n = weight_matrix.shape[1]
    cols = list(range(n))
    gt_of_col = np.argmax(weight_matrix, axis=0)
    best = weight_matrix[gt_of_col, cols]
    hit = np.nonzero(best >= threshold)[0]
    return gt_of_col[hit], hit


# ----------------------------------------------------------------------------
# ssd_encoder_decoder/ssd_input_encoder.py
# ----------------------------------------------------------------------------

class DegenerateBoxError(Exception):
    """ssd_input_encoder.py:613-617."""
    pass


def anchor_boxes_for_layer(img_height, img_width, feature_map_size, aspect_ratios, this_scale,
                           next_scale, two_boxes_for_ar1=True, this_steps=None, this_offsets=None,
                           clip_boxes=False, normalize_coords=True, coords='centroids'):
    """ssd_input_encoder.py:420-548.  Returns (boxes (fh,fw,nb,4), (cy,cx),
    wh_list, (step_h, step_w), (off_h, off_w))."""
    size = min(img_height, img_width)
    wh = []
    for ar in aspect_ratios:
        if ar == 1:
            side = this_scale * size
            wh.append((side, side))
            if two_boxes_for_ar1:
                side = np.sqrt(this_scale * next_scale) * size
                wh.append((side, side))
        else:
            wh.append((this_scale * size * np.sqrt(ar), this_scale * size / np.sqrt(ar)))
    wh = np.array(wh)
    nb = len(wh)
    fh, fw = feature_map_size[0], feature_map_size[1]

    if this_steps is None:
        step_h = img_height / fh
        step_w = img_width / fw
    elif isinstance(this_steps, (list, tuple)) and len(this_steps) == 2:
        step_h, step_w = this_steps[0], this_steps[1]
    elif isinstance(this_steps, (int, float)):
        step_h = step_w = this_steps
    if this_offsets is None:
        off_h = off_w = 0.5
    elif isinstance(this_offsets, (list, tuple)) and len(this_offsets) == 2:
        off_h, off_w = this_offsets[0], this_offsets[1]
    elif isinstance(this_offsets, (int, float)):
        off_h = off_w = this_offsets

    cy = np.linspace(off_h * step_h, (off_h + fh - 1) * step_h, fh)
    cx = np.linspace(off_w * step_w, (off_w + fw - 1) * step_w, fw)
    gx, gy = np.meshgrid(cx, cy)

    t = np.zeros((fh, fw, nb, 4))
    t[:, :, :, 0] = gx[:, :, None]
    t[:, :, :, 1] = gy[:, :, None]
    t[:, :, :, 2] = wh[:, 0]
    t[:, :, :, 3] = wh[:, 1]
    t = convert_coordinates(t, 0, 'centroids2corners')
    if clip_boxes:
        xs = t[:, :, :, [0, 2]]
        xs[xs >= img_width] = img_width - 1
        xs[xs < 0] = 0
        t[:, :, :, [0, 2]] = xs
        ys = t[:, :, :, [1, 3]]
        ys[ys >= img_height] = img_height - 1
        ys[ys < 0] = 0
        t[:, :, :, [1, 3]] = ys
    if normalize_coords:
        t[:, :, :, [0, 2]] /= img_width
        t[:, :, :, [1, 3]] /= img_height
    if coords == 'centroids':
        t = convert_coordinates(t, 0, 'corners2centroids', border_pixels='half')
    elif coords == 'minmax':
        t = convert_coordinates(t, 0, 'corners2minmax', border_pixels='half')
    return t, (cy, cx), wh, (step_h, step_w), (off_h, off_w)


class SSDInputEncoder(object):
    """ssd_input_encoder.py:27-611 (argument checks :142-180 omitted: they are
    host-side validation reproduced and tested in the product shim)."""

    def __init__(self, img_height, img_width, n_classes, predictor_sizes, min_scale=0.1,
                 max_scale=0.9, scales=None, aspect_ratios_global=[0.5, 1.0, 2.0],
                 aspect_ratios_per_layer=None, two_boxes_for_ar1=True, steps=None, offsets=None,
                 clip_boxes=False, variances=[0.1, 0.1, 0.2, 0.2], matching_type='multi',
                 pos_iou_threshold=0.5, neg_iou_limit=0.3, border_pixels='half',
                 coords='centroids', normalize_coords=True, background_id=0, log_wh=True):
        predictor_sizes = np.array(predictor_sizes)
        if predictor_sizes.ndim == 1:
            predictor_sizes = predictor_sizes[None, :]
        L = predictor_sizes.shape[0]
        self.img_height, self.img_width = img_height, img_width
        self.n_classes = n_classes + 1
        self.predictor_sizes = predictor_sizes
        self.scales = np.linspace(min_scale, max_scale, L + 1) if scales is None else np.array(scales)
        self.aspect_ratios = [aspect_ratios_global] * L if aspect_ratios_per_layer is None else aspect_ratios_per_layer
        self.two_boxes_for_ar1 = two_boxes_for_ar1
        self.steps = [None] * L if steps is None else steps
        self.offsets = [None] * L if offsets is None else offsets
        self.clip_boxes = clip_boxes
        self.variances = np.array(variances)
        self.matching_type = matching_type
        self.pos_iou_threshold = pos_iou_threshold
        self.neg_iou_limit = neg_iou_limit
        self.border_pixels = border_pixels
        self.coords = coords
        self.normalize_coords = normalize_coords
        self.background_id = background_id
        self.log_wh = log_wh
        self.boxes_list = []
        for i in range(L):
            boxes = anchor_boxes_for_layer(img_height, img_width, predictor_sizes[i], self.aspect_ratios[i],
                                           self.scales[i], self.scales[i + 1], two_boxes_for_ar1,
                                           self.steps[i], self.offsets[i], clip_boxes, normalize_coords, coords)[0]
            self.boxes_list.append(boxes)

    def anchors(self):
        """(A, 4) float64 in the order layer, y, x, box (:576-591)."""
        return np.concatenate([b.reshape(-1, 4) for b in self.boxes_list], axis=0)

    def generate_encoding_template(self, batch_size):
        """ssd_input_encoder.py:550-611."""
        anc = self.anchors()
        A = anc.shape[0]
        boxes = np.tile(anc[None], (batch_size, 1, 1))
        classes = np.zeros((batch_size, A, self.n_classes))
        var = np.zeros_like(boxes)
        var += self.variances
        return np.concatenate((classes, boxes, boxes, var), axis=2)

    def __call__(self, ground_truth_labels, diagnostics=False, return_matches=False):
        """ssd_input_encoder.py:277-418.  `return_matches` additionally returns,
        per image, an int array (A,) with the matched ground-truth index, -1 for
        background and -2 for neutral (the parity observable)."""
        B = len(ground_truth_labels)
        y = self.generate_encoding_template(B)
        y[:, :, self.background_id] = 1
        eye = np.eye(self.n_classes)
        A = y.shape[1]
        match_idx = np.full((B, A), -1, dtype=np.int32)
        for i in range(B):
            if ground_truth_labels[i].size == 0:
                continue
            lab = ground_truth_labels[i].astype(float)
            if np.any(lab[:, [3]] - lab[:, [1]] <= 0) or np.any(lab[:, [4]] - lab[:, [2]] <= 0):
                raise DegenerateBoxError("SSDInputEncoder detected degenerate ground truth bounding boxes for batch item {} with bounding boxes {}, ".format(i, lab) +
                                         "i.e. bounding boxes where xmax <= xmin and/or ymax <= ymin. Degenerate ground truth " +
                                         "bounding boxes will lead to NaN errors during the training.")
            if self.normalize_coords:
                lab[:, [2, 4]] /= self.img_height
                lab[:, [1, 3]] /= self.img_width
            if self.coords == 'centroids':
                lab = convert_coordinates(lab, 1, 'corners2centroids', border_pixels=self.border_pixels)
            elif self.coords == 'minmax':
                lab = convert_coordinates(lab, 1, 'corners2minmax')
            onehot = eye[lab[:, 0].astype(int)]
            rows = np.concatenate([onehot, lab[:, 1:5]], axis=-1)

            sim = iou(lab[:, 1:5], y[i, :, -12:-8], coords=self.coords, mode='outer_product',
                      border_pixels=self.border_pixels)
            bip = match_bipartite_greedy(sim)
            y[i, bip, :-8] = rows
            for g in range(len(bip)):              # duplicate indices: last write wins
                match_idx[i, bip[g]] = g
            sim[:, bip] = 0
            if self.matching_type == 'multi':
                gts, ancs = match_multi(sim, self.pos_iou_threshold)
                y[i, ancs, :-8] = rows[gts]
                match_idx[i, ancs] = gts
                sim[:, ancs] = 0
            neutral = np.nonzero(np.amax(sim, axis=0) >= self.neg_iou_limit)[0]
            y[i, neutral, self.background_id] = 0
            bg_only = match_idx[i, neutral] == -1
            match_idx[i, neutral[bg_only]] = -2

        if self.coords == 'centroids':
            y[:, :, [-12, -11]] -= y[:, :, [-8, -7]]
            y[:, :, [-12, -11]] /= y[:, :, [-6, -5]] * y[:, :, [-4, -3]]
            y[:, :, [-10, -9]] /= y[:, :, [-6, -5]]
            if self.log_wh:
                y[:, :, [-10, -9]] = np.log(y[:, :, [-10, -9]]) / y[:, :, [-2, -1]]
            else:                                   # *_no_log twin (:400)
                y[:, :, [-10, -9]] = y[:, :, [-10, -9]] / y[:, :, [-2, -1]]
        elif self.coords == 'corners':
            y[:, :, -12:-8] -= y[:, :, -8:-4]
            y[:, :, [-12, -10]] /= (y[:, :, -6] - y[:, :, -8])[..., None]
            y[:, :, [-11, -9]] /= (y[:, :, -5] - y[:, :, -7])[..., None]
            y[:, :, -12:-8] /= y[:, :, -4:]
        elif self.coords == 'minmax':
            y[:, :, -12:-8] -= y[:, :, -8:-4]
            y[:, :, [-12, -11]] /= (y[:, :, -7] - y[:, :, -8])[..., None]
            y[:, :, [-10, -9]] /= (y[:, :, -5] - y[:, :, -6])[..., None]
            y[:, :, -12:-8] /= y[:, :, -4:]

        outs = [y]
        if diagnostics:
            ym = np.copy(y)
            ym[:, :, -12:-8] = 0
            outs.append(ym)
        if return_matches:
            outs.append(match_idx)
        return outs[0] if len(outs) == 1 else tuple(outs)
