"""Drop-in for `ssd_encoder_decoder/ssd_output_decoder.py` of the reference
(/root/reference/localisation_part/ssd_encoder_decoder/ssd_output_decoder.py).

Same function names, argument order, defaults, return types and exceptions; the
numpy bodies are replaced by calls into libssdcodec (sm_100a CUDA kernels:
anchor-offset decode, confidence filter + compaction, segmented sort, greedy NMS,
cross-class top-k).  numpy in, numpy out.
"""
from __future__ import division

import numpy as np

try:  # imported as part of the package ...
    from .. import _lib
except ImportError:  # ... or with the package directory itself on sys.path (reference layout)
    import _lib

_IMG_SIZE_MSG = ("If relative box coordinates are supposed to be converted to absolute coordinates, the decoder "
                 "needs the image size in order to decode the predictions, but `img_height == {}` and `img_width == {}`")
LOG_WH = True  # the *_no_log twin module flips this


def _split(rows, counts, empty):
    out, pos = [], 0
    for c in counts:
        c = int(c)
        out.append(rows[pos:pos + c] if c else empty())
        pos += c
    return out


def _check_border(border_pixels):
    if border_pixels not in _lib.BORDER:
        raise ValueError("`border_pixels` must be one of 'half', 'include' and 'exclude', but got '{}'.".format(border_pixels))


def _nms_indices(boxes, scores, iou_threshold, coords, border_pixels):
    ctx = _lib.get_context()
    n = boxes.shape[0]
    b = np.ascontiguousarray(boxes, dtype=np.float64)
    s = np.ascontiguousarray(scores, dtype=np.float64)
    keep = np.empty(max(n, 1), dtype=np.int32)
    k = _lib.C.c_int64(0)
    with ctx.call_lock:      # shares the decode scratch: never inside another thread's submit .. collect transaction
        _lib.check(ctx.lib.ssdc_greedy_nms(ctx.handle, _lib.ptr(b), _lib.ptr(s), n, float(iou_threshold),
                                          _lib.COORDS[coords], _lib.BORDER[border_pixels], _lib.ptr(keep), _lib.C.byref(k)))
    return keep[:k.value]


def greedy_nms(y_pred_decoded, iou_threshold=0.45, coords='corners', border_pixels='half'):
    """reference :27-75 - rows `[class_id, score, 4 coordinates]` per batch item."""
    _check_border(border_pixels)
    if coords not in _lib.COORDS:
        raise ValueError("Unexpected value for `coords`. Supported values are 'minmax', 'corners' and 'centroids'.")
    out = []
    for item in y_pred_decoded:
        item = np.asarray(item)
        if item.shape[0] == 0:
            out.append(np.array([]))
            continue
        keep = _nms_indices(item[:, 2:6], item[:, 1], iou_threshold, coords, border_pixels)
        out.append(np.array(item[keep]))
    return out


def _greedy_nms(predictions, iou_threshold=0.45, coords='corners', border_pixels='half'):
    """reference :77-92 - rows `[score, 4 coordinates]`."""
    predictions = np.asarray(predictions)
    if predictions.shape[0] == 0:
        return np.array([])
    keep = _nms_indices(predictions[:, 1:5], predictions[:, 0], iou_threshold, coords, border_pixels)
    return np.array(predictions[keep])


def _greedy_nms2(predictions, iou_threshold=0.45, coords='corners', border_pixels='half'):
    """reference :94-109 - rows `[class_id, score, 4 coordinates]`."""
    predictions = np.asarray(predictions)
    if predictions.shape[0] == 0:
        return np.array([])
    keep = _nms_indices(predictions[:, 2:6], predictions[:, 1], iou_threshold, coords, border_pixels)
    return np.array(predictions[keep])


def _greedy_nms_debug(predictions, iou_threshold=0.45, coords='corners', border_pixels='half'):
    """reference :469-486 - rows `[box_id, score, 4 coordinates]`."""
    predictions = np.asarray(predictions)
    if predictions.shape[0] == 0:
        return np.array([])
    keep = _nms_indices(predictions[:, 2:6], predictions[:, 1], iou_threshold, coords, border_pixels)
    return np.array(predictions[keep])


def decode_detections(y_pred,
                      confidence_thresh=0.01,
                      iou_threshold=0.45,
                      top_k=200,
                      input_coords='centroids',
                      normalize_coords=True,
                      img_height=None,
                      img_width=None,
                      border_pixels='half'):
    """reference :111-226.  Returns a list (one entry per batch item) of float64
    arrays `(k, 6)` with rows `[class_id, confidence, xmin, ymin, xmax, ymax]`;
    an item without predictions is `np.array([])` of shape `(0,)`."""
    if normalize_coords and ((img_height is None) or (img_width is None)):
        raise ValueError(_IMG_SIZE_MSG.format(img_height, img_width))
    if input_coords not in _lib.COORDS:
        raise ValueError("Unexpected value for `input_coords`. Supported input coordinate formats are 'minmax', 'corners' and 'centroids'.")
    _check_border(border_pixels)
    rows, counts, _ = _lib.run_decode(y_pred, _lib.MODE_PER_CLASS, confidence_thresh, iou_threshold, top_k,
                                      input_coords, normalize_coords, img_height, img_width, border_pixels,
                                      log_wh=LOG_WH)
    return _split(rows, counts, lambda: np.array([]))


def decode_detections_fast(y_pred,
                           confidence_thresh=0.5,
                           iou_threshold=0.45,
                           top_k='all',
                           input_coords='centroids',
                           normalize_coords=True,
                           img_height=None,
                           img_width=None,
                           border_pixels='half'):
    """reference :228-333 (argmax class, one class-agnostic NMS per image)."""
    if normalize_coords and ((img_height is None) or (img_width is None)):
        raise ValueError(_IMG_SIZE_MSG.format(img_height, img_width))
    if input_coords not in _lib.COORDS:
        raise ValueError("Unexpected value for `coords`. Supported values are 'minmax', 'corners' and 'centroids'.")
    _check_border(border_pixels)
    do_nms = bool(iou_threshold)           # `if iou_threshold:` (:326)
    rows, counts, _ = _lib.run_decode(y_pred, _lib.MODE_FAST, confidence_thresh, iou_threshold if do_nms else 0.0,
                                      top_k, input_coords, normalize_coords, img_height, img_width, border_pixels,
                                      log_wh=LOG_WH, do_nms=do_nms)
    y = np.asarray(y_pred)
    # the reference never leaves float32 when input_coords == 'corners' (no convert_coordinates call)
    out_dtype = np.float32 if (y.dtype == np.float32 and input_coords == 'corners') else np.float64
    if out_dtype != np.float64:
        rows = rows.astype(out_dtype)
    if do_nms:
        empty = lambda: np.array([])
    else:
        empty = lambda: np.zeros((0, 6), dtype=out_dtype)
    return _split(rows, counts, empty)


def decode_detections_debug(y_pred,
                            confidence_thresh=0.01,
                            iou_threshold=0.45,
                            top_k=200,
                            input_coords='centroids',
                            normalize_coords=True,
                            img_height=None,
                            img_width=None,
                            variance_encoded_in_target=False,
                            border_pixels='half'):
    """reference :342-467: like `decode_detections` with the anchor (box) index
    prepended: rows `[box_id, class_id, confidence, xmin, ymin, xmax, ymax]`."""
    if normalize_coords and ((img_height is None) or (img_width is None)):
        raise ValueError(_IMG_SIZE_MSG.format(img_height, img_width))
    if input_coords not in _lib.COORDS:
        raise ValueError("Unexpected value for `input_coords`. Supported input coordinate formats are 'minmax', 'corners' and 'centroids'.")
    _check_border(border_pixels)
    y = np.asarray(y_pred)
    if variance_encoded_in_target and input_coords == 'centroids':
        # :405-409: the variances are already folded into the targets => decode with unit variances
        y = np.array(y, copy=True)
        y[:, :, -4:] = 1
    rows, counts, idx = _lib.run_decode(y, _lib.MODE_PER_CLASS, confidence_thresh, iou_threshold, top_k,
                                        input_coords, normalize_coords, img_height, img_width, border_pixels,
                                        log_wh=LOG_WH)
    rows7 = np.concatenate([idx.astype(np.float64)[:, None], rows], axis=1)
    out, pos = [], 0
    for c in counts:
        c = int(c)
        if c == 0:   # the reference's np.concatenate([]) fails here (:461)
            raise ValueError("need at least one array to concatenate")
        out.append(rows7[pos:pos + c])
        pos += c
    return out


def get_num_boxes_per_pred_layer(predictor_sizes, aspect_ratios, two_boxes_for_ar1):
    """reference :488-501."""
    counts = []
    for size, ars in zip(predictor_sizes, aspect_ratios):
        per_cell = len(ars) + 1 if two_boxes_for_ar1 else len(ars)
        counts.append(size[0] * size[1] * per_cell)
    return counts


def get_pred_layers(y_pred_decoded, num_boxes_per_pred_layer):
    """reference :503-530: predictor layer index of every row of a
    `decode_detections_debug` result."""
    edges = np.cumsum(num_boxes_per_pred_layer)
    result = []
    for item in y_pred_decoded:
        layers = []
        for row in item:
            if (row[0] < 0) or (row[0] >= edges[-1]):
                raise ValueError("Box index is out of bounds of the possible indices as given by the values in `num_boxes_per_pred_layer`.")
            layers.append(int(np.searchsorted(edges, row[0], side='right')))
        result.append(layers)
    return result
