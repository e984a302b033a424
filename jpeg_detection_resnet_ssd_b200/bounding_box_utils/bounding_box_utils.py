"""Drop-in for `bounding_box_utils/bounding_box_utils.py` of the reference
(/root/reference/localisation_part/bounding_box_utils/bounding_box_utils.py).
The arithmetic runs on the device (csrc/thin.cu) with the same device functions
the decode / encode kernels use."""
from __future__ import division

import numpy as np

try:
    from .. import _lib
except ImportError:
    import _lib

_CONV_MSG = ("Unexpected conversion value. Supported values are 'minmax2centroids', 'centroids2minmax', "
             "'corners2centroids', 'centroids2corners', 'minmax2corners', and 'corners2minmax'.")


def convert_coordinates(tensor, start_index, conversion, border_pixels='half'):
    """reference :24-87.  float64 copy of `tensor` with the four coordinates starting at
    `start_index` of the last axis converted; the conversion arithmetic is done in the
    input's dtype like the reference does."""
    if conversion not in _lib.CONVERSIONS:
        raise ValueError(_CONV_MSG)
    if border_pixels not in _lib.BORDER:
        raise ValueError("`border_pixels` must be one of 'half', 'include' and 'exclude'.")
    t = np.asarray(tensor)
    if t.dtype == np.float32:
        dt = _lib.F32
    else:
        dt = _lib.F64
        t = t.astype(np.float64, copy=False)
    t = np.ascontiguousarray(t)
    width = t.shape[-1]
    start = start_index if start_index >= 0 else width + start_index
    rows = t.size // width if width else 0
    out = np.empty(t.shape, dtype=np.float64)
    if rows:
        ctx = _lib.get_context()
        _lib.check(ctx.lib.ssdc_convert_coordinates(ctx.handle, _lib.ptr(t), dt, rows, width, start,
                                                   _lib.CONVERSIONS[conversion], _lib.BORDER[border_pixels],
                                                   _lib.ptr(out)))
    return out


def convert_coordinates2(tensor, start_index, conversion):
    """reference :89-117: the matrix-product formulation, 'minmax2centroids' and
    'centroids2minmax' only (same results as `convert_coordinates` up to rounding)."""
    if conversion not in ('minmax2centroids', 'centroids2minmax'):
        raise ValueError("Unexpected conversion value. Supported values are 'minmax2centroids' and 'centroids2minmax'.")
    return convert_coordinates(np.asarray(tensor, dtype=np.float64), start_index, conversion, border_pixels='half')


def _prepare(boxes1, boxes2, coords, mode):
    boxes1 = np.asarray(boxes1)
    boxes2 = np.asarray(boxes2)
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
        # dtype-aware conversion first (the reference converts in the input dtype, :335-336)
        boxes1 = convert_coordinates(boxes1, 0, 'centroids2corners')
        boxes2 = convert_coordinates(boxes2, 0, 'centroids2corners')
        coords = 'corners'
    elif coords not in ('minmax', 'corners'):
        raise ValueError("Unexpected value for `coords`. Supported values are 'minmax', 'corners' and 'centroids'.")
    b1 = np.ascontiguousarray(boxes1, dtype=np.float64)
    b2 = np.ascontiguousarray(boxes2, dtype=np.float64)
    return b1, b2, coords


def _pairwise(fn_name, b1, b2, coords, mode, border_pixels):
    if border_pixels not in _lib.BORDER:
        raise ValueError("`border_pixels` must be one of 'half', 'include' and 'exclude'.")
    m, n = b1.shape[0], b2.shape[0]
    if mode == 'outer_product':
        out = np.empty((m, n), dtype=np.float64)
        md = _lib.IOU_OUTER
    else:
        if not (m == n or m == 1 or n == 1):
            raise ValueError("operands could not be broadcast together with shapes ({},2) ({},2)".format(m, n))
        out = np.empty((max(m, n) if (m and n) else 0,), dtype=np.float64)
        md = _lib.IOU_ELEMENTWISE
    if out.size:
        ctx = _lib.get_context()
        fn = getattr(ctx.lib, fn_name)
        _lib.check(fn(ctx.handle, _lib.ptr(b1), m, _lib.ptr(b2), n, _lib.COORDS[coords], md,
                      _lib.BORDER[border_pixels], _lib.ptr(out)))
    return out


def iou(boxes1, boxes2, coords='centroids', mode='outer_product', border_pixels='half'):
    """reference :283-383.  `(m, n)` matrix ('outer_product') or a vector ('element-wise')
    of float64 IoU values; like the reference, `border_pixels` only enters the union."""
    b1, b2, coords = _prepare(boxes1, boxes2, coords, mode)
    return _pairwise('ssdc_iou', b1, b2, coords, mode, border_pixels)


def intersection_area(boxes1, boxes2, coords='centroids', mode='outer_product', border_pixels='half'):
    """reference :119-224.  Intersection areas; unlike inside `iou`, here `border_pixels`
    does enter the side lengths (+d)."""
    b1, b2, coords = _prepare(boxes1, boxes2, coords, mode)
    return _pairwise('ssdc_intersection_area', b1, b2, coords, mode, border_pixels)


def intersection_area_(boxes1, boxes2, coords='corners', mode='outer_product', border_pixels='half'):
    """reference :226-280: the same without the argument checks ('corners' / 'minmax' only)."""
    b1 = np.ascontiguousarray(boxes1, dtype=np.float64)
    b2 = np.ascontiguousarray(boxes2, dtype=np.float64)
    return _pairwise('ssdc_intersection_area', b1, b2, coords, mode, border_pixels)
