"""Drop-in for `ssd_encoder_decoder/matching_utils.py` of the reference
(/root/reference/localisation_part/ssd_encoder_decoder/matching_utils.py).
Both matchers run on the device (csrc/thin.cu); inside `SSDInputEncoder.__call__`
the same algorithms are fused into the encode kernels (csrc/encode.cu)."""
from __future__ import division

import numpy as np

try:
    from .. import _lib
except ImportError:
    import _lib


def match_bipartite_greedy(weight_matrix):
    """reference :22-79.  `(m, n)` weights -> `(m,)` matched column per row."""
    w = np.ascontiguousarray(weight_matrix, dtype=np.float64)
    if w.ndim != 2:
        raise ValueError("weight_matrix must be 2D, got shape {}".format(w.shape))
    m, n = w.shape
    out = np.zeros(m, dtype=np.int64)
    if m == 0:
        return out.astype(int)
    ctx = _lib.get_context()
    _lib.check(ctx.lib.ssdc_match_bipartite_greedy(ctx.handle, _lib.ptr(w), m, n, _lib.ptr(out)))
    return out.astype(int)


def match_multi(weight_matrix, threshold):
    """reference :81-116.  Returns `(gt_indices, anchor_indices)` of every column whose
    best row weight is `>= threshold`, in ascending column order."""
    w = np.ascontiguousarray(weight_matrix, dtype=np.float64)
    if w.ndim != 2:
        raise ValueError("weight_matrix must be 2D, got shape {}".format(w.shape))
    m, n = w.shape
    if m == 0:
        raise ValueError("attempt to get argmax of an empty sequence")
    gts = np.empty(max(n, 1), dtype=np.int64)
    ancs = np.empty(max(n, 1), dtype=np.int64)
    k = _lib.C.c_int64(0)
    ctx = _lib.get_context()
    _lib.check(ctx.lib.ssdc_match_multi(ctx.handle, _lib.ptr(w), m, n, float(threshold),
                                       _lib.ptr(gts), _lib.ptr(ancs), _lib.C.byref(k)))
    return gts[:k.value].copy(), ancs[:k.value].copy()
