"""The SSD multibox loss of the reference (`keras_loss_function/keras_ssd_loss.py`), forward pass, as a
numpy-in / numpy-out callable on the device.

In the reference `SSDLoss.compute_loss` builds a TensorFlow graph (it is handed to `model.compile`);
`DeviceSSDLoss` keeps the constructor and the method name and evaluates the same arithmetic (`ssdc_ssd_loss`,
csrc/loss.cu) for monitoring / validation: `compute_loss(y_true, y_pred) -> (batch_size,) float32`.  It is not a Keras
loss object (no gradients).  `SSDLoss` is the reference's own class whenever that can be imported (drop-in mode with
TensorFlow present), else `DeviceSSDLoss`.  `y_true` is what `SSDInputEncoder` returns (float64 or float32), `y_pred` the float32 model
output of the same shape.  Parity is unpinned (TensorFlow cannot be executed here): the device code and the
numpy oracle follow the TensorFlow graph op by op in float32.
"""
from __future__ import division

import numpy as np

try:
    from .. import _lib, _dropin
except ImportError:
    import _lib
    import _dropin


class DeviceSSDLoss:
    def __init__(self, neg_pos_ratio=3, n_neg_min=0, alpha=1.0):
        """reference :26-51."""
        self.neg_pos_ratio = neg_pos_ratio
        self.n_neg_min = n_neg_min
        self.alpha = alpha

    def compute_loss(self, y_true, y_pred):
        """reference :98-211.  Returns the per-image loss, shape `(batch_size,)`, float32."""
        y_true = np.asarray(y_true)
        y_pred = np.asarray(y_pred)
        if y_true.ndim != 3 or y_true.shape != y_pred.shape or y_true.shape[2] < 13:
            raise ValueError("`y_true` and `y_pred` must both have shape (batch_size, #boxes, #classes + 12), got {} and {}".format(y_true.shape, y_pred.shape))
        if y_true.dtype not in (np.float32, np.float64):
            y_true = y_true.astype(np.float64)
        y_true = np.ascontiguousarray(y_true)
        y_pred = np.ascontiguousarray(y_pred, dtype=np.float32)
        B, A, W = y_true.shape
        out = np.empty(B, dtype=np.float32)
        ctx = _lib.get_context()
        with ctx.call_lock:
            _lib.check(ctx.lib.ssdc_ssd_loss(ctx.handle, _lib.ptr(y_true), _lib.F32 if y_true.dtype == np.float32 else _lib.F64,
                                            _lib.ptr(y_pred), 0, B, A, W - 12, int(self.neg_pos_ratio), int(self.n_neg_min),
                                            float(self.alpha), _lib.ptr(out)))
        return out


# Drop-in mode: `from keras_loss_function.keras_ssd_loss import SSDLoss` in the reference's training scripts must keep
# returning the TensorFlow loss that `model.compile` needs.  If the reference's module is importable further down the
# package path, its class is re-exported unchanged (the device forward pass stays available as `DeviceSSDLoss`);
# otherwise `SSDLoss` is the device version.
_ref = _dropin.load_shadowed(__package__ or 'keras_loss_function', 'keras_ssd_loss') if (__package__ or '').split('.')[0] == 'keras_loss_function' else None
SSDLoss = _ref.SSDLoss if _ref is not None and hasattr(_ref, 'SSDLoss') else DeviceSSDLoss
