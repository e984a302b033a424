"""The `DecodeDetectionsFast` layer contract of the reference
(/root/reference/localisation_part/keras_layers/keras_layer_DecodeDetectionsFast.py:111-248)
as a numpy callable: argmax class per box, background dropped, `conf > float32(thresh)`, one
class-agnostic TensorFlow-style NMS (at most `nms_max_output_size` boxes), top-k, zero padding.
PARITY UNPINNED (see keras_layer_DecodeDetections.py).  In drop-in mode with Keras importable
`DecodeDetectionsFast` is the reference's own `Layer`; the device callable is
`DeviceDecodeDetectionsFast`."""
from __future__ import division

try:
    from .. import _lib, _dropin
except ImportError:
    import _lib
    import _dropin
from .keras_layer_DecodeDetections import DeviceDecodeDetections as _Base


class DeviceDecodeDetectionsFast(_Base):
    def __init__(self, *args, **kwargs):
        super(DeviceDecodeDetectionsFast, self).__init__(*args, **kwargs)
        self._mode = _lib.MODE_LAYER_FAST


_ref = _dropin.load_shadowed(__package__ or 'keras_layers', 'keras_layer_DecodeDetectionsFast') if (__package__ or '').split('.')[0] == 'keras_layers' else None
DecodeDetectionsFast = _ref.DecodeDetectionsFast if _ref is not None and hasattr(_ref, 'DecodeDetectionsFast') else DeviceDecodeDetectionsFast
