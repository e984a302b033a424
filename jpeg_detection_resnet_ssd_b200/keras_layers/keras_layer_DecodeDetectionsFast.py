"""The `DecodeDetectionsFast` layer contract of the reference
(/root/reference/localisation_part/keras_layers/keras_layer_DecodeDetectionsFast.py:111-248)
as a numpy callable: argmax class per box, background dropped, `conf > float32(thresh)`, one
class-agnostic TensorFlow-style NMS (at most `nms_max_output_size` boxes), top-k, zero padding.
PARITY UNPINNED (see keras_layer_DecodeDetections.py)."""
from __future__ import division

try:
    from .. import _lib
except ImportError:
    import _lib
from .keras_layer_DecodeDetections import DecodeDetections as _Base


class DecodeDetectionsFast(_Base):
    def __init__(self, *args, **kwargs):
        super(DecodeDetectionsFast, self).__init__(*args, **kwargs)
        self._mode = _lib.MODE_LAYER_FAST
