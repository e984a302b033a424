"""The `DecodeDetections` layer contract of the reference
(/root/reference/localisation_part/keras_layers/keras_layer_DecodeDetections.py:109-265)
without Keras / TensorFlow: a callable `layer(y_pred) -> (batch, top_k, 6)` float32 array,
zero rows = padding, rows sorted by confidence.

The reference layer is float32 TensorFlow ops around `tf.image.non_max_suppression`
(a CPU op in the pinned TF 1.x); here the same pipeline runs in libssdcodec's kernels in
"layer mode": float32 decode with the layer's operation order, `conf > float32(thresh)`,
TensorFlow's float32 IoU, at most `nms_max_output_size` boxes per class, top-k by score with
the lower (class, rank) index first on ties.  PARITY UNPINNED: TensorFlow is not available
to execute the original, so this mode is checked against a CPU restatement and TensorFlow's
published NMS unit-test vectors only.

Drop-in mode: the reference's models build their inference graph with this class
(models/keras_ssd300_dct_j2d_resnet.py:42-43,885), so when the reference's own module is
importable further down the package path (Keras / TensorFlow present) `DecodeDetections`
IS the reference's Keras `Layer`, unchanged; the device callable is always available as
`DeviceDecodeDetections` (e.g. on the raw output of a `model_mode='training'` model).
"""
from __future__ import division

import numpy as np

try:
    from .. import _lib, _dropin
except ImportError:
    import _lib
    import _dropin

_MODE = _lib.MODE_LAYER


class DeviceDecodeDetections(object):
    def __init__(self,
                 confidence_thresh=0.01,
                 iou_threshold=0.45,
                 top_k=200,
                 nms_max_output_size=400,
                 coords='centroids',
                 normalize_coords=True,
                 img_height=None,
                 img_width=None,
                 **kwargs):
        if normalize_coords and ((img_height is None) or (img_width is None)):
            raise ValueError("If relative box coordinates are supposed to be converted to absolute coordinates, the decoder needs the image size in order to decode the predictions, but `img_height == {}` and `img_width == {}`".format(img_height, img_width))
        if coords != 'centroids':
            raise ValueError("The DetectionOutput layer currently only supports the 'centroids' coordinate format.")
        self.confidence_thresh = confidence_thresh
        self.iou_threshold = iou_threshold
        self.top_k = top_k
        self.normalize_coords = normalize_coords
        self.img_height = img_height
        self.img_width = img_width
        self.coords = coords
        self.nms_max_output_size = nms_max_output_size
        self.name = kwargs.get('name', 'decoded_predictions')
        self._mode = _MODE

    def call(self, y_pred, mask=None):
        y = np.ascontiguousarray(y_pred, dtype=np.float32)
        rows, counts, _ = _lib.run_decode(y, self._mode, self.confidence_thresh, self.iou_threshold, self.top_k,
                                          'centroids', self.normalize_coords, self.img_height, self.img_width,
                                          'half', log_wh=True, nms_cap=self.nms_max_output_size)
        B = y.shape[0]
        out = np.zeros((B, self.top_k, 6), dtype=np.float32)
        pos = 0
        for b in range(B):
            c = int(counts[b])
            out[b, :c] = rows[pos:pos + c]
            pos += c
        return out

    __call__ = call

    def compute_output_shape(self, input_shape):
        batch_size, n_boxes, last_axis = input_shape
        return (batch_size, self.top_k, 6)

    def get_config(self):
        return {
            'name': self.name,
            'confidence_thresh': self.confidence_thresh,
            'iou_threshold': self.iou_threshold,
            'top_k': self.top_k,
            'nms_max_output_size': self.nms_max_output_size,
            'coords': self.coords,
            'normalize_coords': self.normalize_coords,
            'img_height': self.img_height,
            'img_width': self.img_width,
        }


_ref = _dropin.load_shadowed(__package__ or 'keras_layers', 'keras_layer_DecodeDetections') if (__package__ or '').split('.')[0] == 'keras_layers' else None
DecodeDetections = _ref.DecodeDetections if _ref is not None and hasattr(_ref, 'DecodeDetections') else DeviceDecodeDetections
