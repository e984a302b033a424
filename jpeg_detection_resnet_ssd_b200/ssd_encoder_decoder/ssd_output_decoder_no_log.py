"""Drop-in for `ssd_encoder_decoder/ssd_output_decoder_no_log.py`: identical to
`ssd_output_decoder` except that box width / height are decoded without the exp
(reference diff: lines 175 and 297).  One boolean in the kernels, not separate code."""
from __future__ import division

import types as _types

from . import ssd_output_decoder as _base


def _rebind(fn):
    g = dict(fn.__globals__)
    g['LOG_WH'] = False
    return _types.FunctionType(fn.__code__, g, fn.__name__, fn.__defaults__, fn.__closure__)


greedy_nms = _base.greedy_nms
_greedy_nms = _base._greedy_nms
_greedy_nms2 = _base._greedy_nms2
_greedy_nms_debug = _base._greedy_nms_debug
decode_detections = _rebind(_base.decode_detections)
decode_detections_fast = _rebind(_base.decode_detections_fast)
decode_detections_debug = _rebind(_base.decode_detections_debug)
get_num_boxes_per_pred_layer = _base.get_num_boxes_per_pred_layer
get_pred_layers = _base.get_pred_layers
