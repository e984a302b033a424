"""Drop-in for `ssd_encoder_decoder/ssd_input_encoder_no_log.py`: the encoder whose width /
height targets are `(w_gt / w_anchor) / variance` without the logarithm (reference diff:
line 400).  One boolean in the encode kernels."""
from __future__ import division

from . import ssd_input_encoder as _base

DegenerateBoxError = _base.DegenerateBoxError


class SSDInputEncoder(_base.SSDInputEncoder):
    def __init__(self, *args, **kwargs):
        super(SSDInputEncoder, self).__init__(*args, **kwargs)
        self._log_wh = False
