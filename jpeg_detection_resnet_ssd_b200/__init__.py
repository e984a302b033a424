"""B200-native SSD box codec: a drop-in for the `ssd_encoder_decoder`,
`bounding_box_utils` and `keras_layers.DecodeDetections*` modules of
Shulk97/JPEG_detection_Resnet_SSD (`localisation_part/`).

Host code is plain Python + numpy + ctypes; all arithmetic of the hot path runs
in hand-written sm_100a CUDA kernels inside `lib/libssdcodec.so`
(C ABI: `include/ssdcodec.h`).  No PyTorch, no CPU fallback.
"""
from ._lib import Context, SSDCodecError, get_context, set_context, pinned_empty  # noqa: F401

__all__ = ['Context', 'SSDCodecError', 'get_context', 'set_context', 'pinned_empty']
