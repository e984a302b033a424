"""ctypes binding of libssdcodec.so (include/ssdcodec.h) and the process-wide context.

There is deliberately no CPU fallback: if the shared library is missing or no
B200 is visible, every codec entry point raises `SSDCodecError`.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SSDC_LIB_PATH') or os.path.join(_HERE, 'lib', 'libssdcodec.so')      # (override: A/B builds)

# ---- constants mirrored from include/ssdcodec.h ------------------------------
OK, ERR_CUDA, ERR_ARG, ERR_CAPACITY, ERR_DEGENERATE, ERR_NODEVICE, ERR_STATE = 0, -1, -2, -3, -4, -5, -6
F32, F64 = 0, 1
COORDS = {'centroids': 0, 'minmax': 1, 'corners': 2}
BORDER = {'half': 0, 'include': 1, 'exclude': 2}
MODE_PER_CLASS, MODE_FAST, MODE_LAYER, MODE_LAYER_FAST = 0, 1, 2, 3
CONVERSIONS = {'minmax2centroids': 0, 'centroids2minmax': 1, 'corners2centroids': 2,
               'centroids2corners': 3, 'minmax2corners': 4, 'corners2minmax': 5}
IOU_OUTER, IOU_ELEMENTWISE = 0, 1
OPTIONS = {'no_sweep': 0, 'floor_target': 1, 'enc_general': 2, 'enc_no_overlap': 3, 'enc_dense_patch': 4,
           'loss_no_tma': 5, 'h2d_chunk_mb': 6, 'no_pipeline': 7, 'enc_lanes': 8, 'd1_ctas': 9, 'no_l2_hints': 10, 'd1_warps': 11}
K_NAMES = ['decode_filter', 'plan', 'sort', 'nms', 'merge', 'enc_rowbest', 'enc_match', 'enc_write', 'thin', 'enc_patch']
K_COUNT = len(K_NAMES)


class SSDCodecError(RuntimeError):
    """Raised for every failure reported by libssdcodec (code in `.code`)."""

    def __init__(self, code, message):
        super().__init__('libssdcodec error %d: %s' % (code, message))
        self.code = code


class DecodeParams(C.Structure):
    _fields_ = [('mode', C.c_int32), ('input_coords', C.c_int32), ('normalize', C.c_int32),
                ('border_pixels', C.c_int32), ('top_k', C.c_int32), ('nms_cap', C.c_int32),
                ('log_wh', C.c_int32), ('do_nms', C.c_int32),
                ('conf_thresh', C.c_double), ('iou_thresh', C.c_double),
                ('img_h', C.c_double), ('img_w', C.c_double)]


class EncodeParams(C.Structure):
    _fields_ = [('n_classes', C.c_int32), ('background_id', C.c_int32), ('coords', C.c_int32),
                ('border_pixels', C.c_int32), ('matching_multi', C.c_int32), ('normalize', C.c_int32),
                ('log_wh', C.c_int32), ('reserved', C.c_int32),
                ('pos_iou_threshold', C.c_double), ('neg_iou_limit', C.c_double),
                ('img_h', C.c_double), ('img_w', C.c_double)]


_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
_pi64 = C.POINTER(C.c_int64)

# name -> (restype, argtypes); every symbol include/ssdcodec.h declares
SIGNATURES = {
    'ssdc_version': (_i, []),
    'ssdc_last_error': (C.c_char_p, []),
    'ssdc_device_count': (_i, []),
    'ssdc_init': (_i, [C.POINTER(C.c_int), _i, C.POINTER(_vp)]),
    'ssdc_destroy': (None, [_vp]),
    'ssdc_ctx_num_devices': (_i, [_vp]),
    'ssdc_synchronize': (_i, [_vp]),
    'ssdc_set_option': (_i, [_vp, _i, _i64]),
    'ssdc_get_option': (_i64, [_vp, _i]),
    'ssdc_launch_count': (_i64, [_vp]),
    'ssdc_profile_enable': (_i, [_vp, _i]),
    'ssdc_profile_read': (_i, [_vp, C.POINTER(C.c_double), _pi64]),
    'ssdc_timer_start': (_i, [_vp]),
    'ssdc_timer_stop': (_i, [_vp, C.POINTER(C.c_double)]),
    'ssdc_dev_alloc': (_i, [_vp, _i, C.c_uint64, C.POINTER(_vp)]),
    'ssdc_dev_free': (_i, [_vp, _i, _vp]),
    'ssdc_host_alloc': (_i, [C.c_uint64, C.POINTER(_vp)]),
    'ssdc_host_free': (_i, [_vp]),
    'ssdc_memcpy_h2d': (_i, [_vp, _i, _vp, _vp, C.c_uint64]),
    'ssdc_memcpy_d2h': (_i, [_vp, _i, _vp, _vp, C.c_uint64]),
    'ssdc_decode_submit': (_i, [_vp, _vp, _i, _i, _i64, _i64, _i, C.POINTER(DecodeParams)]),
    'ssdc_decode_collect': (_i, [_vp, _vp, _i64, _vp, _vp, _pi64]),
    'ssdc_decode_stats': (_i, [_vp, _pi64]),
    'ssdc_decode': (_i, [_vp, _vp, _i, _i64, _i64, _i, C.POINTER(DecodeParams), _vp, _i64, _vp, _vp, _pi64]),
    'ssdc_decode_results_dev': (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _pi64, _pi64, C.POINTER(C.c_int32)]),
    'ssdc_greedy_nms': (_i, [_vp, _vp, _vp, _i64, _d, _i, _i, _vp, _pi64]),
    'ssdc_encoder_create': (_i, [_vp, _vp, _i64, _vp, C.POINTER(EncodeParams), C.POINTER(_vp)]),
    'ssdc_encoder_destroy': (None, [_vp]),
    'ssdc_encoder_bad_image': (_i64, [_vp]),
    'ssdc_encode': (_i, [_vp, _vp, _vp, _i64, _i, _vp, _vp, _vp]),
    'ssdc_encoding_template': (_i, [_vp, _i64, _vp]),
    'ssdc_iou': (_i, [_vp, _vp, _i64, _vp, _i64, _i, _i, _i, _vp]),
    'ssdc_intersection_area': (_i, [_vp, _vp, _i64, _vp, _i64, _i, _i, _i, _vp]),
    'ssdc_convert_coordinates': (_i, [_vp, _vp, _i, _i64, _i, _i, _i, _i, _vp]),
    'ssdc_match_bipartite_greedy': (_i, [_vp, _vp, _i64, _i64, _vp]),
    'ssdc_match_multi': (_i, [_vp, _vp, _i64, _i64, _d, _vp, _vp, _pi64]),
    'ssdc_ssd_loss': (_i, [_vp, _vp, _i, _vp, _i, _i64, _i64, _i, _i, _i, _d, _vp]),
    'ssdc_inverse_transform_rows': (_i, [_vp, _vp, _i64, _i, _vp, _i64, _vp, _vp, _i, _i, _i, _i]),
    'ssdc_results_for_evaluation': (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i64, _pi64]),
    'ssdc_box_filter': (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _i, _d, _d, _d, _i, _vp]),
    'ssdc_voc_match': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i64, _d, _i, _i, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen the in-tree libssdcodec.so and declare every prototype."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise SSDCodecError(ERR_NODEVICE, 'shared library %s not found; build it with '
                                '`python -m jpeg_detection_resnet_ssd_b200.build` (there is no CPU fallback)' % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def last_error():
    msg = load_library().ssdc_last_error()
    return msg.decode('utf-8', 'replace') if msg else ''


def check(code):
    if code != OK:
        raise SSDCodecError(code, last_error())


def ptr(a):
    """Raw data pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    return C.c_void_p(a.ctypes.data)


class Context(object):
    """Owns one `ssdc_ctx` (streams, scratch, pinned staging) over a set of devices."""

    def __init__(self, devices=None):
        lib = load_library()
        if devices is None:
            devices = default_devices()
        self.devices = [int(d) for d in devices]
        arr = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        check(lib.ssdc_init(arr, len(self.devices), C.byref(h)))
        self.lib = lib
        self.handle = h
        self.call_lock = threading.RLock()

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.ssdc_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        """Test / diagnosis switches of the context (`OPTIONS`; include/ssdcodec.h SSDC_OPT_*); returns the old value."""
        old = int(self.lib.ssdc_get_option(self.handle, OPTIONS[name]))
        check(self.lib.ssdc_set_option(self.handle, OPTIONS[name], int(value)))
        return old

    def decode_stats(self):
        """(keys emitted by D1, images with an engaged score floor, images that took the exact fallback) of the last
        collected image-sweep decode."""
        out = (C.c_int64 * 3)()
        check(self.lib.ssdc_decode_stats(self.handle, out))
        return int(out[0]), int(out[1]), int(out[2])

    # -- measurement helpers (bench.py / tests) --
    def synchronize(self):
        check(self.lib.ssdc_synchronize(self.handle))

    def launch_count(self):
        return int(self.lib.ssdc_launch_count(self.handle))

    def profile_enable(self, on=True):
        check(self.lib.ssdc_profile_enable(self.handle, 1 if on else 0))

    def profile_read(self):
        ms = (C.c_double * K_COUNT)()
        n = (C.c_int64 * K_COUNT)()
        check(self.lib.ssdc_profile_read(self.handle, ms, n))
        return {K_NAMES[i]: (float(ms[i]), int(n[i])) for i in range(K_COUNT)}

    def timer_start(self):
        check(self.lib.ssdc_timer_start(self.handle))

    def timer_stop(self):
        ms = C.c_double()
        check(self.lib.ssdc_timer_stop(self.handle, C.byref(ms)))
        return float(ms.value)

    def dev_alloc(self, nbytes, slot=0):
        p = C.c_void_p()
        check(self.lib.ssdc_dev_alloc(self.handle, slot, nbytes, C.byref(p)))
        return p

    def dev_free(self, p, slot=0):
        check(self.lib.ssdc_dev_free(self.handle, slot, p))

    def h2d(self, dst, src_array, slot=0):
        check(self.lib.ssdc_memcpy_h2d(self.handle, slot, dst, ptr(src_array), src_array.nbytes))

    def d2h(self, dst_array, src, slot=0):
        check(self.lib.ssdc_memcpy_d2h(self.handle, slot, ptr(dst_array), src, dst_array.nbytes))


def pinned_empty(shape, dtype):
    """numpy array backed by page-locked host memory.

    The memory is released when the LAST array referring to it dies: the finalizer hangs on the ctypes buffer
    object at the end of every view's `.base` chain (`y[:n]`, `y.reshape(..)`, `np.asarray(y)` all keep it
    alive), not on the array object returned here."""
    import weakref
    lib = load_library()
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    n = count * dtype.itemsize
    p = C.c_void_p()
    check(lib.ssdc_host_alloc(max(n, 1), C.byref(p)))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    weakref.finalize(buf, _free_pinned, p.value)
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


def _free_pinned(address):
    if _lib is not None:
        _lib.ssdc_host_free(C.c_void_p(address))


class _PinnedPool(object):
    """Recycled page-locked buffers for the large arrays the codec RETURNS (`y_encoded`: 2.3 MB per SSD300 image).

    A fresh `np.empty` of that size costs a page fault per 4 KB on its first touch and makes the device-to-host copy a
    staged pageable one (measured: 16 ms per batch of 32 where the copy itself takes 1.4 ms from pinned memory).  The
    arrays handed out here are ordinary ndarrays over pinned memory; when the last view of one dies its buffer goes
    back to the pool instead of to the driver, so a training loop that drops batch i before batch i + 2 arrives never
    allocates again.  Bounded by SSDC_PINNED_POOL_GB (default 4) of outstanding + cached bytes; beyond that, or for
    small arrays, plain numpy memory is returned."""
    GRAIN = 1 << 21

    def __init__(self):
        self.lock = threading.Lock()
        self.free = {}           # size class -> [address]
        self.bytes = 0           # outstanding + cached
        self.limit = int(float(os.environ.get('SSDC_PINNED_POOL_GB', '4')) * 2 ** 30)

    def empty(self, shape, dtype, min_bytes=1 << 22):
        import weakref
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        n = count * dtype.itemsize
        if n < min_bytes:
            return np.empty(shape, dtype)
        cls = (n + self.GRAIN - 1) // self.GRAIN * self.GRAIN
        with self.lock:
            lst = self.free.get(cls)
            addr = lst.pop() if lst else None
            if addr is None:
                if self.bytes + cls > self.limit:
                    return np.empty(shape, dtype)
                self.bytes += cls
        if addr is None:
            p = C.c_void_p()
            if load_library().ssdc_host_alloc(cls, C.byref(p)) != OK:
                with self.lock:
                    self.bytes -= cls
                return np.empty(shape, dtype)
            addr = p.value
        buf = (C.c_char * n).from_address(addr)
        weakref.finalize(buf, self._release, addr, cls)
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def _release(self, addr, cls):
        with self.lock:
            self.free.setdefault(cls, []).append(addr)

    def trim(self):
        """Returns every cached buffer to the driver."""
        with self.lock:
            addrs = [(a, c) for c, lst in self.free.items() for a in lst]
            self.free = {}
            self.bytes -= sum(c for _, c in addrs)
        for a, _ in addrs:
            if _lib is not None:
                _lib.ssdc_host_free(C.c_void_p(a))


pinned_pool = _PinnedPool()


def default_devices():
    """SSDC_DEVICES="0,1,.." if set, else the torchrun LOCAL_RANK, else device 0."""
    env = os.environ.get('SSDC_DEVICES')
    if env:
        return [int(x) for x in env.split(',') if x.strip() != '']
    if os.environ.get('LOCAL_RANK') is not None:
        return [int(os.environ['LOCAL_RANK'])]
    return [0]


_default_ctx = None
_ctx_lock = threading.Lock()


def get_context():
    """Process-wide default context (created on first use)."""
    global _default_ctx
    with _ctx_lock:
        if _default_ctx is None:
            _default_ctx = Context()
        return _default_ctx


def set_context(ctx):
    """Install `ctx` as the process-wide default (returns the previous one)."""
    global _default_ctx
    with _ctx_lock:
        old, _default_ctx = _default_ctx, ctx
        return old


# ---- shared decode driver ------------------------------------------------------

def run_decode(y_pred, mode, confidence_thresh, iou_threshold, top_k, input_coords, normalize_coords,
               img_height, img_width, border_pixels, log_wh=True, nms_cap=0, do_nms=True, ctx=None):
    """Runs libssdcodec's decode pipeline on a host array.  Returns
    (rows (total, 6) float64, counts (B,) int32, anchor_idx (total,) int32)."""
    ctx = ctx or get_context()
    y = np.asarray(y_pred)
    if y.ndim != 3 or y.shape[2] < 14:
        raise ValueError('y_pred must have shape (batch, #boxes, #classes + 12), got {}'.format(y.shape))
    if y.dtype == np.float32:
        dt = F32
    else:
        dt = F64
        if y.dtype != np.float64:
            y = y.astype(np.float64)
    y = np.ascontiguousarray(y)
    B, A, W = y.shape
    p = DecodeParams()
    p.mode = mode
    p.input_coords = COORDS[input_coords]
    p.normalize = 1 if normalize_coords else 0
    p.border_pixels = BORDER[border_pixels]
    p.top_k = 0 if top_k == 'all' else int(top_k)
    p.nms_cap = int(nms_cap)
    p.log_wh = 1 if log_wh else 0
    p.do_nms = 1 if do_nms else 0
    p.conf_thresh = float(confidence_thresh)
    p.iou_thresh = float(iou_threshold) if do_nms else 0.0
    p.img_h = float(img_height) if normalize_coords else 1.0
    p.img_w = float(img_width) if normalize_coords else 1.0
    # Device scratch per image: the input copy, worst-case candidate lists and the decoded boxes.
    # Very large batches are processed in sub-batches so the scratch stays inside a fixed budget
    # (SSDC_SCRATCH_GB, default 32 of the 180 GB).
    elem = 4 if dt == F32 else 8
    n_seg = (W - 13) if mode in (MODE_PER_CLASS, MODE_LAYER) else 1
    per_image = A * W * elem + n_seg * A * (8 if dt == F32 else 16) + A * 4 * elem + 4 * A
    budget = float(os.environ.get('SSDC_SCRATCH_GB', '32')) * 2 ** 30
    n_dev = max(1, len(ctx.devices))
    max_images = max(n_dev, int(budget // max(per_image, 1)) * n_dev)
    if B > max_images:
        parts = [_decode_once(ctx, y[i:i + max_images], dt, p) for i in range(0, B, max_images)]
        return (np.concatenate([q[0] for q in parts], axis=0), np.concatenate([q[1] for q in parts]),
                np.concatenate([q[2] for q in parts]))
    return _decode_once(ctx, y, dt, p)


def _decode_once(ctx, y, dt, p):
    lib = ctx.lib
    B, A, W = y.shape
    counts = np.zeros(B, dtype=np.int32)
    total = C.c_int64(0)
    with ctx.call_lock:          # submit + collect form one transaction on the context
        check(lib.ssdc_decode_submit(ctx.handle, ptr(y), dt, 0, B, A, W - 12, C.byref(p)))
        cap = B * p.top_k if p.top_k > 0 else 0
        rows = np.empty((max(cap, 1), 6), dtype=np.float64)
        idx = np.empty(max(cap, 1), dtype=np.int32)
        rc = lib.ssdc_decode_collect(ctx.handle, ptr(rows), cap, ptr(counts), ptr(idx), C.byref(total))
        if rc == ERR_CAPACITY:
            cap = int(total.value)
            rows = np.empty((max(cap, 1), 6), dtype=np.float64)
            idx = np.empty(max(cap, 1), dtype=np.int32)
            rc = lib.ssdc_decode_collect(ctx.handle, ptr(rows), cap, ptr(counts), ptr(idx), C.byref(total))
        check(rc)
    n = int(total.value)
    return rows[:n], counts, idx[:n]
