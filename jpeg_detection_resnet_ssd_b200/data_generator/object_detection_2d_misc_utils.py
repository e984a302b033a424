"""Drop-in for `data_generator/object_detection_2d_misc_utils.py` of the reference
(/root/reference/localisation_part/data_generator/object_detection_2d_misc_utils.py:22-73).

`apply_inverse_transforms(y_pred_decoded, inverse_transforms)` maps decoded predictions back to the original
image.  The inverters the reference's own transformations hand out are affine maps of the four box columns:

    Resize.__call__        object_detection_2d_geometric_ops.py:75-79      x -> np.round(x * orig / resized, 0)
    RandomPatch.__call__   object_detection_2d_patch_sampling_ops.py:316-320   x -> x + patch offset
    (two more)             ...patch_sampling_ops.py:577, :730                  identity

They are recognised (by their closure, or given as the descriptor objects below) and the whole batch is transformed
by ONE kernel (`ssdc_inverse_transform_rows`, csrc/evalprep.cu).  Any other callable is arbitrary user code that no
library can run on the device: it is called exactly as the reference calls it.  Float64 predictions go to the device
(what `decode_detections` returns; the arithmetic is float64 like numpy's); for other dtypes (the float32 output of the
Keras layer) the reference's closures are called as they are, which keeps their float32 arithmetic.
"""
from __future__ import division

import numpy as np

try:
    from .. import _lib
except ImportError:
    import _lib

_DEFAULT_COLS = (2, 3, 4, 5)          # xmin, ymin, xmax, ymax of a prediction row [class, conf, xmin, ymin, xmax, ymax]


class _AffineInverter(object):
    """An inverter as data.  Calling it on a single `(k, n)` float64 label array runs the same device kernel as
    `apply_inverse_transforms` (there is no host arithmetic behind these objects)."""

    def __call__(self, labels):
        labels = np.asarray(labels)
        if labels.dtype != np.float64 or labels.ndim != 2:
            raise TypeError("inverter descriptors take (k, n) float64 label arrays (what decode_detections returns); "
                            "for other dtypes use the closure the reference's transformation returned")
        return apply_inverse_transforms([labels], [[self]])[0]


class ResizeInverter(_AffineInverter):
    """The inverter `Resize(height, width)(image, return_inverter=True)` returns, as data: boxes predicted on the
    `out_height` x `out_width` image back to the `img_height` x `img_width` original."""
    kind = 1

    def __init__(self, img_height, img_width, out_height, out_width, cols=_DEFAULT_COLS):
        self.a_y = img_height / out_height
        self.a_x = img_width / out_width
        self.cols = tuple(cols)


class TranslateInverter(_AffineInverter):
    """The inverter of the patch samplers: boxes inside a patch back to the image the patch was cut from."""
    kind = 2

    def __init__(self, patch_ymin, patch_xmin, cols=_DEFAULT_COLS):
        self.a_y = patch_ymin
        self.a_x = patch_xmin
        self.cols = tuple(cols)


def describe_inverter(fn):
    """(kind, a_y, a_x, cols) of an inverter that is one of the reference's affine closures or a descriptor object;
    None for anything else."""
    if fn is None:
        return (0, 0.0, 0.0, None)
    if isinstance(fn, (ResizeInverter, TranslateInverter)):
        return (fn.kind, float(fn.a_y), float(fn.a_x), fn.cols)
    code = getattr(fn, '__code__', None)
    cells = getattr(fn, '__closure__', None)
    if code is None:
        return None
    free = dict(zip(code.co_freevars, [c.cell_contents for c in cells])) if cells else {}
    try:
        if not free and code.co_argcount == 1 and code.co_names == () and len(code.co_code) <= 8:      # `return labels`
            return (0, 0.0, 0.0, None)
        if getattr(fn, '__name__', '') != 'inverter':
            return None
        if set(free) == {'img_height', 'img_width', 'self', 'xmax', 'xmin', 'ymax', 'ymin'}:          # Resize
            s = free['self']
            cols = (free['xmin'] + 1, free['ymin'] + 1, free['xmax'] + 1, free['ymax'] + 1)
            return (1, free['img_height'] / s.out_height, free['img_width'] / s.out_width, cols)
        if set(free) == {'patch_xmin', 'patch_ymin', 'xmax', 'xmin', 'ymax', 'ymin'}:                # patch samplers
            cols = (free['xmin'] + 1, free['ymin'] + 1, free['xmax'] + 1, free['ymax'] + 1)
            return (2, float(free['patch_ymin']), float(free['patch_xmin']), cols)
    except Exception:
        return None
    return None


def compile_inverse_transforms(inverse_transforms, n_images):
    """Per-image inverter lists -> (steps (S, 3) float64, step_offsets (B+1,) int64, cols), or None if any inverter
    is not one of the recognised affine maps (or they disagree about the columns)."""
    steps, offs, cols = [], np.zeros(n_images + 1, dtype=np.int64), None
    for i in range(n_images):
        for fn in (inverse_transforms[i] or []):
            d = describe_inverter(fn)
            if d is None:
                return None
            if d[0] == 0:
                continue
            if d[3] is not None:
                if cols is not None and tuple(d[3]) != tuple(cols):
                    return None
                cols = tuple(d[3])
            steps.append((float(d[0]), d[1], d[2]))
        offs[i + 1] = len(steps)
    arr = np.ascontiguousarray(np.array(steps, dtype=np.float64).reshape(-1, 3))
    return arr, offs, (cols or _DEFAULT_COLS)


def _host_path(items, inverse_transforms):
    out = []
    for i in range(len(items)):
        it = np.copy(items[i])
        if it.size > 0:
            for inverter in inverse_transforms[i]:
                if not (inverter is None):
                    it = inverter(it)
        out.append(it)
    return out


def apply_inverse_transforms(y_pred_decoded, inverse_transforms):
    """reference :22-73: list of `(k_i, 6)` arrays (or one `(B, k, 6)` array) -> the same structure, transformed."""
    is_list = isinstance(y_pred_decoded, list)
    if not is_list and not isinstance(y_pred_decoded, np.ndarray):
        raise ValueError("`y_pred_decoded` must be either a list or a Numpy array.")
    n = len(y_pred_decoded)
    items = [np.asarray(y_pred_decoded[i]) for i in range(n)]
    plan = compile_inverse_transforms(inverse_transforms, n)
    device_ok = plan is not None and all(it.size == 0 or (it.ndim == 2 and it.dtype == np.float64 and it.shape[1] == items[0].shape[-1])
                                         for it in items) and any(it.size for it in items)
    if not device_ok:
        out = _host_path(items, inverse_transforms)
        return out if is_list else (np.array(out) if n else np.copy(y_pred_decoded))
    steps, step_offs, cols = plan
    width = max(it.shape[1] for it in items if it.size)
    if max(cols) >= width:
        out = _host_path(items, inverse_transforms)
        return out if is_list else np.array(out)
    offs = np.zeros(n + 1, dtype=np.int64)
    for i, it in enumerate(items):
        offs[i + 1] = offs[i] + (it.shape[0] if it.size else 0)
    flat = np.ascontiguousarray(np.concatenate([it for it in items if it.size], axis=0))
    ctx = _lib.get_context()
    _lib.check(ctx.lib.ssdc_inverse_transform_rows(ctx.handle, _lib.ptr(flat), flat.shape[0], width, _lib.ptr(offs), n,
                                                  _lib.ptr(steps), _lib.ptr(step_offs), cols[0], cols[1], cols[2], cols[3]))
    out = [flat[offs[i]:offs[i + 1]] if items[i].size else np.copy(items[i]) for i in range(n)]
    if is_list:
        return out
    res = np.copy(y_pred_decoded)
    for i in range(n):
        if items[i].size:
            res[i] = out[i]
    return res
