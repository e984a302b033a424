"""Drop-in for `ssd_encoder_decoder/ssd_input_encoder.py` of the reference
(/root/reference/localisation_part/ssd_encoder_decoder/ssd_input_encoder.py).

`SSDInputEncoder` keeps the reference's constructor signature, argument checks,
public attributes and `__call__` contract (list of `(m_i, 5)` ground-truth arrays in,
`(B, #boxes, #classes + 12)` float64 array out).  Anchor generation is construction-time
host configuration (numpy, once); ground truth -> anchor matching and offset encoding run
in libssdcodec's sm_100a kernels (csrc/encode.cu).
"""
from __future__ import division

import numpy as np

try:
    from .. import _lib
except ImportError:
    import _lib

LOG_WH = True  # the *_no_log twin module flips this


class DegenerateBoxError(Exception):
    """Raised when a ground-truth box has non-positive width or height (reference :613-617)."""
    pass


def _corners_from_centroids(t):
    out = np.empty_like(t)
    out[..., 0] = t[..., 0] - t[..., 2] / 2.0
    out[..., 1] = t[..., 1] - t[..., 3] / 2.0
    out[..., 2] = t[..., 0] + t[..., 2] / 2.0
    out[..., 3] = t[..., 1] + t[..., 3] / 2.0
    return out


class SSDInputEncoder(object):
    """Turns ground-truth boxes into SSD training targets (reference :27-611)."""

    def __init__(self,
                 img_height,
                 img_width,
                 n_classes,
                 predictor_sizes,
                 min_scale=0.1,
                 max_scale=0.9,
                 scales=None,
                 aspect_ratios_global=[0.5, 1.0, 2.0],
                 aspect_ratios_per_layer=None,
                 two_boxes_for_ar1=True,
                 steps=None,
                 offsets=None,
                 clip_boxes=False,
                 variances=[0.1, 0.1, 0.2, 0.2],
                 matching_type='multi',
                 pos_iou_threshold=0.5,
                 neg_iou_limit=0.3,
                 border_pixels='half',
                 coords='centroids',
                 normalize_coords=True,
                 background_id=0):
        predictor_sizes = np.array(predictor_sizes)
        if predictor_sizes.ndim == 1:
            predictor_sizes = np.expand_dims(predictor_sizes, axis=0)
        n_layers = predictor_sizes.shape[0]

        # ---- argument checks (reference :142-180, same messages) ----
        if (min_scale is None or max_scale is None) and scales is None:
            raise ValueError("Either `min_scale` and `max_scale` or `scales` need to be specified.")
        if scales:
            if len(scales) != n_layers + 1:
                raise ValueError("It must be either scales is None or len(scales) == len(predictor_sizes)+1, but len(scales) == {} and len(predictor_sizes)+1 == {}".format(len(scales), n_layers + 1))
            scales = np.array(scales)
            if np.any(scales <= 0):
                raise ValueError("All values in `scales` must be greater than 0, but the passed list of scales is {}".format(scales))
        else:
            if not 0 < min_scale <= max_scale:
                raise ValueError("It must be 0 < min_scale <= max_scale, but it is min_scale = {} and max_scale = {}".format(min_scale, max_scale))
        if aspect_ratios_per_layer is not None:
            if len(aspect_ratios_per_layer) != n_layers:
                raise ValueError("It must be either aspect_ratios_per_layer is None or len(aspect_ratios_per_layer) == len(predictor_sizes), but len(aspect_ratios_per_layer) == {} and len(predictor_sizes) == {}".format(len(aspect_ratios_per_layer), n_layers))
            for ars in aspect_ratios_per_layer:
                if np.any(np.array(ars) <= 0):
                    raise ValueError("All aspect ratios must be greater than zero.")
        else:
            if aspect_ratios_global is None:
                raise ValueError("At least one of `aspect_ratios_global` and `aspect_ratios_per_layer` must not be `None`.")
            if np.any(np.array(aspect_ratios_global) <= 0):
                raise ValueError("All aspect ratios must be greater than zero.")
        if len(variances) != 4:
            raise ValueError("4 variance values must be pased, but {} values were received.".format(len(variances)))
        variances = np.array(variances)
        if np.any(variances <= 0):
            raise ValueError("All variances must be >0, but the variances given are {}".format(variances))
        if coords not in ('minmax', 'centroids', 'corners'):
            raise ValueError("Unexpected value for `coords`. Supported values are 'minmax', 'corners' and 'centroids'.")
        if (steps is not None) and (len(steps) != n_layers):
            raise ValueError("You must provide at least one step value per predictor layer.")
        if (offsets is not None) and (len(offsets) != n_layers):
            raise ValueError("You must provide at least one offset value per predictor layer.")
        if border_pixels not in _lib.BORDER:
            raise ValueError("`border_pixels` must be one of 'half', 'include' and 'exclude'.")

        # ---- members (reference :186-236) ----
        self.img_height = img_height
        self.img_width = img_width
        self.n_classes = n_classes + 1
        self.predictor_sizes = predictor_sizes
        self.min_scale = min_scale
        self.max_scale = max_scale
        self.scales = np.linspace(min_scale, max_scale, n_layers + 1) if scales is None else scales
        self.aspect_ratios = [aspect_ratios_global] * n_layers if aspect_ratios_per_layer is None else aspect_ratios_per_layer
        self.two_boxes_for_ar1 = two_boxes_for_ar1
        self.steps = steps if steps is not None else [None] * n_layers
        self.offsets = offsets if offsets is not None else [None] * n_layers
        self.clip_boxes = clip_boxes
        self.variances = variances
        self.matching_type = matching_type
        self.pos_iou_threshold = pos_iou_threshold
        self.neg_iou_limit = neg_iou_limit
        self.border_pixels = border_pixels
        self.coords = coords
        self.normalize_coords = normalize_coords
        self.background_id = background_id
        if aspect_ratios_per_layer is not None:
            self.n_boxes = [len(ars) + 1 if ((1 in ars) & two_boxes_for_ar1) else len(ars)
                            for ars in aspect_ratios_per_layer]
        else:
            self.n_boxes = len(aspect_ratios_global) + 1 if ((1 in aspect_ratios_global) & two_boxes_for_ar1) else len(aspect_ratios_global)

        # ---- anchors, once (reference :238-275) ----
        self.boxes_list = []
        self.wh_list_diag = []
        self.steps_diag = []
        self.offsets_diag = []
        self.centers_diag = []
        for i in range(n_layers):
            boxes, center, wh, step, offset = self.generate_anchor_boxes_for_layer(
                feature_map_size=self.predictor_sizes[i], aspect_ratios=self.aspect_ratios[i],
                this_scale=self.scales[i], next_scale=self.scales[i + 1], this_steps=self.steps[i],
                this_offsets=self.offsets[i], diagnostics=True)
            self.boxes_list.append(boxes)
            self.wh_list_diag.append(wh)
            self.steps_diag.append(step)
            self.offsets_diag.append(offset)
            self.centers_diag.append(center)

        self._log_wh = LOG_WH
        self._handle = None      # ssdc_encoder*, created on first use (anchors are uploaded once)
        self._ctx = None

    # ------------------------------------------------------------------
    def _anchors(self):
        """All anchors `(A, 4)` float64 in layer / y / x / box order (reference :576-591)."""
        return np.ascontiguousarray(np.concatenate([b.reshape(-1, 4) for b in self.boxes_list], axis=0))

    def _encoder(self):
        ctx = _lib.get_context()
        if self._handle is not None and self._ctx is ctx:
            return ctx, self._handle
        self._release()
        anchors = self._anchors()
        var = np.ascontiguousarray(self.variances, dtype=np.float64)
        p = _lib.EncodeParams()
        p.n_classes = int(self.n_classes)
        p.background_id = int(self.background_id)
        p.coords = _lib.COORDS[self.coords]
        p.border_pixels = _lib.BORDER[self.border_pixels]
        p.matching_multi = 1 if self.matching_type == 'multi' else 0
        p.normalize = 1 if self.normalize_coords else 0
        p.log_wh = 1 if self._log_wh else 0
        p.pos_iou_threshold = float(self.pos_iou_threshold)
        p.neg_iou_limit = float(self.neg_iou_limit)
        p.img_h = float(self.img_height)
        p.img_w = float(self.img_width)
        h = _lib.C.c_void_p()
        _lib.check(ctx.lib.ssdc_encoder_create(ctx.handle, _lib.ptr(anchors), anchors.shape[0], _lib.ptr(var),
                                              _lib.C.byref(p), _lib.C.byref(h)))
        self._handle, self._ctx = h, ctx
        return ctx, h

    def _release(self):
        if getattr(self, '_handle', None) is not None and getattr(self._ctx, 'handle', None):
            self._ctx.lib.ssdc_encoder_destroy(self._handle)
        self._handle = None
        self._ctx = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def __call__(self, ground_truth_labels, diagnostics=False, return_matches=False):
        """reference :277-418.  `ground_truth_labels`: list (length = batch size) of 2D arrays
        with rows `(class_id, xmin, ymin, xmax, ymax)` in absolute pixels.  Returns `y_encoded`
        `(B, #boxes, #classes + 12)` float64; with `diagnostics=True` also the copy whose four
        offset columns are zeroed.  `return_matches=True` (extension) additionally returns the
        `(B, #boxes)` int32 assignment: matched ground-truth row, -1 background, -2 neutral."""
        B = len(ground_truth_labels)
        offs = np.zeros(B + 1, dtype=np.int64)
        parts = []
        for i in range(B):
            lab = np.asarray(ground_truth_labels[i])
            if lab.size == 0:       # nothing to match for this batch item (:329)
                offs[i + 1] = offs[i]
                continue
            lab = lab.astype(float)
            if lab.ndim != 2 or lab.shape[1] < 5:
                raise ValueError("ground truth labels must have shape (n_boxes, 5), got {} for batch item {}".format(lab.shape, i))
            parts.append(lab[:, :5])
            offs[i + 1] = offs[i] + lab.shape[0]
        gt = np.ascontiguousarray(np.concatenate(parts, axis=0)) if parts else np.zeros((0, 5))

        ctx, h = self._encoder()
        A = sum(int(b.shape[0] * b.shape[1] * b.shape[2]) for b in self.boxes_list)
        W = self.n_classes + 12
        # (large results come from the recycled pinned pool: no first-touch page faults, direct DMA; see _lib._PinnedPool)
        y = _lib.pinned_pool.empty((B, A, W), np.float64)
        y2 = _lib.pinned_pool.empty((B, A, W), np.float64) if diagnostics else None
        mi = np.empty((B, A), dtype=np.int32) if return_matches else None
        rc = ctx.lib.ssdc_encode(h, _lib.ptr(gt), _lib.ptr(offs), B, 0, _lib.ptr(y), _lib.ptr(y2), _lib.ptr(mi))
        if rc == _lib.ERR_DEGENERATE:
            i = int(ctx.lib.ssdc_encoder_bad_image(h))
            lab = np.asarray(ground_truth_labels[i]).astype(float)
            raise DegenerateBoxError("SSDInputEncoder detected degenerate ground truth bounding boxes for batch item {} with bounding boxes {}, ".format(i, lab) +
                                     "i.e. bounding boxes where xmax <= xmin and/or ymax <= ymin. Degenerate ground truth " +
                                     "bounding boxes will lead to NaN errors during the training.")
        if rc == _lib.ERR_ARG and ctx.lib.ssdc_encoder_bad_image(h) >= 0:
            raise IndexError(_lib.last_error())
        _lib.check(rc)
        outs = [y]
        if diagnostics:
            outs.append(y2)
        if return_matches:
            outs.append(mi)
        return outs[0] if len(outs) == 1 else tuple(outs)

    # ------------------------------------------------------------------
    def generate_anchor_boxes_for_layer(self,
                                        feature_map_size,
                                        aspect_ratios,
                                        this_scale,
                                        next_scale,
                                        this_steps=None,
                                        this_offsets=None,
                                        diagnostics=False):
        """reference :420-548.  `(fh, fw, n_boxes, 4)` float64 anchors of one predictor layer in
        the encoder's coordinate format (construction-time host configuration)."""
        short_side = min(self.img_height, self.img_width)
        sizes = []
        for ar in aspect_ratios:
            if ar == 1:
                s = this_scale * short_side
                sizes.append((s, s))
                if self.two_boxes_for_ar1:
                    s = np.sqrt(this_scale * next_scale) * short_side
                    sizes.append((s, s))
            else:
                sizes.append((this_scale * short_side * np.sqrt(ar), this_scale * short_side / np.sqrt(ar)))
        wh_list = np.array(sizes)
        n_boxes = len(wh_list)
        fh, fw = int(feature_map_size[0]), int(feature_map_size[1])

        if this_steps is None:
            step_height = self.img_height / feature_map_size[0]
            step_width = self.img_width / feature_map_size[1]
        elif isinstance(this_steps, (list, tuple)) and (len(this_steps) == 2):
            step_height, step_width = this_steps[0], this_steps[1]
        elif isinstance(this_steps, (int, float)):
            step_height = step_width = this_steps
        if this_offsets is None:
            offset_height = offset_width = 0.5
        elif isinstance(this_offsets, (list, tuple)) and (len(this_offsets) == 2):
            offset_height, offset_width = this_offsets[0], this_offsets[1]
        elif isinstance(this_offsets, (int, float)):
            offset_height = offset_width = this_offsets

        cy = np.linspace(offset_height * step_height, (offset_height + feature_map_size[0] - 1) * step_height, feature_map_size[0])
        cx = np.linspace(offset_width * step_width, (offset_width + feature_map_size[1] - 1) * step_width, feature_map_size[1])
        grid_x, grid_y = np.meshgrid(cx, cy)

        cent = np.zeros((fh, fw, n_boxes, 4))
        cent[..., 0] = grid_x[:, :, None]
        cent[..., 1] = grid_y[:, :, None]
        cent[..., 2] = wh_list[:, 0]
        cent[..., 3] = wh_list[:, 1]
        boxes = _corners_from_centroids(cent)
        if self.clip_boxes:
            xs = boxes[..., [0, 2]]
            xs[xs >= self.img_width] = self.img_width - 1
            xs[xs < 0] = 0
            boxes[..., [0, 2]] = xs
            ys = boxes[..., [1, 3]]
            ys[ys >= self.img_height] = self.img_height - 1
            ys[ys < 0] = 0
            boxes[..., [1, 3]] = ys
        if self.normalize_coords:
            boxes[..., [0, 2]] /= self.img_width
            boxes[..., [1, 3]] /= self.img_height
        if self.coords == 'centroids':
            out = np.empty_like(boxes)
            out[..., 0] = (boxes[..., 0] + boxes[..., 2]) / 2.0
            out[..., 1] = (boxes[..., 1] + boxes[..., 3]) / 2.0
            out[..., 2] = boxes[..., 2] - boxes[..., 0] + 0
            out[..., 3] = boxes[..., 3] - boxes[..., 1] + 0
            boxes = out
        elif self.coords == 'minmax':
            boxes = boxes[..., [0, 2, 1, 3]].copy()

        if diagnostics:
            return boxes, (cy, cx), wh_list, (step_height, step_width), (offset_height, offset_width)
        return boxes

    def generate_encoding_template(self, batch_size, diagnostics=False):
        """reference :550-611.  `(batch_size, #boxes, #classes + 12)` float64: zero class
        vector, the anchors twice, the variances."""
        ctx, h = self._encoder()
        A = sum(int(b.shape[0] * b.shape[1] * b.shape[2]) for b in self.boxes_list)
        out = np.empty((batch_size, A, self.n_classes + 12), dtype=np.float64)
        _lib.check(ctx.lib.ssdc_encoding_template(h, batch_size, _lib.ptr(out)))
        if diagnostics:
            return out, self.centers_diag, self.wh_list_diag, self.steps_diag, self.offsets_diag
        return out
