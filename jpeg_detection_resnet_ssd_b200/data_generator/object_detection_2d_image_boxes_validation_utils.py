"""Drop-in for `data_generator/object_detection_2d_image_boxes_validation_utils.py` of the reference
(/root/reference/localisation_part/data_generator/object_detection_2d_image_boxes_validation_utils.py):
`BoundGenerator` (:26-77), `BoxFilter` (:79-232) and `ImageValidator` (:234-322), the box checks of the augmentation
chain (patch sampling, resizing).  The checks themselves - degenerate boxes, minimum area, overlap of every box
with the image by centre point / IoU / area fraction - run in `ssdc_box_filter` (csrc/evalprep.cu, the codec's IoU
device function); `BoxFilter.filter_batch` checks the boxes of a whole batch of images in one launch.
"""
from __future__ import division

import numpy as np

try:
    from .. import _lib
except ImportError:
    import _lib

_CRITERIA = {'center_point': 0, 'iou': 1, 'area': 2}


class BoundGenerator:
    """reference :26-77: draws a (lower, upper) pair from a weighted sample space; `None` bounds mean 0.0 / 1.0."""

    def __init__(self,
                 sample_space=((0.1, None),
                               (0.3, None),
                               (0.5, None),
                               (0.7, None),
                               (0.9, None),
                               (None, None)),
                 weights=None):
        n = len(sample_space)
        if weights is not None and len(weights) != n:
            raise ValueError("`weights` must either be `None` for uniform distribution or have the same length as `sample_space`.")
        pairs = []
        for pair in sample_space:
            if len(pair) != 2:
                raise ValueError("All elements of the sample space must be 2-tuples.")
            lo = 0.0 if pair[0] is None else pair[0]
            hi = 1.0 if pair[1] is None else pair[1]
            if lo > hi:
                raise ValueError("For all sample space elements, the lower bound cannot be greater than the upper bound.")
            pairs.append([lo, hi])
        self.sample_space = pairs
        self.sample_space_size = n
        self.weights = [1.0 / n] * n if weights is None else weights

    def __call__(self):
        return self.sample_space[np.random.choice(self.sample_space_size, p=self.weights)]


class BoxFilter:
    """reference :79-232: returns all bounding boxes that are valid with respect to the defined criteria."""

    def __init__(self,
                 check_overlap=True,
                 check_min_area=True,
                 check_degenerate=True,
                 overlap_criterion='center_point',
                 overlap_bounds=(0.3, 1.0),
                 min_area=16,
                 labels_format={'class_id': 0, 'xmin': 1, 'ymin': 2, 'xmax': 3, 'ymax': 4},
                 border_pixels='half'):
        if not isinstance(overlap_bounds, (list, tuple, BoundGenerator)):
            raise ValueError("`overlap_bounds` must be either a 2-tuple of scalars or a `BoundGenerator` object.")
        if isinstance(overlap_bounds, (list, tuple)) and (overlap_bounds[0] > overlap_bounds[1]):
            raise ValueError("The lower bound must not be greater than the upper bound.")
        if not (overlap_criterion in {'iou', 'area', 'center_point'}):
            raise ValueError("`overlap_criterion` must be one of 'iou', 'area', or 'center_point'.")
        self.overlap_criterion = overlap_criterion
        self.overlap_bounds = overlap_bounds
        self.min_area = min_area
        self.check_overlap = check_overlap
        self.check_min_area = check_min_area
        self.check_degenerate = check_degenerate
        self.labels_format = labels_format
        self.border_pixels = border_pixels

    def _bounds(self):
        if isinstance(self.overlap_bounds, BoundGenerator):
            return self.overlap_bounds()
        return self.overlap_bounds

    def _mask(self, boxes, hw, lower, upper):
        n = boxes.shape[0]
        keep = np.empty(max(n, 1), dtype=np.uint8)
        if n:
            ctx = _lib.get_context()
            _lib.check(ctx.lib.ssdc_box_filter(ctx.handle, _lib.ptr(boxes), _lib.ptr(hw), n,
                                              1 if self.check_degenerate else 0, 1 if self.check_min_area else 0,
                                              1 if self.check_overlap else 0, _CRITERIA[self.overlap_criterion],
                                              float(lower), float(upper), float(self.min_area),
                                              _lib.BORDER[self.border_pixels], _lib.ptr(keep)))
        return keep[:n].astype(bool)

    def _coords(self, labels):
        f = self.labels_format
        return np.ascontiguousarray(np.asarray(labels)[:, [f['xmin'], f['ymin'], f['xmax'], f['ymax']]], dtype=np.float64)

    def __call__(self, labels, image_height=None, image_width=None):
        """reference :174-232."""
        labels = np.copy(labels)
        lower, upper = (0.0, 1.0)
        if self.check_overlap:
            lower, upper = self._bounds()
        if labels.shape[0] == 0:
            return labels
        h = 0.0 if image_height is None else float(image_height)
        w = 0.0 if image_width is None else float(image_width)
        hw = np.empty((labels.shape[0], 2), dtype=np.float64)
        hw[:, 0], hw[:, 1] = h, w
        return labels[self._mask(self._coords(labels), hw, lower, upper)]

    def filter_batch(self, labels_list, image_heights, image_widths):
        """The boxes of many images in one launch: `labels_list[i]` is checked against an image of
        `image_heights[i]` x `image_widths[i]` (scalars are broadcast).  One pair of overlap bounds is drawn for the
        whole batch.  Returns the list of filtered label arrays."""
        n = len(labels_list)
        hs = np.broadcast_to(np.asarray(image_heights, dtype=np.float64), (n,))
        ws = np.broadcast_to(np.asarray(image_widths, dtype=np.float64), (n,))
        lower, upper = self._bounds() if self.check_overlap else (0.0, 1.0)
        arrs = [np.asarray(l) for l in labels_list]
        counts = [a.shape[0] if a.size else 0 for a in arrs]
        if sum(counts) == 0:
            return [np.copy(a) for a in arrs]
        boxes = np.concatenate([self._coords(a) for a, c in zip(arrs, counts) if c], axis=0)
        hw = np.concatenate([np.tile([[hs[i], ws[i]]], (c, 1)) for i, c in enumerate(counts) if c], axis=0)
        keep = self._mask(np.ascontiguousarray(boxes), np.ascontiguousarray(hw), lower, upper)
        out, pos = [], 0
        for a, c in zip(arrs, counts):
            out.append(np.copy(a)[keep[pos:pos + c]] if c else np.copy(a))
            pos += c
        return out


class ImageValidator:
    """reference :234-322: True if a given minimum number of boxes meets the overlap requirements with the image."""

    def __init__(self,
                 overlap_criterion='center_point',
                 bounds=(0.3, 1.0),
                 n_boxes_min=1,
                 labels_format={'class_id': 0, 'xmin': 1, 'ymin': 2, 'xmax': 3, 'ymax': 4},
                 border_pixels='half'):
        if not ((isinstance(n_boxes_min, int) and n_boxes_min > 0) or n_boxes_min == 'all'):
            raise ValueError("`n_boxes_min` must be a positive integer or 'all'.")
        self.overlap_criterion = overlap_criterion
        self.bounds = bounds
        self.n_boxes_min = n_boxes_min
        self.labels_format = labels_format
        self.border_pixels = border_pixels
        self.box_filter = BoxFilter(check_overlap=True,
                                    check_min_area=False,
                                    check_degenerate=False,
                                    overlap_criterion=self.overlap_criterion,
                                    overlap_bounds=self.bounds,
                                    labels_format=self.labels_format,
                                    border_pixels=self.border_pixels)

    def __call__(self, labels, image_height, image_width):
        """reference :283-322."""
        bf = self.box_filter
        bf.overlap_bounds, bf.labels_format = self.bounds, self.labels_format
        n_valid = len(bf(labels=labels, image_height=image_height, image_width=image_width))
        if self.n_boxes_min == 'all':
            return n_valid == len(labels)
        return n_valid >= self.n_boxes_min
