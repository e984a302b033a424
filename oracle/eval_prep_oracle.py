"""CPU restatement (numpy) of the decoder -> evaluator glue and of the augmentation box checks - TEST
INFRASTRUCTURE ONLY (imported by tests/, bench.py's CPU legs and oracle/make_golden.py, never by the product).

Follows, in /root/reference/localisation_part/:
    apply_inverse_transforms     data_generator/object_detection_2d_misc_utils.py:22-73
    Resize inverter              data_generator/object_detection_2d_geometric_ops.py:75-79
    patch sampler inverter       data_generator/object_detection_2d_patch_sampling_ops.py:316-320
    result records               eval_utils/average_precision_evaluator.py:402-422 (+ the 'f4' record, :668-675)
    BoxFilter.__call__           data_generator/object_detection_2d_image_boxes_validation_utils.py:174-232
Pinned: oracle/make_golden.py runs the reference's own functions / closures on the same inputs and requires
identical outputs (tests/golden/evalprep.npz).
"""
from __future__ import division

import numpy as np


def resize_inverter(img_height, img_width, out_height, out_width, cols=(2, 3, 4, 5)):
    """geometric_ops.py:75-79 with labels_format columns + 1 = `cols` (xmin, ymin, xmax, ymax)."""
    x0, y0, x1, y1 = cols

    def inverter(labels):
        labels = np.copy(labels)
        labels[:, [y0, y1]] = np.round(labels[:, [y0, y1]] * (img_height / out_height), decimals=0)
        labels[:, [x0, x1]] = np.round(labels[:, [x0, x1]] * (img_width / out_width), decimals=0)
        return labels
    return inverter


def translate_inverter(patch_ymin, patch_xmin, cols=(2, 3, 4, 5)):
    """patch_sampling_ops.py:316-320."""
    x0, y0, x1, y1 = cols

    def inverter(labels):
        labels = np.copy(labels)
        labels[:, [y0, y1]] += patch_ymin
        labels[:, [x0, x1]] += patch_xmin
        return labels
    return inverter


def apply_inverse_transforms(y_pred_decoded, inverse_transforms):
    """misc_utils.py:22-73 (list form)."""
    out = []
    for i in range(len(y_pred_decoded)):
        it = np.copy(y_pred_decoded[i])
        if it.size > 0:
            for inverter in inverse_transforms[i]:
                if inverter is not None:
                    it = inverter(it)
        out.append(it)
    return out


def evaluation_records(y_pred, round_confidences=False):
    """average_precision_evaluator.py:405-422 for one batch, flattened: (image index, class, confidence, box) in the
    loop's (image, row) order, narrowed to float32 like the structured array of :668-675."""
    img, cls, conf, box = [], [], [], []
    for k, item in enumerate(y_pred):
        for b in np.asarray(item).reshape(-1, 6):
            img.append(k)
            cls.append(int(b[0]))
            conf.append(round(b[1], round_confidences) if round_confidences else b[1])
            box.append([round(b[2], 1), round(b[3], 1), round(b[4], 1), round(b[5], 1)])
    return (np.array(img, np.int32), np.array(cls, np.int32), np.array(conf, np.float64).astype(np.float32),
            np.array(box, np.float64).reshape(-1, 4).astype(np.float32))


def box_filter(labels, image_height, image_width, check_overlap=True, check_min_area=True, check_degenerate=True,
               overlap_criterion='center_point', lower=0.3, upper=1.0, min_area=16, cols=(1, 2, 3, 4), border_pixels='half'):
    """image_boxes_validation_utils.py:174-232; returns the boolean mask `requirements_met`."""
    from oracle import ssd_codec_oracle as orc
    labels = np.copy(labels)
    xmin, ymin, xmax, ymax = cols
    ok = np.ones(labels.shape[0], dtype=bool)
    if check_degenerate:
        ok &= (labels[:, xmax] > labels[:, xmin]) & (labels[:, ymax] > labels[:, ymin])
    if check_min_area:
        ok &= (labels[:, xmax] - labels[:, xmin]) * (labels[:, ymax] - labels[:, ymin]) >= min_area
    if check_overlap:
        if overlap_criterion == 'iou':
            image_coords = np.array([0, 0, image_width, image_height])
            v = orc.iou(image_coords, labels[:, [xmin, ymin, xmax, ymax]], coords='corners', mode='element-wise', border_pixels=border_pixels)
            ok &= (v > lower) & (v <= upper)
        elif overlap_criterion == 'area':
            d = {'half': 0, 'include': 1, 'exclude': -1}[border_pixels]
            areas = (labels[:, xmax] - labels[:, xmin] + d) * (labels[:, ymax] - labels[:, ymin] + d)
            cl = np.copy(labels)
            cl[:, [ymin, ymax]] = np.clip(labels[:, [ymin, ymax]], a_min=0, a_max=image_height - 1)
            cl[:, [xmin, xmax]] = np.clip(labels[:, [xmin, xmax]], a_min=0, a_max=image_width - 1)
            inter = (cl[:, xmax] - cl[:, xmin] + d) * (cl[:, ymax] - cl[:, ymin] + d)
            lo = inter > lower * areas if lower == 0.0 else inter >= lower * areas
            ok &= lo & (inter <= upper * areas)
        else:
            cy = (labels[:, ymin] + labels[:, ymax]) / 2
            cx = (labels[:, xmin] + labels[:, xmax]) / 2
            ok &= (cy >= 0.0) & (cy <= image_height - 1) & (cx >= 0.0) & (cx <= image_width - 1)
    return ok
