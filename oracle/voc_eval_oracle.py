"""CPU oracle of the VOC mAP evaluation core.  TEST INFRASTRUCTURE ONLY.

numpy restatement of `Evaluator.get_num_gt_per_class / match_predictions / compute_precision_recall /
compute_average_precisions / compute_mean_average_precision` of
/root/reference/localisation_part/eval_utils/average_precision_evaluator.py (lines cited per function).

Pinned: `oracle/make_golden.py` runs the real reference class (imported with stub modules for the
third-party packages that are absent here: bs4, h5py, keras ...) on seeded inputs and requires identical
outputs; vectors in `tests/golden/voc_*.npz`.

Data model (what the reference keeps in `self.prediction_results`, `data_generator.labels`, ...):
  prediction_results : list of length n_classes + 1; entry c is a list of tuples
                       (image_id, confidence, xmin, ymin, xmax, ymax); entry 0 is unused
  labels             : list (one entry per image) of arrays with rows [class_id, xmin, ymin, xmax, ymax]
  image_ids          : list of image ids (any hashable; compared as str like the reference does)
  eval_neutral       : None or list (per image) of bool arrays
"""
from __future__ import division

import numpy as np

from .ssd_codec_oracle import iou


def _filtered_labels(labels_i, ignore_under_area):
    # average_precision_evaluator.py:538-545 / :628-635
    if ignore_under_area > 0:
        return [l for l in labels_i if not ((l[4] - l[2]) * (l[3] - l[1]) < ignore_under_area)]
    return labels_i


def get_num_gt_per_class(labels, eval_neutral, n_classes, ignore_neutral_boxes=True, ignore_under_area=0):
    """:494-568"""
    counts = np.zeros(n_classes + 1, dtype=int)
    for i in range(len(labels)):
        boxes = np.asarray(_filtered_labels(labels[i], ignore_under_area))
        for j in range(boxes.shape[0]):
            if ignore_neutral_boxes and eval_neutral is not None:
                if not eval_neutral[i][j]:
                    counts[int(boxes[j, 0])] += 1
            else:
                counts[int(boxes[j, 0])] += 1
    return counts


def match_predictions(prediction_results, labels, image_ids, eval_neutral, n_classes,
                      ignore_neutral_boxes=True, matching_iou_threshold=0.5, border_pixels='include',
                      sorting_algorithm='quicksort', verbose=True, ignore_under_area=0):
    """:570-777.  Returns (true_positives, false_positives, cumulative_tp, cumulative_fp), each a list of
    length n_classes + 1 (entry 0 is `[]`).  `verbose=False` reproduces the reference's quirk: its loop then
    runs over `range(len(predictions.shape))`, i.e. only the single highest-confidence prediction of every
    class is evaluated (:692-696)."""
    neutral_avail = eval_neutral is not None
    ground_truth = {}
    for i in range(len(image_ids)):
        lab = _filtered_labels(labels[i], ignore_under_area)
        if ignore_neutral_boxes and neutral_avail:
            ground_truth[str(image_ids[i])] = (np.asarray(lab), np.asarray(eval_neutral[i]))
        else:
            ground_truth[str(image_ids[i])] = np.asarray(lab)

    tps, fps, ctps, cfps = [[]], [[]], [[]], [[]]
    for class_id in range(1, n_classes + 1):
        preds = prediction_results[class_id]
        tp = np.zeros(len(preds), dtype=int)
        fp = np.zeros(len(preds), dtype=int)
        if len(preds) == 0:
            tps.append(tp)
            fps.append(fp)
            continue                      # (the reference appends nothing to the cumulative lists here)
        nchar = len(str(preds[0][0])) + 6
        dt = np.dtype([('image_id', 'U{}'.format(nchar)), ('confidence', 'f4'), ('xmin', 'f4'),
                       ('ymin', 'f4'), ('xmax', 'f4'), ('ymax', 'f4')])
        arr = np.array(preds, dtype=dt)
        order = np.argsort(-arr['confidence'], kind=sorting_algorithm)
        srt = arr[order]
        matched = {}
        todo = range(len(preds)) if verbose else range(len(arr.shape))
        for i in todo:
            p = srt[i]
            image_id = p['image_id']
            box = np.asarray(list(p[['xmin', 'ymin', 'xmax', 'ymax']]))
            if ignore_neutral_boxes and neutral_avail:
                gt, neutral = ground_truth[image_id]
            else:
                gt = ground_truth[image_id]
            gt = np.asarray(gt)
            if gt.size == 0:
                # (np.asarray([]) has no second axis; the reference would raise here - an image without any
                # ground truth is treated as "no object of this class")
                fp[i] = 1
                continue
            mask = gt[:, 0] == class_id
            gt = gt[mask]
            if ignore_neutral_boxes and neutral_avail:
                neutral = neutral[mask]
            if gt.size == 0:
                fp[i] = 1
                continue
            with np.errstate(all='ignore'):
                ov = iou(gt[:, [1, 2, 3, 4]], box, coords='corners', mode='element-wise', border_pixels=border_pixels)
            j = np.argmax(ov)
            if ov[j] < matching_iou_threshold:
                fp[i] = 1
            else:
                if not (ignore_neutral_boxes and neutral_avail) or (neutral[j] == False):   # noqa: E712
                    if image_id not in matched:
                        tp[i] = 1
                        matched[image_id] = np.zeros(gt.shape[0], dtype=bool)
                        matched[image_id][j] = True
                    elif not matched[image_id][j]:
                        tp[i] = 1
                        matched[image_id][j] = True
                    else:
                        fp[i] = 1
        tps.append(tp)
        fps.append(fp)
        ctps.append(np.cumsum(tp))
        cfps.append(np.cumsum(fp))
    return tps, fps, ctps, cfps


def compute_precision_recall(cum_tp, cum_fp, num_gt_per_class, n_classes):
    """:779-822"""
    precisions, recalls = [[]], [[]]
    for c in range(1, n_classes + 1):
        tp, fp = cum_tp[c], cum_fp[c]
        with np.errstate(all='ignore'):
            precisions.append(np.where(tp + fp > 0, tp / (tp + fp), 0))
            recalls.append(tp / num_gt_per_class[c])
    return precisions, recalls


def compute_average_precisions(precisions, recalls, n_classes, mode='sample', num_recall_points=11):
    """:824-925"""
    if mode not in ('sample', 'integrate'):
        raise ValueError("`mode` can be either 'sample' or 'integrate', but received '{}'".format(mode))
    aps = [0.0]
    for c in range(1, n_classes + 1):
        prec, rec = precisions[c], recalls[c]
        ap = 0.0
        if mode == 'sample':
            for t in np.linspace(start=0, stop=1, num=num_recall_points, endpoint=True):
                sel = prec[rec >= t]
                ap += 0.0 if sel.size == 0 else np.amax(sel)
            ap /= num_recall_points
        else:
            ur, ui, _ = np.unique(rec, return_index=True, return_counts=True)
            maxp = np.zeros_like(ur)
            dr = np.zeros_like(ur)
            for i in range(len(ur) - 2, -1, -1):
                maxp[i] = np.maximum(np.amax(prec[ui[i]:ui[i + 1]]), maxp[i + 1])
                dr[i] = ur[i + 1] - ur[i]
            ap = np.sum(maxp * dr)
        aps.append(ap)
    return aps


def compute_mean_average_precision(average_precisions):
    """:927-947"""
    return np.average(average_precisions[1:])
