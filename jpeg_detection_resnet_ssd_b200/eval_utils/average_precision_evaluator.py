"""Drop-in for the evaluation core of `eval_utils/average_precision_evaluator.py` of the reference
(/root/reference/localisation_part/eval_utils/average_precision_evaluator.py): the `Evaluator` methods that
turn decoded detections into VOC average precisions.

`match_predictions` - in the reference a Python loop over every prediction (10^5 .. 10^6 for a VOC test set)
- runs on the device (`ssdc_voc_match`, csrc/voc.cu); the precision / recall / AP reductions over the
resulting sorted arrays are a few numpy vector operations exactly as in the reference.

Out of scope here: `predict_on_dataset` (needs the Keras model and the data generator, SURVEY section 2
#11/#12).  Feed the detections with `set_predictions(...)` or assign `evaluator.prediction_results`
directly (same structure the reference builds at :405-422: a list with one entry per class id, each a
list of `(image_id, confidence, xmin, ymin, xmax, ymax)` tuples).
"""
from __future__ import division

import numpy as np

try:
    from .. import _lib, _dropin
except ImportError:
    import _lib
    import _dropin


class DeviceEvaluator(object):

    def __init__(self,
                 model,
                 n_classes,
                 data_generator,
                 model_mode='inference',
                 pred_format={'class_id': 0, 'conf': 1, 'xmin': 2, 'ymin': 3, 'xmax': 4, 'ymax': 5},
                 gt_format={'class_id': 0, 'xmin': 1, 'ymin': 2, 'xmax': 3, 'ymax': 4},
                 ignore_under_area=0):
        """Same constructor as the reference (:50-95).  `data_generator` only needs the attributes the
        evaluation core reads: `labels`, `image_ids` and `eval_neutral` (may be None)."""
        self.model = model
        self.data_generator = data_generator
        self.n_classes = n_classes
        self.model_mode = model_mode
        self.pred_format = pred_format
        self.gt_format = gt_format
        self.ignore_under_area = ignore_under_area
        self.prediction_results = None
        self._acc = None                 # flat records of add_decoded_batch (alternative to prediction_results)
        self.num_gt_per_class = None
        self.true_positives = None
        self.false_positives = None
        self.cumulative_true_positives = None
        self.cumulative_false_positives = None
        self.cumulative_precisions = None
        self.cumulative_recalls = None
        self.average_precisions = None
        self.mean_average_precision = None

    # ------------------------------------------------------------------
    def set_predictions(self, prediction_results):
        """`prediction_results[class_id]` = list of `(image_id, confidence, xmin, ymin, xmax, ymax)`."""
        self.prediction_results = prediction_results
        self._acc = None

    def reset_predictions(self):
        self.prediction_results = None
        self._acc = None

    def add_decoded_batch(self, batch_image_ids, inverse_transforms=None, round_confidences=False):
        """The body of the reference's prediction loop after `model.predict` (:381-422) without a Python loop over the
        detections: takes the result of the LAST `decode_detections` call of this process (or `ssdc_decode_submit`
        on a device-resident batch) where it lies - in device memory -, maps the boxes back to the original images
        (`inverse_transforms`: per image a list of the inverters the reference's Resize / patch samplers return, or
        the descriptor objects of data_generator.object_detection_2d_misc_utils), rounds confidences / coordinates
        like :411-418 and appends flat records that `match_predictions` hands to the device matcher as they are.
        Needs the image-sweep decode (finite `top_k`, float32 predictions)."""
        try:
            from ..data_generator import object_detection_2d_misc_utils as mu
        except ImportError:
            from data_generator import object_detection_2d_misc_utils as mu
        n_img = len(batch_image_ids)
        steps = step_offs = None
        if inverse_transforms is not None:
            plan = mu.compile_inverse_transforms(inverse_transforms, n_img)
            if plan is None or tuple(plan[2]) != (self.pred_format['xmin'], self.pred_format['ymin'], self.pred_format['xmax'], self.pred_format['ymax']):
                raise ValueError("add_decoded_batch only takes the affine inverters of the reference's transformations; for arbitrary "
                                 "callables use decode_detections + apply_inverse_transforms + set_predictions.")
            steps, step_offs = plan[0], plan[1]
        ctx = _lib.get_context()
        cap = 1 << 12
        while True:
            img = np.empty(cap, np.int32); cls = np.empty(cap, np.int32)
            conf = np.empty(cap, np.float32); box = np.empty((cap, 4), np.float32)
            n = _lib.C.c_int64(0)
            rc = ctx.lib.ssdc_results_for_evaluation(ctx.handle, _lib.ptr(steps), _lib.ptr(step_offs),
                                                     int(round_confidences) if round_confidences else -1,
                                                     _lib.ptr(img), _lib.ptr(cls), _lib.ptr(conf), _lib.ptr(box), cap, _lib.C.byref(n))
            if rc == _lib.ERR_CAPACITY and cap < (1 << 28):
                cap *= 8
                continue
            _lib.check(rc)
            break
        n = int(n.value)
        if n and int(img[:n].max()) >= n_img:
            raise ValueError("the decoded batch holds more images than `batch_image_ids`")
        ids = np.asarray([str(x) for x in batch_image_ids], dtype=object)
        if getattr(self, '_acc', None) is None:
            self._acc = {'ids': [], 'cls': [], 'conf': [], 'box': []}
        self._acc['ids'].append(ids[img[:n]])
        self._acc['cls'].append(cls[:n].copy())
        self._acc['conf'].append(conf[:n].copy())
        self._acc['box'].append(box[:n].copy())
        return n

    def materialize_prediction_results(self):
        """`prediction_results` in the reference's list-of-tuples structure (:405-422) from the records collected by
        `add_decoded_batch` (for `write_predictions_to_txt` and other consumers of the original structure)."""
        if getattr(self, '_acc', None) is None:
            return self.prediction_results
        res = [list() for _ in range(self.n_classes + 1)]
        for ids, cls, conf, box in zip(self._acc['ids'], self._acc['cls'], self._acc['conf'], self._acc['box']):
            for i in range(len(cls)):
                res[int(cls[i])].append((ids[i], float(conf[i]), float(box[i, 0]), float(box[i, 1]), float(box[i, 2]), float(box[i, 3])))
        return res

    def predict_on_dataset(self, *args, **kwargs):
        raise NotImplementedError("predict_on_dataset needs the Keras model and the data generator, which are outside "
                                  "the box codec; decode the model output with ssd_output_decoder.decode_detections and "
                                  "pass the per-class results to set_predictions().")

    def __call__(self, mode='sample', num_recall_points=11, ignore_neutral_boxes=True, matching_iou_threshold=0.5,
                 border_pixels='include', sorting_algorithm='quicksort', return_precisions=False,
                 return_recalls=False, return_average_precisions=False, verbose=True, **unused):
        """The evaluation part of the reference's `__call__` (:97-259), starting from `prediction_results`."""
        if self.prediction_results is None and getattr(self, '_acc', None) is None:
            raise ValueError("There are no prediction results. Provide them with `set_predictions()` or `add_decoded_batch()`.")
        self.get_num_gt_per_class(ignore_neutral_boxes=ignore_neutral_boxes, verbose=False, ret=False)
        self.match_predictions(ignore_neutral_boxes=ignore_neutral_boxes, matching_iou_threshold=matching_iou_threshold,
                               border_pixels=border_pixels, sorting_algorithm=sorting_algorithm, verbose=verbose, ret=False)
        self.compute_precision_recall(verbose=False, ret=False)
        self.compute_average_precisions(mode=mode, num_recall_points=num_recall_points, verbose=False, ret=False)
        mean_average_precision = self.compute_mean_average_precision(ret=True)
        if return_average_precisions or return_precisions or return_recalls:
            ret = [mean_average_precision]
            if return_average_precisions:
                ret.append(self.average_precisions)
            if return_precisions:
                ret.append(self.cumulative_precisions)
            if return_recalls:
                ret.append(self.cumulative_recalls)
            return ret
        return mean_average_precision

    # ------------------------------------------------------------------
    def _labels_of(self, i):
        labels = self.data_generator.labels[i]
        if self.ignore_under_area > 0:      # :538-545, :628-635
            g = self.gt_format
            labels = [l for l in labels
                      if not ((l[g['ymax']] - l[g['ymin']]) * (l[g['xmax']] - l[g['xmin']]) < self.ignore_under_area)]
        return labels

    def _neutral_of(self, i):
        """Neutral flags of image i aligned with `_labels_of(i)`: the `ignore_under_area` filter is applied to the
        flags too, in the recall denominators and in the matching alike (documented deviation: the reference filters
        the labels only and then indexes the unfiltered flags, :551 / :711)."""
        nb = np.asarray(self.data_generator.eval_neutral[i], dtype=bool).reshape(-1)
        if self.ignore_under_area > 0:
            g = self.gt_format
            full = np.asarray(self.data_generator.labels[i], dtype=np.float64)
            if full.size:
                keep = ~((full[:, g['ymax']] - full[:, g['ymin']]) * (full[:, g['xmax']] - full[:, g['xmin']]) < self.ignore_under_area)
                nb = nb[:len(keep)][keep]
        return nb

    def get_num_gt_per_class(self, ignore_neutral_boxes=True, verbose=True, ret=False):
        """reference :494-568."""
        if self.data_generator.labels is None:
            raise ValueError("Computing the number of ground truth boxes per class not possible, no ground truth given.")
        counts = np.zeros(self.n_classes + 1, dtype=int)
        ci = self.gt_format['class_id']
        neutral = self.data_generator.eval_neutral
        for i in range(len(self.data_generator.labels)):
            boxes = np.asarray(self._labels_of(i))
            if boxes.size == 0:
                continue
            cls = boxes[:, ci].astype(int)
            if ignore_neutral_boxes and neutral is not None:
                cls = cls[~self._neutral_of(i)]
            np.add.at(counts, cls, 1)
        self.num_gt_per_class = counts
        if ret:
            return counts

    def match_predictions(self, ignore_neutral_boxes=True, matching_iou_threshold=0.5, border_pixels='include',
                          sorting_algorithm='quicksort', verbose=True, ret=False):
        """reference :570-777, on the device.  Predictions are ordered by descending confidence with ties
        in their original order (what the reference gets with `sorting_algorithm='mergesort'`; numpy's
        'quicksort' leaves the order of exactly equal confidences unspecified).  `verbose=False` keeps the
        reference's behaviour of evaluating only the best prediction of every class (:692-696)."""
        if self.data_generator.labels is None:
            raise ValueError("Matching predictions to ground truth boxes not possible, no ground truth given.")
        if self.prediction_results is None and getattr(self, '_acc', None) is None:
            raise ValueError("There are no prediction results. You must run `predict_on_dataset()` before calling this method.")
        if border_pixels not in _lib.BORDER:
            raise ValueError("`border_pixels` must be one of 'half', 'include' and 'exclude'.")
        g = self.gt_format
        image_ids = [str(x) for x in self.data_generator.image_ids]
        index_of = {}
        for i, s in enumerate(image_ids):
            index_of[s] = i                      # (later duplicates win, like the reference's dict)
        neutral_avail = self.data_generator.eval_neutral is not None
        use_neutral = ignore_neutral_boxes and neutral_avail

        # ground truth, flattened per image
        gt_rows, gt_neu, gt_off = [], [], np.zeros(len(image_ids) + 1, dtype=np.int64)
        for i in range(len(image_ids)):
            lab = np.asarray(self._labels_of(i), dtype=np.float64)
            n = 0 if lab.size == 0 else lab.shape[0]
            if n:
                gt_rows.append(lab[:, [g['class_id'], g['xmin'], g['ymin'], g['xmax'], g['ymax']]])
                if use_neutral:
                    gt_neu.append(self._neutral_of(i).astype(np.uint8))
            gt_off[i + 1] = gt_off[i] + n
        gt = np.ascontiguousarray(np.concatenate(gt_rows, axis=0)) if gt_rows else np.zeros((0, 5))
        neu = np.ascontiguousarray(np.concatenate(gt_neu)) if (use_neutral and gt_neu) else None

        # predictions, flattened per class ('f4' fields like the reference's structured array, :668-675)
        C = self.n_classes
        coff = np.zeros(C + 2, dtype=np.int64)
        if getattr(self, '_acc', None) is not None and self.prediction_results is None:
            # flat records of add_decoded_batch: a stable sort by class gives the per-class lists in append order
            cls_all = np.concatenate(self._acc['cls']) if self._acc['cls'] else np.zeros(0, np.int32)
            ids_all = np.concatenate(self._acc['ids']) if self._acc['ids'] else np.zeros(0, object)
            by_class = np.argsort(cls_all, kind='stable')
            by_class = by_class[(cls_all[by_class] >= 1) & (cls_all[by_class] <= C)]
            coff[2:] = np.cumsum(np.bincount(cls_all[by_class], minlength=C + 1)[1:C + 1])
            uniq, inv = np.unique(ids_all, return_inverse=True) if len(ids_all) else (np.zeros(0, object), np.zeros(0, np.int64))
            lut = np.array([index_of[u] for u in uniq], dtype=np.int32) if len(uniq) else np.zeros(0, np.int32)
            pim = np.ascontiguousarray(lut[inv][by_class]) if len(ids_all) else np.zeros(0, np.int32)
            pcf = np.ascontiguousarray(np.concatenate(self._acc['conf'])[by_class]) if len(ids_all) else np.zeros(0, np.float32)
            pbx = np.ascontiguousarray(np.concatenate(self._acc['box'], axis=0)[by_class]) if len(ids_all) else np.zeros((0, 4), np.float32)
        else:
            imgs, confs, boxes = [], [], []
            for c in range(1, C + 1):
                preds = self.prediction_results[c]
                coff[c + 1] = coff[c] + len(preds)
                if len(preds):
                    imgs.append(np.fromiter((index_of[str(p[0])] for p in preds), dtype=np.int32, count=len(preds)))
                    arr = np.array([p[1:6] for p in preds], dtype=np.float64).astype(np.float32)
                    confs.append(arr[:, 0])
                    boxes.append(arr[:, 1:5])
            pim = np.ascontiguousarray(np.concatenate(imgs)) if imgs else np.zeros(0, np.int32)
            pcf = np.ascontiguousarray(np.concatenate(confs)) if confs else np.zeros(0, np.float32)
            pbx = np.ascontiguousarray(np.concatenate(boxes, axis=0)) if boxes else np.zeros((0, 4), np.float32)
        P = int(coff[C + 1])
        order = np.empty(max(P, 1), np.int32)
        tp = np.empty(max(P, 1), np.int32)
        fp = np.empty(max(P, 1), np.int32)
        ctp = np.empty(max(P, 1), np.int32)
        cfp = np.empty(max(P, 1), np.int32)
        ctx = _lib.get_context()
        _lib.check(ctx.lib.ssdc_voc_match(ctx.handle, _lib.ptr(pim), _lib.ptr(pcf), _lib.ptr(pbx), _lib.ptr(coff), C,
                                         _lib.ptr(gt), _lib.ptr(neu), _lib.ptr(gt_off), len(image_ids),
                                         float(matching_iou_threshold), _lib.BORDER[border_pixels], 0 if verbose else 1,
                                         _lib.ptr(order), _lib.ptr(tp), _lib.ptr(fp), _lib.ptr(ctp), _lib.ptr(cfp)))
        true_positives, false_positives, cum_tp, cum_fp = [[]], [[]], [[]], [[]]
        self.sorted_indices = [[]]
        for c in range(1, C + 1):
            a, b = int(coff[c]), int(coff[c + 1])
            true_positives.append(tp[a:b].astype(int))
            false_positives.append(fp[a:b].astype(int))
            # (the reference appends nothing to the cumulative lists for a class without predictions, which
            # shifts the following classes; here every class keeps its own slot)
            cum_tp.append(ctp[a:b].astype(int))
            cum_fp.append(cfp[a:b].astype(int))
            self.sorted_indices.append(order[a:b].astype(int))
        self.true_positives = true_positives
        self.false_positives = false_positives
        self.cumulative_true_positives = cum_tp
        self.cumulative_false_positives = cum_fp
        if ret:
            return true_positives, false_positives, cum_tp, cum_fp

    def compute_precision_recall(self, verbose=True, ret=False):
        """reference :779-822."""
        if (self.cumulative_true_positives is None) or (self.cumulative_false_positives is None):
            raise ValueError("True and false positives not available. You must run `match_predictions()` before you call this method.")
        if self.num_gt_per_class is None:
            raise ValueError("Number of ground truth boxes per class not available. You must run `get_num_gt_per_class()` before you call this method.")
        precisions, recalls = [[]], [[]]
        for c in range(1, self.n_classes + 1):
            tp = self.cumulative_true_positives[c]
            fp = self.cumulative_false_positives[c]
            with np.errstate(divide='ignore', invalid='ignore'):
                precisions.append(np.where(tp + fp > 0, tp / (tp + fp), 0))
                recalls.append(tp / self.num_gt_per_class[c])
        self.cumulative_precisions = precisions
        self.cumulative_recalls = recalls
        if ret:
            return precisions, recalls

    def compute_average_precisions(self, mode='sample', num_recall_points=11, verbose=True, ret=False):
        """reference :824-925 (pre-2010 k-point sampling or post-2010 integration)."""
        if (self.cumulative_precisions is None) or (self.cumulative_recalls is None):
            raise ValueError("Precisions and recalls not available. You must run `compute_precision_recall()` before you call this method.")
        if mode not in ('sample', 'integrate'):
            raise ValueError("`mode` can be either 'sample' or 'integrate', but received '{}'".format(mode))
        aps = [0.0]
        for c in range(1, self.n_classes + 1):
            prec = np.asarray(self.cumulative_precisions[c])
            rec = np.asarray(self.cumulative_recalls[c])
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
        self.average_precisions = aps
        if ret:
            return aps

    def compute_mean_average_precision(self, ret=True):
        """reference :927-947."""
        if self.average_precisions is None:
            raise ValueError("Average precisions not available. You must run `compute_average_precisions()` before you call this method.")
        self.mean_average_precision = np.average(self.average_precisions[1:])
        if ret:
            return self.mean_average_precision


# Drop-in mode: the reference's Evaluator also runs the model over the dataset (`predict_on_dataset`, Keras + the data
# generator), writes result files, ...  When its module can be imported further down the package path, `Evaluator` is
# the reference's class with ONLY the matching core replaced by the device version; otherwise it is the stand-alone
# class above.
_ref = _dropin.load_shadowed(__package__ or 'eval_utils', 'average_precision_evaluator') if (__package__ or '').split('.')[0] == 'eval_utils' else None
if _ref is not None and hasattr(_ref, 'Evaluator'):
    class Evaluator(_ref.Evaluator):
        match_predictions = DeviceEvaluator.match_predictions
        _labels_of = DeviceEvaluator._labels_of
else:
    Evaluator = DeviceEvaluator
