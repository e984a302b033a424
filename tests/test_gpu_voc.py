"""GPU parity: the VOC matching / average precision core (`eval_utils.average_precision_evaluator.Evaluator`,
`ssdc_voc_match`) against the golden vectors of the real reference and against the live oracle."""
import numpy as np
import pytest

from oracle import cases
from oracle import voc_eval_oracle as voc

from helpers import load_golden

pytestmark = pytest.mark.gpu


class Dataset(object):
    def __init__(self, inp):
        self.labels, self.image_ids, self.eval_neutral = inp['labels'], inp['image_ids'], inp['eval_neutral']


def run_product(case, inp=None, ignore_neutral_boxes=True):
    from jpeg_detection_resnet_ssd_b200.eval_utils.average_precision_evaluator import Evaluator
    inp = inp or cases.build_voc_input(case)
    ev = Evaluator(model=None, n_classes=inp['n_classes'], data_generator=Dataset(inp),
                   ignore_under_area=case.get('ignore_under_area', 0))
    ev.set_predictions(inp['prediction_results'])
    num = ev.get_num_gt_per_class(ignore_neutral_boxes=ignore_neutral_boxes, verbose=False, ret=True)
    tp, fp, ctp, cfp = ev.match_predictions(ignore_neutral_boxes=ignore_neutral_boxes, verbose=case.get('verbose', True),
                                            ret=True, **case['kwargs'])
    ev.compute_precision_recall(verbose=False)
    ap_s = ev.compute_average_precisions(mode='sample', num_recall_points=11, verbose=False, ret=True)
    ap_i = ev.compute_average_precisions(mode='integrate', verbose=False, ret=True)
    return ev, num, tp, fp, ctp, cfp, ap_s, ap_i


@pytest.mark.parametrize('case', cases.VOC_CASES, ids=lambda c: c['name'])
def test_voc_golden(ctx, case):
    g = load_golden(case['name'])
    ev, num, tp, fp, ctp, cfp, ap_s, ap_i = run_product(case)
    assert np.array_equal(num, g['num_gt'])
    for c in range(1, ev.n_classes + 1):
        assert np.array_equal(tp[c], g['tp_%d' % c]), c
        assert np.array_equal(fp[c], g['fp_%d' % c]), c
        assert np.array_equal(ctp[c], np.cumsum(g['tp_%d' % c]))
        assert np.array_equal(cfp[c], np.cumsum(g['fp_%d' % c]))
    assert np.array_equal(np.asarray(ap_s, dtype=float), g['ap_sample'])
    assert np.array_equal(np.asarray(ap_i, dtype=float), g['ap_integrate'])


@pytest.mark.parametrize('border', ['half', 'include', 'exclude'])
@pytest.mark.parametrize('use_neutral', [False, True])
def test_voc_live_oracle(ctx, border, use_neutral):
    case = dict(name='live', seed=977 + 3 * use_neutral, n_images=50, n_classes=6, neutral=True, quantize=50,
                kwargs=dict(matching_iou_threshold=0.4, border_pixels=border, sorting_algorithm='mergesort'))
    inp = cases.build_voc_input(case)
    ev, num, tp, fp, ctp, cfp, ap_s, ap_i = run_product(case, inp, ignore_neutral_boxes=use_neutral)
    C = inp['n_classes']
    o_num = voc.get_num_gt_per_class(inp['labels'], inp['eval_neutral'], C, use_neutral, 0)
    o_tp, o_fp, o_ctp, o_cfp = voc.match_predictions(inp['prediction_results'], inp['labels'], inp['image_ids'], inp['eval_neutral'], C,
                                                     ignore_neutral_boxes=use_neutral, **case['kwargs'])
    assert np.array_equal(num, o_num)
    for c in range(1, C + 1):
        assert np.array_equal(tp[c], o_tp[c]) and np.array_equal(fp[c], o_fp[c])
        assert np.array_equal(ctp[c], o_ctp[c]) and np.array_equal(cfp[c], o_cfp[c])
    o_prec, o_rec = voc.compute_precision_recall(o_ctp, o_cfp, o_num, C)
    assert np.array_equal(np.asarray(ap_s, float), np.asarray(voc.compute_average_precisions(o_prec, o_rec, C, 'sample', 11), float))
    assert np.array_equal(np.asarray(ap_i, float), np.asarray(voc.compute_average_precisions(o_prec, o_rec, C, 'integrate'), float))


def test_voc_call_and_empty_classes(ctx):
    """`__call__` from stored predictions; a class without predictions; an image without ground truth."""
    from jpeg_detection_resnet_ssd_b200.eval_utils.average_precision_evaluator import Evaluator
    case = dict(name='c', seed=31, n_images=12, n_classes=4, kwargs=dict())
    inp = cases.build_voc_input(case)
    inp['prediction_results'][2] = []
    ev = Evaluator(model=None, n_classes=4, data_generator=Dataset(inp))
    with pytest.raises(ValueError):
        ev()
    with pytest.raises(NotImplementedError):
        ev.predict_on_dataset()
    ev.set_predictions(inp['prediction_results'])
    m, aps = ev(mode='sample', sorting_algorithm='mergesort', return_average_precisions=True)
    assert aps[2] == 0.0 and len(ev.true_positives[2]) == 0
    o_num = voc.get_num_gt_per_class(inp['labels'], None, 4)
    o_tp, o_fp, _, _ = voc.match_predictions(inp['prediction_results'], inp['labels'], inp['image_ids'], None, 4, sorting_algorithm='mergesort')
    for c in (1, 3, 4):
        assert np.array_equal(ev.true_positives[c], o_tp[c]) and np.array_equal(ev.false_positives[c], o_fp[c])
    assert np.array_equal(ev.num_gt_per_class, o_num)
    assert 0.0 <= m <= 1.0
    with pytest.raises(ValueError):
        ev.match_predictions(border_pixels='outside')
    with pytest.raises(ValueError):
        ev.compute_average_precisions(mode='area')


def test_voc_big_class_and_properties(ctx):
    """One class with more predictions than the shared-memory sort holds (global-memory path).  Properties
    that need no oracle: tp + fp <= 1, every ground-truth box is matched at most once, the true positives
    of a class never exceed its ground-truth count, order is a stable descending sort."""
    from jpeg_detection_resnet_ssd_b200.eval_utils.average_precision_evaluator import Evaluator
    case = dict(name='big', seed=5, n_images=300, n_classes=2, dets_per_image=150, quantize=1000, kwargs=dict())
    inp = cases.build_voc_input(case)
    assert max(len(p) for p in inp['prediction_results']) > 16384
    ev = Evaluator(model=None, n_classes=2, data_generator=Dataset(inp))
    ev.set_predictions(inp['prediction_results'])
    num = ev.get_num_gt_per_class(ret=True)
    tp, fp, ctp, cfp = ev.match_predictions(sorting_algorithm='mergesort', ret=True)
    o_tp, o_fp, _, _ = voc.match_predictions(inp['prediction_results'], inp['labels'], inp['image_ids'], None, 2, sorting_algorithm='mergesort')
    for c in (1, 2):
        conf = np.array([p[1] for p in inp['prediction_results'][c]], dtype=np.float32)
        assert np.array_equal(ev.sorted_indices[c], np.argsort(-conf, kind='mergesort'))
        assert np.all(tp[c] + fp[c] == 1) and tp[c].sum() <= num[c]
        assert np.array_equal(tp[c], o_tp[c]) and np.array_equal(fp[c], o_fp[c])
        assert np.array_equal(ctp[c], np.cumsum(tp[c])) and np.array_equal(cfp[c], np.cumsum(fp[c]))
