"""GPU parity tests of the decoder -> evaluator glue and the augmentation box checks
against the goldens written from the reference's own functions (tests/golden/evalprep.npz) and the numpy oracle."""
import numpy as np
import pytest

from oracle import cases
from oracle import eval_prep_oracle as ep
from oracle import ssd_codec_oracle as orc
from oracle import voc_eval_oracle as voc
import synth
from jpeg_detection_resnet_ssd_b200 import _lib
from jpeg_detection_resnet_ssd_b200.data_generator import object_detection_2d_misc_utils as mu
from jpeg_detection_resnet_ssd_b200.data_generator.object_detection_2d_image_boxes_validation_utils import (
    BoxFilter, BoundGenerator, ImageValidator)
from jpeg_detection_resnet_ssd_b200.eval_utils.average_precision_evaluator import Evaluator
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_output_decoder import decode_detections

from helpers import load_golden

pytestmark = pytest.mark.gpu


def _descriptors(specs):
    out = []
    for sp in specs:
        if sp is None:
            out.append(None)
        elif sp[0] == 'resize':
            out.append(mu.ResizeInverter(*sp[1:]))
        elif sp[0] == 'translate':
            out.append(mu.TranslateInverter(*sp[1:]))
        else:
            out.append(None)
    return out


def test_apply_inverse_transforms_matches_reference(ctx):
    g = load_golden('evalprep')
    inp = cases.build_evalprep_input()
    inv = [_descriptors(sp) for sp in inp['specs']]
    n0 = ctx.launch_count()
    out = mu.apply_inverse_transforms([np.copy(p) for p in inp['preds']], inv)
    assert ctx.launch_count() == n0 + 1                       # one kernel for the whole batch
    assert isinstance(out, list) and len(out) == len(inp['preds'])
    for i, a in enumerate(out):
        assert a.dtype == np.float64 and np.array_equal(a, g['inv_%d' % i]), i
    # array form (B, k, 6)
    k = 5
    arr = np.stack([p[:k] for p in inp['preds'] if p.shape[0] >= k][:4])
    sel = [i for i, p in enumerate(inp['preds']) if p.shape[0] >= k][:4]
    got = mu.apply_inverse_transforms(arr, [inv[i] for i in sel])
    assert isinstance(got, np.ndarray) and got.shape == arr.shape
    for j, i in enumerate(sel):
        assert np.array_equal(got[j], g['inv_%d' % i][:k])
    # an arbitrary callable keeps working (host path, like the reference)
    out2 = mu.apply_inverse_transforms([np.copy(inp['preds'][0])], [[lambda l: l + 1.0]])
    assert np.array_equal(out2[0], inp['preds'][0] + 1.0)
    with pytest.raises(ValueError):
        mu.apply_inverse_transforms(3, [])


@pytest.mark.parametrize('round_conf', [False, 3])
def test_evaluation_records_from_device_resident_decode(round_conf, ctx):
    """decode_detections leaves its rows on the device; `add_decoded_batch` turns them into the Evaluator's records
    (inverse transforms, rounding, float32) without a host loop - equal to the reference's loop on the host results."""
    enc = synth.make_encoder(SSDInputEncoder, 'ssd300')
    B = 7
    y = synth.synth_y_pred(synth.anchors_of(enc), enc.variances, 21, B, 41, bg_bias=8.5, hot=30)
    dec = decode_detections(y, 0.01, 0.45, 200, 'centroids', True, 300, 300)
    specs = [[('resize', 375 + 10 * i, 500 - 7 * i, 300, 300)] if i % 3 else [('translate', 3 * i, -i), ('resize', 333, 444, 300, 300)]
             for i in range(B)]
    specs[5] = []
    inv = [_descriptors(sp) for sp in specs]
    from tests_support_inverters import oracle_inverters
    want_rows = ep.apply_inverse_transforms(dec, [oracle_inverters(sp) for sp in specs])
    w_img, w_cls, w_conf, w_box = ep.evaluation_records(want_rows, round_conf)

    class DS(object):
        pass
    ds = DS()
    ids = ['img%03d' % i for i in range(B)]
    ds.image_ids, ds.eval_neutral = ids, None
    rng = np.random.default_rng(3)
    ds.labels = [np.concatenate([rng.integers(1, 21, size=(4, 1)), np.sort(rng.uniform(0, 400, size=(4, 4)), axis=1)[:, [0, 1, 2, 3]]], axis=1) for _ in range(B)]
    ev = Evaluator(model=None, n_classes=20, data_generator=ds)
    n = ev.add_decoded_batch(ids, inverse_transforms=inv, round_confidences=round_conf)
    assert n == len(w_img)
    acc = ev._acc
    assert np.array_equal(acc['cls'][0], w_cls) and np.array_equal(acc['conf'][0], w_conf) and np.array_equal(acc['box'][0], w_box)
    assert list(acc['ids'][0]) == [ids[i] for i in w_img]
    # the matcher fed with the flat records == the matcher fed with the reference's list-of-tuples structure
    res = [list() for _ in range(21)]
    for i in range(len(w_img)):
        res[int(w_cls[i])].append((ids[w_img[i]], float(w_conf[i]), float(w_box[i, 0]), float(w_box[i, 1]), float(w_box[i, 2]), float(w_box[i, 3])))
    a = ev.match_predictions(sorting_algorithm='mergesort', ret=True)
    ev2 = Evaluator(model=None, n_classes=20, data_generator=ds)
    ev2.set_predictions(res)
    b = ev2.match_predictions(sorting_algorithm='mergesort', ret=True)
    for xa, xb in zip(a, b):
        for c in range(1, 21):
            assert np.array_equal(xa[c], xb[c])
    assert ev.materialize_prediction_results()[1:] == [[(t[0], np.float32(t[1]), *[np.float32(v) for v in t[2:]]) for t in r] for r in res[1:]]


def test_box_filter_and_image_validator_match_reference(ctx):
    g = load_golden('evalprep')
    inp = cases.build_evalprep_input()
    for ci, cfg in enumerate(cases.BOXFILTER_CONFIGS):
        bf = BoxFilter(**cfg)
        for li, lab in enumerate(inp['labels']):
            got = bf(lab, image_height=300 + 7 * li, image_width=280 + 11 * li)
            assert np.array_equal(got, lab[g['bf_%d_%d' % (ci, li)]]), (ci, li)
        # the whole batch in one launch
        n0 = ctx.launch_count()
        outs = bf.filter_batch(inp['labels'], [300 + 7 * li for li in range(12)], [280 + 11 * li for li in range(12)])
        assert ctx.launch_count() == n0 + 1
        for li, lab in enumerate(inp['labels']):
            assert np.array_equal(outs[li], lab[g['bf_%d_%d' % (ci, li)]])
    iv = ImageValidator(overlap_criterion='area', bounds=(0.5, 1.0), n_boxes_min=3)
    assert [bool(iv(lab, 300, 300)) if len(lab) else False for lab in inp['labels']] == g['iv_area3'].tolist()
    iv = ImageValidator(overlap_criterion='center_point', n_boxes_min='all')
    assert [bool(iv(lab, 300, 300)) for lab in inp['labels']] == g['iv_all'].tolist()
    # integer label arrays (what the generators hand over) and a BoundGenerator
    lab = inp['labels'][0].astype(np.int64)
    bf = BoxFilter(overlap_criterion='iou', overlap_bounds=BoundGenerator(sample_space=((0.05, None),)))
    want = ep.box_filter(lab.astype(float), 300, 300, overlap_criterion='iou', lower=0.05, upper=1.0)
    assert np.array_equal(bf(lab, 300, 300), lab[want])
    with pytest.raises(ValueError):
        BoxFilter(overlap_criterion='nope')
    with pytest.raises(ValueError):
        ImageValidator(n_boxes_min=0)
    with pytest.raises(ValueError):
        BoundGenerator(sample_space=((0.5, 0.1),))


def test_inverter_descriptors_are_device_callables(ctx):
    inp = cases.build_evalprep_input()
    g = load_golden('evalprep')
    p = inp['preds'][0]
    n0 = ctx.launch_count()
    out = mu.ResizeInverter(*inp['specs'][0][0][1:])(p)
    assert ctx.launch_count() == n0 + 1 and np.array_equal(out, g['inv_0'])
    with pytest.raises(TypeError):
        mu.TranslateInverter(1, 2)(p.astype(np.float32))
