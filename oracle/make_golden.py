"""Generates tests/golden/*.npz by running the REAL reference (from /root/reference) and the
oracle restatement on the same seeded inputs, requiring bit-identical outputs (this is what
pins the oracle), and stores the reference outputs in a compact canonical form.

    python oracle/make_golden.py            # only works where /root/reference exists

Test infrastructure; never imported by the product.
"""
from __future__ import division

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cases, ref_loader                      # noqa: E402
from oracle import ssd_codec_oracle as orc                # noqa: E402
import synth

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
EXP_PROBE = np.linspace(-3.0, 3.0, 4001).astype(np.float32)


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


def golden_decode(ref, case):
    y = cases.build_decode_input(case, ref.encoder.SSDInputEncoder)
    log_wh = case.get('log_wh', True)
    mod = ref.decoder if log_wh else ref.decoder_no_log
    kw = dict(case['kwargs'])
    t = time.time()
    r = getattr(mod, case['fn'])(y, **kw)
    t_ref = time.time() - t
    ofn = getattr(orc, case['fn'])
    o = ofn(y, log_wh=log_wh, exp_mode='numpy', **kw)
    assert len(r) == len(o)
    for a, b in zip(r, o):
        assert same(a, b), 'oracle != reference for ' + case['name']
    oa = ofn(y, log_wh=log_wh, exp_mode='numpy', with_anchor_index=True, **kw)
    oc = ofn(y, log_wh=log_wh, exp_mode='cr', with_anchor_index=True, **kw)
    for a, b in zip(r, oa):
        assert np.size(a) == np.size(b) == 0 or same(a, np.asarray(b)[:, :6].astype(a.dtype)), case['name']
    if case['fn'] == 'decode_detections' and all(np.size(a) for a in r) and kw.get('top_k') != 'all' and log_wh:
        # anchor ids straight from the reference's own debug decoder
        dbg = ref.decoder.decode_detections_debug(y, **kw)
        for a, b in zip(dbg, oa):
            b = np.asarray(b)
            # (the debug decoder associates (off * size) * var, so its coordinates may differ by an
            # ulp from decode_detections'; box id, class and confidence must agree exactly)
            assert same(a[:, :3], np.concatenate([b[:, 6:7], b[:, :2]], axis=1)), 'debug anchors differ ' + case['name']

    def canon(per_image):
        moved = []
        for p in per_image:
            if np.size(p) == 0:
                moved.append(np.zeros((0, 7)))
            else:
                p = np.asarray(p, dtype=np.float64)
                moved.append(np.concatenate([p[:, 6:7], p[:, :6]], axis=1))
        return cases.canonical_rows(moved)

    rows, counts = canon(oa)
    rows_cr, counts_cr = canon(oc)
    out_dtype = str(np.asarray(r[0]).dtype) if len(r) else 'float64'
    empty_shapes = json.dumps([list(np.asarray(a).shape) for a in r])
    np.savez_compressed(os.path.join(GOLDEN, case['name'] + '.npz'),
                        input_sha=cases.sha256_of(y), rows=rows, counts=counts, rows_cr=rows_cr,
                        counts_cr=counts_cr, out_dtype=out_dtype, shapes=empty_shapes,
                        case=json.dumps({k: v for k, v in case.items()}))
    print('%-26s ref %.2fs  rows %d  cr-identical %s' % (case['name'], t_ref, rows.shape[0],
                                                        same(rows, rows_cr)))


def golden_encode(ref, case):
    gt = cases.build_encode_input(case)
    kw = synth.layout_kwargs(case['layout'], **case.get('overrides', {}))
    log_wh = case.get('log_wh', True)
    rmod = ref.encoder if log_wh else ref.encoder_no_log
    renc = rmod.SSDInputEncoder(**kw)
    oenc = orc.SSDInputEncoder(log_wh=log_wh, **kw)
    assert same(synth.anchors_of(renc), synth.anchors_of(oenc))
    diag = case.get('diagnostics', False)
    t = time.time()
    with np.errstate(all='ignore'):
        rr = renc(gt, diagnostics=diag)
        t_ref = time.time() - t
        oo = oenc(gt, diagnostics=diag, return_matches=True)
    y_ref = rr[0] if diag else rr
    y_orc, mi = oo[0], oo[-1]
    assert same(y_ref, y_orc), 'oracle != reference for ' + case['name']
    if diag:
        assert same(rr[1], oo[1])
    B, A, W = y_ref.shape
    flat_mi = mi.reshape(-1)
    nz = np.nonzero(flat_mi != -1)[0]
    np.savez_compressed(os.path.join(GOLDEN, case['name'] + '.npz'),
                        input_sha=cases.sha256_of(*gt) if len(gt) else '', y_sha=cases.sha256_of(y_ref),
                        shape=np.array([B, A, W]), nz_idx=nz.astype(np.int64), nz_match=flat_mi[nz],
                        nz_rows=y_ref.reshape(B * A, W)[nz], template_sha=cases.sha256_of(renc.generate_encoding_template(2)),
                        case=json.dumps({k: v for k, v in case.items()}))
    print('%-26s ref %.2fs  matched %d neutral %d' % (case['name'], t_ref, (flat_mi >= 0).sum(), (flat_mi == -2).sum()))


def golden_thin(ref):
    inp = cases.thin_inputs()
    out = dict(inp)
    for fmt in ('corners', 'minmax', 'centroids'):
        for border in ('half', 'include', 'exclude'):
            r = ref.bbox.iou(inp['b1_' + fmt], inp['b2_' + fmt], coords=fmt, mode='outer_product', border_pixels=border)
            o = orc.iou(inp['b1_' + fmt], inp['b2_' + fmt], coords=fmt, mode='outer_product', border_pixels=border)
            assert same(r, o)
            out['iou_outer_%s_%s' % (fmt, border)] = r
            r = ref.bbox.iou(inp['b1_' + fmt], inp['b3_' + fmt], coords=fmt, mode='element-wise', border_pixels=border)
            o = orc.iou(inp['b1_' + fmt], inp['b3_' + fmt], coords=fmt, mode='element-wise', border_pixels=border)
            assert same(r, o)
            out['iou_elem_%s_%s' % (fmt, border)] = r
            r = ref.bbox.iou(inp['b1_' + fmt], inp['b2_' + fmt][0], coords=fmt, mode='element-wise', border_pixels=border)
            out['iou_bcast_%s_%s' % (fmt, border)] = r
            if fmt != 'centroids':
                out['inter_outer_%s_%s' % (fmt, border)] = ref.bbox.intersection_area(
                    inp['b1_' + fmt], inp['b2_' + fmt], coords=fmt, mode='outer_product', border_pixels=border)
    for conv in ('minmax2centroids', 'centroids2minmax', 'corners2centroids', 'centroids2corners',
                 'minmax2corners', 'corners2minmax'):
        for border in ('half', 'include', 'exclude'):
            for key, start in (('conv32', 3), ('conv64', -5)):
                r = ref.bbox.convert_coordinates(inp[key], start_index=start, conversion=conv, border_pixels=border)
                o = orc.convert_coordinates(inp[key], start_index=start, conversion=conv, border_pixels=border)
                assert same(r, o)
                out['conv_%s_%s_%s' % (key, conv, border)] = r
    for key in ('w_small', 'w_quirk'):
        r = ref.matching.match_bipartite_greedy(inp[key])
        assert same(np.asarray(r), np.asarray(orc.match_bipartite_greedy(inp[key])))
        out['bip_' + key] = np.asarray(r, dtype=np.int64)
        g, a = ref.matching.match_multi(inp[key], 0.5)
        go, ao = orc.match_multi(inp[key], 0.5)
        assert same(np.asarray(g), np.asarray(go)) and same(np.asarray(a), np.asarray(ao))
        out['multi_gt_' + key] = np.asarray(g, dtype=np.int64)
        out['multi_anchor_' + key] = np.asarray(a, dtype=np.int64)
    rows = inp['nms_rows']
    r = ref.decoder.greedy_nms([rows, rows[:40]], iou_threshold=0.3, coords='corners', border_pixels='half')
    o = orc.greedy_nms([rows, rows[:40]], iou_threshold=0.3, coords='corners', border_pixels='half')
    assert all(same(a, b) for a, b in zip(r, o))
    out['nms_full'] = r[0]
    out['nms_40'] = r[1]
    out['nms1'] = ref.decoder._greedy_nms(rows[:, 1:], iou_threshold=0.45, coords='corners', border_pixels='include')
    out['nms2'] = ref.decoder._greedy_nms2(rows, iou_threshold=0.45, coords='corners', border_pixels='half')
    np.savez_compressed(os.path.join(GOLDEN, 'thin_ops.npz'), **out)
    print('thin ops: %d arrays' % len(out))


def golden_voc(evmod, case):
    from oracle import voc_eval_oracle as vo
    inp = cases.build_voc_input(case)
    C = inp['n_classes']

    class DG(object):
        pass
    dg = DG()
    dg.labels, dg.image_ids, dg.eval_neutral = inp['labels'], inp['image_ids'], inp['eval_neutral']
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ev = evmod.Evaluator(model=None, n_classes=C, data_generator=dg, model_mode='inference',
                             ignore_under_area=case.get('ignore_under_area', 0))
    ev.prediction_results = inp['prediction_results']
    verbose = case.get('verbose', True)
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        num = ev.get_num_gt_per_class(ignore_neutral_boxes=True, verbose=False, ret=True)
        tp, fp, ctp, cfp = ev.match_predictions(ignore_neutral_boxes=True, verbose=verbose, ret=True, **case['kwargs'])
        prec, rec = ev.compute_precision_recall(verbose=False, ret=True)
        ap_s = ev.compute_average_precisions(mode='sample', num_recall_points=11, verbose=False, ret=True)
        ap_i = ev.compute_average_precisions(mode='integrate', verbose=False, ret=True)
    o_num = vo.get_num_gt_per_class(inp['labels'], inp['eval_neutral'], C, True, case.get('ignore_under_area', 0))
    o_tp, o_fp, o_ctp, o_cfp = vo.match_predictions(inp['prediction_results'], inp['labels'], inp['image_ids'], inp['eval_neutral'], C,
                                                    ignore_neutral_boxes=True, verbose=verbose,
                                                    ignore_under_area=case.get('ignore_under_area', 0), **case['kwargs'])
    assert same(np.asarray(num), np.asarray(o_num))
    for c in range(1, C + 1):
        assert same(np.asarray(tp[c]), np.asarray(o_tp[c])) and same(np.asarray(fp[c]), np.asarray(o_fp[c])), case['name']
        assert same(np.asarray(ctp[c]), np.asarray(o_ctp[c])) and same(np.asarray(cfp[c]), np.asarray(o_cfp[c]))
    o_prec, o_rec = vo.compute_precision_recall(o_ctp, o_cfp, o_num, C)
    for c in range(1, C + 1):
        assert same(np.asarray(prec[c]), np.asarray(o_prec[c])) and same(np.asarray(rec[c]), np.asarray(o_rec[c]))
    assert np.array_equal(np.asarray(ap_s, dtype=float), np.asarray(vo.compute_average_precisions(o_prec, o_rec, C, 'sample', 11), dtype=float))
    assert np.array_equal(np.asarray(ap_i, dtype=float), np.asarray(vo.compute_average_precisions(o_prec, o_rec, C, 'integrate'), dtype=float))
    out = dict(num_gt=np.asarray(num), ap_sample=np.asarray(ap_s, dtype=float), ap_integrate=np.asarray(ap_i, dtype=float),
               case=json.dumps(case))
    for c in range(1, C + 1):
        out['tp_%d' % c] = np.asarray(tp[c]); out['fp_%d' % c] = np.asarray(fp[c])
    np.savez_compressed(os.path.join(GOLDEN, case['name'] + '.npz'), **out)
    print('%-26s preds %d  tp %d  mAP(sample) %.4f' % (case['name'], sum(len(p) for p in inp['prediction_results']),
                                                      sum(int(np.sum(tp[c])) for c in range(1, C + 1)), float(np.mean(ap_s[1:]))))


def golden_evalprep():
    """Decoder -> evaluator glue and BoxFilter: the reference's own functions and inverter closures on the seeded
    inputs; the numpy restatement must agree bit for bit."""
    from oracle import eval_prep_oracle as ep
    ns = ref_loader.load_data_generator_utils()
    inp = cases.build_evalprep_input()
    ref_inv = [ref_loader.reference_inverters(ns, sp) for sp in inp['specs']]
    r = ns.misc.apply_inverse_transforms([np.copy(p) for p in inp['preds']], ref_inv)

    def orc_inverters(specs):
        out = []
        for sp in specs:
            if sp is None:
                out.append(None)
            elif sp[0] == 'resize':
                out.append(ep.resize_inverter(*sp[1:]))
            elif sp[0] == 'translate':
                out.append(ep.translate_inverter(*sp[1:]))
            else:
                out.append(lambda labels: labels)
        return out
    o = ep.apply_inverse_transforms(inp['preds'], [orc_inverters(sp) for sp in inp['specs']])
    assert all(same(a, b) for a, b in zip(r, o))
    out = {}
    for i, a in enumerate(r):
        out['inv_%d' % i] = np.asarray(a)
    # result records: the loop of average_precision_evaluator.py:405-422, executed literally on the reference's output
    for rc in (False, 2):
        img, cls, conf, box = [], [], [], []
        for k, batch_item in enumerate(r):
            for b in np.asarray(batch_item).reshape(-1, 6):
                img.append(k); cls.append(int(b[0]))
                conf.append(round(b[1], rc) if rc else b[1])
                box.append([round(b[2], 1), round(b[3], 1), round(b[4], 1), round(b[5], 1)])
        rec = np.array([tuple([c] + bx) for c, bx in zip(conf, box)], dtype=[('c', 'f4'), ('x0', 'f4'), ('y0', 'f4'), ('x1', 'f4'), ('y1', 'f4')])
        oi, oc, of, ob = ep.evaluation_records(o, rc)
        assert same(np.array(img, np.int32), oi) and same(np.array(cls, np.int32), oc) and same(rec['c'], of)
        assert same(np.stack([rec['x0'], rec['y0'], rec['x1'], rec['y1']], 1), ob)
        tag = 'rc%d' % int(rc)
        out['rec_img_' + tag], out['rec_cls_' + tag], out['rec_conf_' + tag], out['rec_box_' + tag] = oi, oc, of, ob
    # BoxFilter
    for ci, cfg in enumerate(cases.BOXFILTER_CONFIGS):
        bf = ns.validation.BoxFilter(**cfg)
        for li, lab in enumerate(inp['labels']):
            H, W = 300 + 7 * li, 280 + 11 * li
            want = bf(lab, image_height=H, image_width=W)
            lower, upper = cfg.get('overlap_bounds', (0.3, 1.0))
            mask = ep.box_filter(lab, H, W, check_overlap=cfg.get('check_overlap', True), check_min_area=cfg.get('check_min_area', True),
                                 check_degenerate=cfg.get('check_degenerate', True), overlap_criterion=cfg.get('overlap_criterion', 'center_point'),
                                 lower=lower, upper=upper, min_area=cfg.get('min_area', 16), border_pixels=cfg.get('border_pixels', 'half'))
            assert same(np.asarray(want), lab[mask]), (ci, li)
            out['bf_%d_%d' % (ci, li)] = mask
    # ImageValidator on the same labels
    iv = ns.validation.ImageValidator(overlap_criterion='area', bounds=(0.5, 1.0), n_boxes_min=3)
    out['iv_area3'] = np.array([bool(iv(lab, 300, 300)) if len(lab) else False for lab in inp['labels']])
    iv = ns.validation.ImageValidator(overlap_criterion='center_point', n_boxes_min='all')
    out['iv_all'] = np.array([bool(iv(lab, 300, 300)) for lab in inp['labels']])
    np.savez_compressed(os.path.join(GOLDEN, 'evalprep.npz'), **out)
    print('evalprep: %d arrays' % len(out))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref = ref_loader.load()
    only = sys.argv[1:]
    np.savez_compressed(os.path.join(GOLDEN, 'exp_probe.npz'), x=EXP_PROBE, y=np.exp(EXP_PROBE))
    for case in cases.DECODE_CASES:
        if not only or case['name'] in only:
            golden_decode(ref, case)
    for case in cases.ENCODE_CASES:
        if not only or case['name'] in only:
            golden_encode(ref, case)
    if not only or 'thin' in only:
        golden_thin(ref)
    if not only or 'evalprep' in only:
        golden_evalprep()
    evmod = ref_loader.load_evaluator()
    for case in cases.VOC_CASES:
        if not only or case['name'] in only:
            golden_voc(evmod, case)


if __name__ == '__main__':
    main()
