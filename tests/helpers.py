"""Shared helpers of the parity tests."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    return g


def golden_case(g):
    return json.loads(str(g['case']))


def host_exp_matches_golden():
    """True when this host's float32 np.exp reproduces the generating host's bit for bit."""
    p = load_golden('exp_probe')
    return np.array_equal(np.exp(p['x']), p['y'])


def to_rows7(per_image_with_anchor_last):
    """list of (k, 7) [class, conf, 4 coords, anchor] -> canonical (rows7, counts)."""
    from oracle import cases
    moved = []
    for p in per_image_with_anchor_last:
        if np.size(p) == 0:
            moved.append(np.zeros((0, 7)))
        else:
            p = np.asarray(p, dtype=np.float64)
            moved.append(np.concatenate([p[:, 6:7], p[:, :6]], axis=1))
    return cases.canonical_rows(moved)


def product_rows7(rows, counts, idx):
    """run_decode output -> canonical (rows7, counts)."""
    from oracle import cases
    per = []
    pos = 0
    for c in counts:
        c = int(c)
        r = np.concatenate([idx[pos:pos + c, None].astype(np.float64), rows[pos:pos + c]], axis=1)
        per.append(r)
        pos += c
    return cases.canonical_rows(per)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.maximum(np.abs(b), 1e-30)
    with np.errstate(invalid='ignore'):
        e = np.abs(a - b) / denom
    e[(a == b) | (np.isnan(a) & np.isnan(b))] = 0
    return e


def explain_index_mismatches(got, got_counts, want, want_counts, iou_thr, border='half', agnostic=False,
                             top_k=None, margin=1e-5):
    """For canonical row sets that differ in their kept-box indices, proves that every first
    divergence of a greedy-NMS keep list is *threshold-ambiguous*: the box present in only one of
    the two lists overlaps an earlier commonly kept box with |IoU - iou_thr| <= margin, i.e. the
    decision flips with an ulp-level change of the box coordinates (float32 np.exp is not
    correctly rounded and host dependent, SURVEY section 7 hard part 2).  Returns the number of
    explained divergences; raises AssertionError for an unexplained one."""
    from oracle import ssd_codec_oracle as orc
    explained = 0
    go = np.concatenate([[0], np.cumsum(got_counts)]).astype(int)
    wo = np.concatenate([[0], np.cumsum(want_counts)]).astype(int)
    for b in range(len(got_counts)):
        g = got[go[b]:go[b + 1]]
        w = want[wo[b]:wo[b + 1]]
        keys = [None] if agnostic else sorted(set(g[:, 1]) | set(w[:, 1]))
        for c in keys:
            gs = g if c is None else g[g[:, 1] == c]
            ws = w if c is None else w[w[:, 1] == c]
            gs = gs[np.lexsort((gs[:, 0], -gs[:, 2]))]
            ws = ws[np.lexsort((ws[:, 0], -ws[:, 2]))]
            n = min(len(gs), len(ws))
            i = 0
            while i < n and gs[i, 0] == ws[i, 0]:
                i += 1
            if i == len(gs) and i == len(ws):
                continue
            truncated = top_k not in (None, 'all') and (got_counts[b] >= top_k or want_counts[b] >= top_k)
            if truncated and (i >= len(gs) or i >= len(ws)):
                continue                      # a tail cut by the cross-class top-k, not an NMS decision
            # the candidate that comes first in canonical order and is missing from the other list
            if i >= len(ws) or (i < len(gs) and (gs[i, 2], -gs[i, 0]) > (ws[i, 2], -ws[i, 0])):
                x = gs[i]
            else:
                x = ws[i]
            prefix = gs[:i]
            assert len(prefix) > 0, 'first kept box differs: not an NMS threshold effect'
            ious = orc.iou(prefix[:, 3:7], x[3:7], coords='corners', mode='element-wise', border_pixels=border)
            gap = np.min(np.abs(ious - iou_thr))
            assert gap <= margin, ('unexplained kept-index mismatch in image %d segment %s: nearest IoU gap %.3g'
                                   % (b, c, gap))
            explained += 1
    return explained
