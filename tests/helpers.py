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
