#!/usr/bin/env python
"""bench.py - benchmark of the SSD box codec hot path on B200 (one JSON line per run).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config 0..4] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

`--config` selects one of BASELINE.json's five configurations (0-based; default 2, the one the metric is
quoted on).  A "step" is one pass of the hot path over one batch of synthetic input:

  0  decode_detections, SSD300, batch 8 (the reference's own CPU-runnable case)
  1  SSDInputEncoder.__call__, ssd_custom / SSD300 anchors, batch 32 (the reference's training batch)
  2  decode_detections (decode + NMS), SSD300, batch 1024 per GPU            <- headline
  3  decode_detections, SSD512 layout (24564 anchors), conf 0.001, dense candidates, batch 512
  4  SSDInputEncoder -> decode_detections_fast round trip on float64 targets, batch 4096 over the GPUs

Every line carries
  value      whole-job images/s with the step's inputs resident in HBM (device-timed with CUDA events on the
             library's own stream over EXACTLY `--steps` steps, max over ranks)
  sustained  the same loop continued for >= 0.2 s (the contract's K steps last a few milliseconds; the clock
             sampler and the driver's own sampler need a longer window)
  e2e        the same metric through the public Python call with HOST buffers (H2D + kernels + D2H every step),
             plus the measured plain-copy ceiling of the same bytes (`h2d_peak_gbs` / `d2h_peak_gbs`)
  roofline   the dominant kernel: algorithmic bytes per launch / its event-timed duration vs the measured HBM peak
  cpu_baseline  the numpy oracle (restatement of the reference, same cost structure) on ONE host core over a
             bounded sample of the same workload
  parity_checked  the GPU results of that same sample compared with the oracle's (indices / classes / matches
             exact, coordinates and offsets <= 1e-5 relative); the run FAILS if they differ

`--impl reference` times the reference's CPU implementation (the numpy oracle port: the reference is Python and does
not travel to the GPU box) on all host cores, same config dict, bounded sample per step.
`--scaling strong` shards the configuration's batch over the ranks (fixed total work); default weak (fixed work per
GPU), except config 4 whose batch of 4096 is defined over the whole job.
"""
from __future__ import division

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_CLASSES = 21          # VOC incl. background
W_ROW = N_CLASSES + 12
UNIT = 'images/s'

CONFIGS = {
    0: dict(kind='decode', layout='ssd300', img=300, A=8732, batch=8, conf=0.01, iou=0.45, top_k=200, bg_bias=8.0, unique=8,
            scaling='weak', metric='decoded images/sec (decode+NMS, SSD300 VOC)',
            workload='configs[0]: decode_detections conf 0.01 / IoU 0.45 / top_k 200 on synthetic SSD300 y_pred '
                     '(8732 anchors, 21 VOC classes, float32), batch 8'),
    1: dict(kind='encode', layout='ssd300', img=300, A=8732, batch=32, scaling='weak',
            metric='encoded images/sec (SSDInputEncoder, ssd_custom / SSD300 anchors)',
            workload='configs[1]: SSDInputEncoder.__call__ on synthetic Pascal VOC ground truth (1..20 boxes / image), '
                     'ssd_custom ResNet50-DCT anchor layout (8732 anchors), batch 32, float64 y_encoded'),
    2: dict(kind='decode', layout='ssd300', img=300, A=8732, batch=1024, conf=0.01, iou=0.45, top_k=200, bg_bias=8.0, unique=256,
            scaling='weak', metric='decoded images/sec (decode+NMS, SSD300 VOC)',
            workload='SSD300 decode+NMS (configs[2]): decode_detections conf 0.01 / IoU 0.45 / top_k 200 on synthetic '
                     'y_pred (8732 anchors, 21 VOC classes, float32)'),
    3: dict(kind='decode', layout='ssd512', img=512, A=24564, batch=512, conf=0.001, iou=0.45, top_k=200, bg_bias=6.0, unique=32,
            scaling='weak', metric='decoded images/sec (decode+NMS, SSD512 layout, dense candidates)',
            workload='configs[3]: decode_detections conf 0.001 / IoU 0.45 / top_k 200 on synthetic SSD512-layout y_pred '
                     '(24564 anchors, 21 classes, float32), dense low-threshold candidates, batch 512'),
    4: dict(kind='roundtrip', layout='ssd300', img=300, A=8732, batch=4096, scaling='strong',
            metric='round-trip images/sec (SSDInputEncoder + decode_detections_fast, SSD300 VOC)',
            workload='configs[4]: SSDInputEncoder -> decode_detections_fast(conf 0.5, IoU 0.45, top_k all) on the float64 '
                     'y_encoded, synthetic VOC ground truth, batch 4096 over the whole job'),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument('--scaling', default=None, choices=['weak', 'strong'])
    ap.add_argument('--batch', type=int, default=None, help='override the configuration\'s batch')
    ap.add_argument('--bg-bias', type=float, default=None, help='background logit bias: candidate density of y_pred')
    ap.add_argument('--no-extra', action='store_true', help='skip the secondary numbers (loss, VOC matcher, density sweep)')
    ap.add_argument('--no-cpu', action='store_true', help='skip cpu_baseline + parity gate (profiling runs only)')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-buffer end-to-end leg (profiling runs only)')
    ap.add_argument('--sustain-ms', type=float, default=200.0)
    ap.add_argument('--floor-target', type=int, default=None, help='diagnosis: SSDC_OPT_FLOOR_TARGET of the context')
    ap.add_argument('--no-pipeline', type=int, default=None, help='diagnosis: SSDC_OPT_NO_PIPELINE of the context')
    ap.add_argument('--enc-lanes', type=int, default=None, help='diagnosis: SSDC_OPT_ENC_LANES of the context')
    ap.add_argument('--d1-ctas', type=int, default=None, help='diagnosis: SSDC_OPT_D1_CTAS of the context')
    ap.add_argument('--no-l2-hints', type=int, default=None, help='diagnosis: SSDC_OPT_NO_L2_HINTS of the context')
    ap.add_argument('--d1-warps', type=int, default=None, help='diagnosis: SSDC_OPT_D1_WARPS of the context')
    ap.add_argument('--sync-steps', action='store_true', help='diagnosis: synchronise after every warm-up step')
    return ap.parse_args()


def load_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def shard_range(total, n, i):
    """Contiguous batch shards of ceil(total / n) images (same rule as libssdcodec's contexts)."""
    per = (total + n - 1) // n
    return min(total, per * i), min(total, per * (i + 1))


def whole_job_throughput(images_all_ranks_per_step, steps, max_ms):
    """images/s of the whole job: all ranks together processed `images_all_ranks_per_step` images per step."""
    return images_all_ranks_per_step * steps / (max_ms / 1e3)


# ---------------------------------------------------------------------------------------------------
# plumbing: torch.distributed (barrier + max over ranks only), clocks, NUMA
# ---------------------------------------------------------------------------------------------------
def init_dist(world, local_rank, backend=None):
    """torch.distributed is plumbing only (barrier + max over ranks); one process per GPU."""
    if world <= 1:
        return None
    import torch
    import torch.distributed as dist
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend='nccl', device_id=torch.device('cuda', local_rank))
    else:
        dist.init_process_group(backend=backend)
    return dist


def max_over_ranks(x, dist):
    """Durations are combined as the MAX over ranks."""
    if dist is None:
        return float(x)
    import torch
    dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bind_to_gpu_numa(index):
    """Pins this process (and so its pinned host allocations, first touch) to the CPUs of the GPU's NUMA node.
    Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(':')[0]) == 8:
            bus = bus[4:]
        with open('/sys/bus/pci/devices/%s/numa_node' % bus) as fh:
            node = int(fh.read().strip())
        if node < 0:
            return 'numa node unknown'
        with open('/sys/devices/system/node/node%d/cpulist' % node) as fh:
            cpus = set()
            for part in fh.read().strip().split(','):
                a, _, b = part.partition('-')
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return 'node %d (%d cpus)' % (node, len(allowed))
        return 'node %d (no allowed cpu there)' % node
    except Exception as exc:
        return 'unbound (%s)' % type(exc).__name__


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4): 'sw_power_cap',
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.001)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------
_ENCODERS = {}


def make_encoder(layout):
    """The product SSDInputEncoder of a named anchor layout (cached)."""
    import synth
    from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
    if layout not in _ENCODERS:
        _ENCODERS[layout] = synth.make_encoder(SSDInputEncoder, layout)
    return _ENCODERS[layout]


def anchors_and_variances(layout, product=True):
    """(A, 4) anchors + variances of a layout.  The CPU arm takes them from the oracle's encoder (no GPU library)."""
    import synth
    if product:
        enc = make_encoder(layout)
    else:
        from oracle import ssd_codec_oracle as orc
        enc = synth.make_encoder(orc.SSDInputEncoder, layout)
    return synth.anchors_of(enc), enc.variances


def synth_decode_input(cfg, batch, seed, pinned, bg_bias, product=True):
    """Synthetic y_pred (batch, A, 33) float32 of the configuration's layout: `unique` distinct images tiled."""
    import synth
    anchors, variances = anchors_and_variances(cfg['layout'], product)
    unique = max(1, min(cfg['unique'], batch))
    base = synth.synth_y_pred(anchors, variances, N_CLASSES, unique, seed, bg_bias=bg_bias, hot=40)
    shape = (batch, cfg['A'], W_ROW)
    if pinned:
        from jpeg_detection_resnet_ssd_b200 import pinned_empty
        y = pinned_empty(shape, np.float32)
    else:
        y = np.empty(shape, np.float32)
    for i in range(0, batch, unique):
        n = min(unique, batch - i)
        y[i:i + n] = base[:n]
    cands = float((base[:, :, 1:N_CLASSES] > cfg['conf']).sum()) / unique
    return y, cands


def decode_params(_lib, cfg):
    p = _lib.DecodeParams()
    p.mode = _lib.MODE_PER_CLASS
    p.input_coords = _lib.COORDS['centroids']
    p.normalize = 1
    p.border_pixels = _lib.BORDER['half']
    p.top_k = cfg['top_k']
    p.nms_cap = 0
    p.log_wh = 1
    p.do_nms = 1
    p.conf_thresh = cfg['conf']
    p.iou_thresh = cfg['iou']
    p.img_h = float(cfg['img'])
    p.img_w = float(cfg['img'])
    return p


def fast_params(_lib):
    pf = _lib.DecodeParams()
    pf.mode, pf.input_coords, pf.normalize, pf.border_pixels = _lib.MODE_FAST, 0, 1, 0
    pf.top_k, pf.nms_cap, pf.log_wh, pf.do_nms = 0, 0, 1, 1
    pf.conf_thresh, pf.iou_thresh, pf.img_h, pf.img_w = 0.5, 0.45, 300.0, 300.0
    return pf


def canonical7(per_image):
    """list of (k, 7) [anchor, class, conf, 4 coords] -> rows sorted per image by (class, -conf, anchor) + counts."""
    counts = np.array([0 if np.size(p) == 0 else np.asarray(p).shape[0] for p in per_image], dtype=np.int64)
    blocks = []
    for p in per_image:
        if np.size(p) == 0:
            continue
        p = np.asarray(p, dtype=np.float64)
        blocks.append(p[np.lexsort((p[:, 0], -p[:, 2], p[:, 1]))])
    return (np.concatenate(blocks, axis=0) if blocks else np.zeros((0, 7))), counts


def compare_rows7(got, got_counts, want, want_counts, what):
    """Parity bar of north_star: kept-box indices and classes bit-exact, coordinates <= 1e-5 relative (float32)."""
    if not np.array_equal(got_counts, want_counts):
        bad = np.nonzero(got_counts != want_counts)[0]
        raise SystemExit('PARITY FAILURE (%s): detections per image differ in images %s' % (what, bad[:8].tolist()))
    if not np.array_equal(got[:, :2], want[:, :2]):
        raise SystemExit('PARITY FAILURE (%s): kept anchor indices / classes differ from the oracle' % what)
    denom = np.maximum(np.abs(want[:, 2:]), 1e-30)
    err = np.abs(got[:, 2:] - want[:, 2:]) / denom
    err[got[:, 2:] == want[:, 2:]] = 0.0
    mx = float(err.max(initial=0.0))
    if not mx <= 1e-5:
        raise SystemExit('PARITY FAILURE (%s): confidence / coordinates off by %.3g relative (> 1e-5)' % (what, mx))
    return mx, bool(np.array_equal(got, want))


class DecodeWorkload(object):
    """configs[0], [2], [3]: decode_detections on synthetic y_pred."""

    def __init__(self, cfg, args, ctx, _lib, rank, world, batch):
        self.cfg, self.ctx, self._lib, self.B = cfg, ctx, _lib, batch
        self.bg_bias = cfg['bg_bias'] if args.bg_bias is None else args.bg_bias
        self.A = cfg['A']
        self.y, self.cands = synth_decode_input(cfg, batch, 1234 + rank, True, self.bg_bias)
        self.in_bytes = int(self.y.nbytes)
        # inputs smaller than 2x the 126 MB L2 are rotated through enough distinct device copies that no step finds
        # its rows in L2 (footprint >= 384 MB); larger ones exceed L2 by themselves
        self.copies = 1 if self.in_bytes > 252e6 else int(math.ceil(384e6 / self.in_bytes))
        self.d_y = [ctx.dev_alloc(self.in_bytes) for _ in range(self.copies)]
        for d in self.d_y:
            ctx.h2d(d, self.y)
        self.p = decode_params(_lib, cfg)
        self.i = 0
        self.alg_bytes = batch * self.A * W_ROW * 4

    def l2_note(self):
        if self.copies == 1:
            return 'input per step (%.0f MB) exceeds the 126 MB L2; no flush needed' % (self.in_bytes / 1e6)
        return 'input per step is %.1f MB: steps rotate through %d distinct device copies (%.0f MB footprint > 3x L2)' % (
            self.in_bytes / 1e6, self.copies, self.copies * self.in_bytes / 1e6)

    def step(self):
        _lib = self._lib
        d = self.d_y[self.i % self.copies]
        self.i += 1
        _lib.check(self.ctx.lib.ssdc_decode_submit(self.ctx.handle, d, _lib.F32, 1, self.B, self.A, N_CLASSES, _lib.C.byref(self.p)))

    def device_results(self, n):
        """Rows of the first n images of the last step, as [anchor, class, conf, coords]."""
        _lib, B, K = self._lib, self.B, self.cfg['top_k']
        counts = np.zeros(B, np.int32)
        total = _lib.C.c_int64(0)
        rows = np.empty((B * K, 6))
        idx = np.empty(B * K, np.int32)
        _lib.check(self.ctx.lib.ssdc_decode_collect(self.ctx.handle, _lib.ptr(rows), B * K, _lib.ptr(counts), _lib.ptr(idx), _lib.C.byref(total)))
        assert total.value > 0 and counts.max() <= K
        per, pos = [], 0
        for b in range(n):
            c = int(counts[b])
            per.append(np.concatenate([idx[pos:pos + c, None].astype(np.float64), rows[pos:pos + c]], axis=1))
            pos += c
        return per

    def roofline(self, prof, prof_steps, peak, peak_src):
        d1_ms = prof['decode_filter'][0] / max(prof['decode_filter'][1], 1)
        achieved = self.alg_bytes / (d1_ms / 1e3) / 1e9
        step_ms = sum(v[0] for v in prof.values()) / prof_steps
        traffic = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
                tj = json.load(fh)
            key = 'decode_filter_tma_kernel@config%d' % self.cfg['id']
            if key in tj and abs(self.bg_bias - self.cfg['bg_bias']) < 1e-9 and self.B == self.cfg['batch']:
                traffic = tj[key]['dram_bytes_per_launch']
        except Exception:
            pass
        return {'bound': 'hbm', 'kernel': 'decode_filter_tma_kernel<float,false> (D1: stream y_pred, threshold, compaction)',
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                'peak_source': peak_src, 'kernel_ms': d1_ms, 'algorithmic_bytes_per_launch': self.alg_bytes,
                'whole_step_frac_of_hbm_bound': self.alg_bytes / (step_ms / 1e3) / 1e9 / peak,
                'kernel_ms_per_step': {k: round(v[0] / prof_steps, 4) for k, v in prof.items() if v[1]},
                'share_of_step': {k: round(v[0] / prof_steps / step_ms, 4) for k, v in prof.items() if v[1]}}

    # -- public API with host buffers --
    def e2e_call(self, y):
        from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_output_decoder import decode_detections
        c = self.cfg
        return decode_detections(y, c['conf'], c['iou'], c['top_k'], 'centroids', True, c['img'], c['img'])

    def e2e(self, steps, barrier, dist, world_images):
        out = None
        for _ in range(2):
            out = self.e2e_call(self.y)
        if self.in_bytes < 64e6:
            steps = max(steps, 100)              # (a call on a small batch is a fraction of a millisecond: a longer loop, stated in `steps`)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = self.e2e_call(self.y)
        dt = max_over_ranks(time.perf_counter() - t0, dist)
        n_rows = sum(o.shape[0] for o in out if o.size)
        e = {'value': world_images * steps / dt, 'unit': UNIT, 'h2d_bytes_per_step': self.in_bytes,
             'd2h_bytes_per_step': int(n_rows * 52 + self.B * 4 + 8), 'steps': steps,
             'api': 'ssd_output_decoder.decode_detections(y_pred: pinned host ndarray) -> list of ndarrays'}
        # the plain-copy ceiling of the same bytes, all ranks at once
        reps = 3
        self.ctx.h2d(self.d_y[0], self.y)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            self.ctx.h2d(self.d_y[0], self.y)
        tc = max_over_ranks(time.perf_counter() - t0, dist) / reps
        e['h2d_peak_gbs'] = self.in_bytes / tc / 1e9
        e['frac_of_h2d_peak'] = (self.in_bytes * steps / dt) / (self.in_bytes / tc)
        # what model.predict() hands the drop-in is pageable memory
        yp = np.array(self.y)
        psteps = max(2, steps // 3)
        self.e2e_call(yp)
        barrier()
        t0 = time.perf_counter()
        for _ in range(psteps):
            self.e2e_call(yp)
        dtp = max_over_ranks(time.perf_counter() - t0, dist)
        e['pageable_value'] = world_images * psteps / dtp
        e['pageable_note'] = 'same call with a plain (pageable) numpy y_pred, what model.predict() returns'
        return e

    # -- CPU side --
    def cpu_sample(self, budget_s=12.0, max_images=128):
        from oracle import ssd_codec_oracle as orc
        c = self.cfg

        def run(y):
            return orc.decode_detections(y, c['conf'], c['iou'], c['top_k'], 'centroids', True, c['img'], c['img'],
                                         exp_mode='cr', with_anchor_index=True)
        t = time.perf_counter()
        first = run(self.y[:1])
        per = time.perf_counter() - t
        n = int(max(1, min(max_images, self.B, budget_s / max(per, 1e-6))))
        if n > 1:
            t = time.perf_counter()
            out = run(self.y[:n])
            dt = time.perf_counter() - t
        else:
            out, dt = first, per
        # oracle rows are [class, conf, coords, anchor] -> [anchor, class, conf, coords]
        want = [np.zeros((0, 7)) if np.size(o) == 0 else np.concatenate([o[:, 6:7], o[:, :6]], axis=1) for o in out]
        base = {'value': n / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                'sample': 'numpy oracle (restatement of the reference decode_detections), first %d image(s) of the '
                          'workload, 1 thread, %.2f s' % (n, dt)}
        return base, n, want

    def parity(self, n, want):
        got, gc = canonical7(self.device_results(n))
        wr, wc = canonical7(want)
        mx, exact = compare_rows7(got, gc, wr, wc, 'decode_detections, first %d images' % n)
        keys, floored, fallback = self.ctx.decode_stats()
        return {'images': n, 'ok': True, 'rows': int(gc.sum()), 'indices_classes': 'exact', 'max_rel_err': mx, 'bit_identical': exact,
                'whole_batch': {'candidates_emitted_per_image': round(keys / self.B, 1), 'images_with_score_floor': floored,
                                'images_rescanned_without_floor': fallback}}

    def free(self):
        for d in self.d_y:
            self.ctx.dev_free(d)


class EncodeWorkload(object):
    """configs[1]: SSDInputEncoder.__call__ (y_encoded stays on the device in the device-timed loop)."""

    def __init__(self, cfg, args, ctx, _lib, rank, world, batch):
        import synth
        self.cfg, self.ctx, self._lib, self.B = cfg, ctx, _lib, batch
        self.A = cfg['A']
        self.enc = make_encoder(cfg['layout'])
        self.gt = synth.synth_ground_truth(cfg['img'], cfg['img'], 20, batch, seed=77 + rank)
        self.flat, self.offs = synth.flatten_ground_truth(self.gt)
        _, self.h = self.enc._encoder()
        self.out_bytes = batch * self.A * W_ROW * 8
        self.copies = 1 if self.out_bytes > 252e6 else int(math.ceil(384e6 / self.out_bytes))
        self.d_out = [ctx.dev_alloc(self.out_bytes) for _ in range(self.copies)]
        self.i = 0
        self.alg_bytes = self.out_bytes

    def l2_note(self):
        if self.copies == 1:
            return 'output per step (%.0f MB) exceeds the 126 MB L2; no flush needed' % (self.out_bytes / 1e6)
        return 'output per step is %.1f MB: steps rotate through %d distinct device buffers (%.0f MB footprint > 3x L2)' % (
            self.out_bytes / 1e6, self.copies, self.copies * self.out_bytes / 1e6)

    def step(self):
        _lib = self._lib
        d = self.d_out[self.i % self.copies]
        self.i += 1
        _lib.check(self.ctx.lib.ssdc_encode(self.h, _lib.ptr(self.flat), _lib.ptr(self.offs), self.B, 1, d, None, None))

    def roofline(self, prof, prof_steps, peak, peak_src):
        w_ms = prof['enc_write'][0] / max(prof['enc_write'][1], 1)
        achieved = self.alg_bytes / (w_ms / 1e3) / 1e9
        return {'bound': 'hbm', 'kernel': 'template_tma_kernel (E3: y_encoded write-out, TMA bulk stores)',
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None,
                'peak_source': peak_src, 'kernel_ms': w_ms, 'algorithmic_bytes_per_launch': self.alg_bytes,
                'kernel_ms_per_step': {k: round(v[0] / prof_steps, 4) for k, v in prof.items() if v[1]},
                'note': 'kernel_ms from a serialised profiling pass; in the timed step the write-out stream overlaps the matching kernels'}

    def e2e(self, steps, barrier, dist, world_images):
        y = None
        for _ in range(2):
            y = self.enc(self.gt)
        if self.out_bytes < 256e6:
            steps = max(steps, 40)               # (a call on a small batch is 1-2 ms: a longer loop, stated in `steps`)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            y = self.enc(self.gt)
        dt = max_over_ranks(time.perf_counter() - t0, dist)
        e = {'value': world_images * steps / dt, 'unit': UNIT, 'h2d_bytes_per_step': int(self.flat.nbytes + self.offs.nbytes),
             'd2h_bytes_per_step': int(y.nbytes), 'steps': steps,
             'api': 'SSDInputEncoder.__call__(list of (m_i, 5) ndarrays) -> (B, A, 33) float64 host ndarray'}
        reps = 3
        self.ctx.d2h(y, self.d_out[0])
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            self.ctx.d2h(y, self.d_out[0])
        tc = max_over_ranks(time.perf_counter() - t0, dist) / reps
        e['d2h_peak_gbs'] = y.nbytes / tc / 1e9
        e['frac_of_d2h_peak'] = (y.nbytes * steps / dt) / (y.nbytes / tc)
        e['note'] = 'the result array is pageable numpy memory, like the reference\'s; d2h_peak is the same copy alone'
        return e

    def cpu_sample(self, budget_s=12.0, max_images=256):
        import synth
        from oracle import ssd_codec_oracle as orc
        oenc = synth.make_encoder(orc.SSDInputEncoder, self.cfg['layout'])
        t = time.perf_counter()
        oenc(self.gt[:2])
        per = (time.perf_counter() - t) / 2
        n = int(max(1, min(max_images, self.B, budget_s / max(per, 1e-6))))
        t = time.perf_counter()
        want = oenc(self.gt[:n], return_matches=True)
        dt = time.perf_counter() - t
        base = {'value': n / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                'sample': 'numpy oracle (restatement of the reference SSDInputEncoder.__call__), first %d images of the '
                          'workload, 1 thread, %.2f s' % (n, dt)}
        return base, n, want

    def parity(self, n, want):
        y_ref, mi_ref = want
        y, mi = self.enc(self.gt[:n], return_matches=True)
        if not np.array_equal(mi, mi_ref):
            raise SystemExit('PARITY FAILURE (encode): matched-anchor assignments differ from the oracle')
        C = N_CLASSES
        if not (np.array_equal(y[:, :, :C], y_ref[:, :, :C]) and np.array_equal(y[:, :, C + 4:], y_ref[:, :, C + 4:])):
            raise SystemExit('PARITY FAILURE (encode): class columns / anchor tail differ from the oracle')
        denom = np.maximum(np.abs(y_ref), 1e-30)
        err = np.abs(y - y_ref) / denom
        err[y == y_ref] = 0
        mx = float(err.max())
        if not mx <= 1e-5:
            raise SystemExit('PARITY FAILURE (encode): offsets off by %.3g relative' % mx)
        # ... and the device-resident output the timed loop wrote is the same tensor
        got = np.empty((self.B, self.A, W_ROW))
        self.ctx.synchronize()
        self.ctx.d2h(got, self.d_out[(self.i - 1) % self.copies])
        y_all = self.enc(self.gt)
        if not np.array_equal(got, y_all):
            raise SystemExit('PARITY FAILURE (encode): device-resident y_encoded of the timed loop differs from the host-path result')
        return {'images': n, 'ok': True, 'matched_anchors': int((mi >= 0).sum()), 'matches': 'exact', 'max_rel_err': mx,
                'bit_identical': bool(np.array_equal(y, y_ref))}

    def free(self):
        for d in self.d_out:
            self.ctx.dev_free(d)


class RoundTripWorkload(object):
    """configs[4]: SSDInputEncoder -> decode_detections_fast on the float64 y_encoded (every positive confidence is
    exactly 1.0, so every NMS decision is a tie broken by the anchor index)."""

    def __init__(self, cfg, args, ctx, _lib, rank, world, batch):
        import synth
        self.cfg, self.ctx, self._lib, self.B = cfg, ctx, _lib, batch
        self.A = cfg['A']
        self.enc = make_encoder(cfg['layout'])
        self.gt = synth.synth_ground_truth(300, 300, 20, batch, seed=78 + 1000 * rank)
        self.flat, self.offs = synth.flatten_ground_truth(self.gt)
        _, self.h = self.enc._encoder()
        self.enc_bytes = batch * self.A * W_ROW * 8
        self.d_enc = ctx.dev_alloc(self.enc_bytes)
        self.pf = fast_params(_lib)
        self.alg_bytes = self.enc_bytes          # per kernel: written once by the encoder, read once by the decoder

    def l2_note(self):
        return 'y_encoded per step (%.0f MB) exceeds the 126 MB L2; no flush needed' % (self.enc_bytes / 1e6)

    def step(self):
        _lib, lib = self._lib, self.ctx.lib
        _lib.check(lib.ssdc_encode(self.h, _lib.ptr(self.flat), _lib.ptr(self.offs), self.B, 1, self.d_enc, None, None))
        _lib.check(lib.ssdc_decode_submit(self.ctx.handle, self.d_enc, _lib.F64, 1, self.B, self.A, N_CLASSES, _lib.C.byref(self.pf)))

    def device_results(self, n):
        _lib = self._lib
        counts = np.zeros(self.B, np.int32)
        total = _lib.C.c_int64(0)
        cap = int(self.offs[-1]) + self.B
        rows = np.empty((cap, 6))
        idx = np.empty(cap, np.int32)
        _lib.check(self.ctx.lib.ssdc_decode_collect(self.ctx.handle, _lib.ptr(rows), cap, _lib.ptr(counts), _lib.ptr(idx), _lib.C.byref(total)))
        per, pos = [], 0
        for b in range(n):
            c = int(counts[b])
            per.append(np.concatenate([idx[pos:pos + c, None].astype(np.float64), rows[pos:pos + c]], axis=1))
            pos += c
        return per, int(total.value)

    def roofline(self, prof, prof_steps, peak, peak_src):
        w_ms = prof['enc_write'][0] / max(prof['enc_write'][1], 1)
        d_ms = prof['decode_filter'][0] / max(prof['decode_filter'][1], 1)
        if w_ms >= d_ms:
            dom, ms = 'template_tma_kernel (E3: y_encoded write-out)', w_ms
        else:
            dom, ms = 'decode_filter_tma_kernel<double,true> (D1 on float64 y_encoded)', d_ms
        achieved = self.alg_bytes / (ms / 1e3) / 1e9
        step_ms = sum(v[0] for v in prof.values()) / prof_steps
        return {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': None, 'peak_source': peak_src, 'kernel_ms': ms, 'algorithmic_bytes_per_launch': self.alg_bytes,
                'whole_step_frac_of_hbm_bound': 2 * self.alg_bytes / (step_ms / 1e3) / 1e9 / peak,
                'kernel_ms_per_step': {k: round(v[0] / prof_steps, 4) for k, v in prof.items() if v[1]},
                'note': 'the step moves y_encoded twice (written by the encoder, read by the decoder): whole_step_frac counts both'}

    def e2e_call(self):
        from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_output_decoder import decode_detections_fast
        y = self.enc(self.gt)
        return y, decode_detections_fast(y, 0.5, 0.45, 'all', 'centroids', True, 300, 300)

    def e2e(self, steps, barrier, dist, world_images):
        steps = max(2, min(steps, 3))
        y, out = self.e2e_call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            y, out = self.e2e_call()
        dt = max_over_ranks(time.perf_counter() - t0, dist)
        n_rows = sum(o.shape[0] for o in out if o.size)
        return {'value': world_images * steps / dt, 'unit': UNIT,
                'h2d_bytes_per_step': int(self.flat.nbytes + self.offs.nbytes + y.nbytes),
                'd2h_bytes_per_step': int(y.nbytes + n_rows * 52 + self.B * 4 + 8), 'steps': steps,
                'api': 'y = SSDInputEncoder.__call__(gt); decode_detections_fast(y, 0.5, 0.45, "all") - host ndarrays in between, as the '
                       'reference\'s scripts would call them (y_encoded crosses PCIe twice)'}

    def cpu_sample(self, budget_s=12.0, max_images=64):
        import synth
        from oracle import ssd_codec_oracle as orc
        oenc = synth.make_encoder(orc.SSDInputEncoder, self.cfg['layout'])

        def run(gt):
            y = oenc(gt)
            return y, orc.decode_detections_fast(y, 0.5, 0.45, 'all', 'centroids', True, 300, 300)
        t = time.perf_counter()
        run(self.gt[:2])
        per = (time.perf_counter() - t) / 2
        n = int(max(1, min(max_images, self.B, budget_s / max(per, 1e-6))))
        t = time.perf_counter()
        _, out = run(self.gt[:n])
        dt = time.perf_counter() - t
        base = {'value': n / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                'sample': 'numpy oracle (SSDInputEncoder + decode_detections_fast), first %d images of the workload, 1 thread, %.2f s' % (n, dt)}
        return base, n, out

    def parity(self, n, want):
        got_per, total = self.device_results(n)
        # rows [class, conf, coords]: compare as ordered lists (descending score = keep order; all scores tie at 1.0)
        for b in range(n):
            w = np.asarray(want[b], dtype=np.float64).reshape(-1, 6) if np.size(want[b]) else np.zeros((0, 6))
            g = got_per[b][:, 1:]
            if g.shape != w.shape or not np.array_equal(g[:, :2], w[:, :2]):
                raise SystemExit('PARITY FAILURE (round trip): kept boxes of image %d differ from the oracle' % b)
            denom = np.maximum(np.abs(w[:, 2:]), 1e-30)
            err = np.abs(g[:, 2:] - w[:, 2:]) / denom
            if err.size and not float(err.max()) <= 1e-5:
                raise SystemExit('PARITY FAILURE (round trip): coordinates of image %d off by %.3g' % (b, float(err.max())))
        return {'images': n, 'ok': True, 'rows': int(sum(p.shape[0] for p in got_per)), 'decoded_boxes_whole_batch': total,
                'ground_truth_boxes_whole_batch': int(self.offs[-1]), 'indices_classes': 'exact'}

    def free(self):
        self.ctx.dev_free(self.d_enc)


WORKLOADS = {'decode': DecodeWorkload, 'encode': EncodeWorkload, 'roundtrip': RoundTripWorkload}


# ---------------------------------------------------------------------------------------------------
# the reference arm: the numpy oracle port on all host cores
# ---------------------------------------------------------------------------------------------------
def _ref_decode_chunk(args):
    y, conf, iou, top_k, img = args
    from oracle import ssd_codec_oracle as orc
    out = orc.decode_detections(y, conf, iou, top_k, 'centroids', True, img, img)
    return sum(int(np.size(o) // 6) for o in out)


_REF_ENC = {}


def _ref_encode_chunk(args):
    gt, layout, roundtrip = args
    import synth
    from oracle import ssd_codec_oracle as orc
    enc = _REF_ENC.get(layout)
    if enc is None:
        enc = _REF_ENC[layout] = synth.make_encoder(orc.SSDInputEncoder, layout)
    y = enc(gt)
    if roundtrip:
        out = orc.decode_detections_fast(y, 0.5, 0.45, 'all', 'centroids', True, 300, 300)
        return sum(int(np.size(o) // 6) for o in out)
    return int(y.shape[0])


def run_reference(args, cfg, rank, world):
    """--impl reference: the reference's CPU implementation of the path (numpy oracle port) on all host cores; each
    step is a bounded sample of the configuration's workload (per-image cost is additive)."""
    if rank != 0:
        return
    import multiprocessing as mp
    import synth
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    batch, scaling = job_batch(args, cfg)
    if cfg['kind'] == 'decode':
        bg = cfg['bg_bias'] if args.bg_bias is None else args.bg_bias
        # per-image cost decides the sample: one image per core per step for the dense SSD512 case, else up to 4
        per_step = cores if cfg['conf'] < 0.005 else max(cores, min(64, 4 * cores))
        per_step = min(per_step, max(batch, 1))
        full, cands = synth_decode_input(cfg, min(batch, max(per_step, cfg['unique'])), 1234, False, bg, product=False)
        y = full[:per_step]
        chunks = [(np.ascontiguousarray(c), cfg['conf'], cfg['iou'], cfg['top_k'], cfg['img'])
                  for c in np.array_split(y, min(cores, per_step)) if c.shape[0]]
        fn = _ref_decode_chunk
    else:
        cands = None
        per_step = min(batch, max(cores, min(256, 8 * cores)))
        gt = synth.synth_ground_truth(cfg['img'], cfg['img'], 20, per_step, seed=77)
        n_chunks = min(cores, per_step)
        bounds = np.linspace(0, per_step, n_chunks + 1).astype(int)
        chunks = [(gt[bounds[i]:bounds[i + 1]], cfg['layout'], cfg['kind'] == 'roundtrip') for i in range(n_chunks)]
        fn = _ref_encode_chunk
    procs = min(cores, len(chunks))
    with mp.get_context('fork').Pool(procs) as pool:
        for _ in range(max(1, min(args.warmup, 2))):
            pool.map(fn, chunks)
        t = time.perf_counter()
        for _ in range(args.steps):
            pool.map(fn, chunks)
        dt = time.perf_counter() - t
    value = per_step * args.steps / dt
    line = {
        'impl': 'reference', 'metric': cfg['metric'], 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None, 'dtype': 'f32' if cfg['kind'] == 'decode' else 'f64',
        'data': 'synthetic', 'config': workload_config(args, cfg, batch, cands),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': procs, 'kind': 'port',
                         'sample': 'numpy oracle (restatement of the reference; the reference is Python and does not travel to the GPU '
                                   'box) over %d processes; each step is a bounded sample of %d images of the workload (the cost is '
                                   'additive per image)' % (procs, per_step)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def job_batch(args, cfg):
    """(batch, scaling): weak = the configuration's batch on every GPU, strong = that batch sharded over the ranks."""
    scaling = args.scaling or cfg['scaling']
    batch = args.batch if args.batch is not None else cfg['batch']
    return batch, scaling


def workload_config(args, cfg, batch, cands):
    """Identical for both arms (the driver compares the dicts): describes the workload, not the run."""
    scaling = args.scaling or cfg['scaling']
    c = {'workload': cfg['workload'], 'config_index': cfg['id'], 'anchors': cfg['A'], 'classes': N_CLASSES,
         'parallelism': 'batch shards, one process per GPU, no collective'}
    if scaling == 'weak':
        c['batch_per_gpu'] = batch
    else:
        c['global_batch'] = batch
    if cfg['kind'] == 'decode':
        c['bg_bias'] = cfg['bg_bias'] if args.bg_bias is None else args.bg_bias
        c['candidates_per_image'] = None if cands is None else round(cands, 1)
        c['unique_images'] = min(cfg['unique'], batch)
    c['l2'] = 'device arm: per-step inputs / outputs exceed the 126 MB L2 or rotate through distinct buffers (see l2_policy)'
    return c


# ---------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    cfg = dict(CONFIGS[args.config], id=args.config)

    if args.impl == 'reference':
        run_reference(args, cfg, rank, world)
        return

    dist = init_dist(world, local_rank)
    numa = bind_to_gpu_numa(local_rank)

    from __graft_entry__ import build
    if rank == 0:
        build()
    if dist is not None:
        dist.barrier()

    from jpeg_detection_resnet_ssd_b200 import _lib, Context, set_context
    ctx = Context([local_rank])
    set_context(ctx)
    if args.floor_target is not None:
        ctx.set_option('floor_target', args.floor_target)
    if args.no_pipeline is not None:
        ctx.set_option('no_pipeline', args.no_pipeline)
    if args.enc_lanes is not None:
        ctx.set_option('enc_lanes', args.enc_lanes)
    if args.d1_ctas is not None:
        ctx.set_option('d1_ctas', args.d1_ctas)
    if args.d1_warps is not None:
        ctx.set_option('d1_warps', args.d1_warps)
    if args.no_l2_hints is not None:
        ctx.set_option('no_l2_hints', args.no_l2_hints)

    batch, scaling = job_batch(args, cfg)
    if scaling == 'strong':
        b0, b1 = shard_range(batch, world, rank)
        my_batch = b1 - b0
        images_all = batch
    else:
        my_batch = batch
        images_all = batch * world
    wl = WORKLOADS[cfg['kind']](cfg, args, ctx, _lib, rank, world, max(my_batch, 1))

    def barrier():
        ctx.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- device-resident timing: EXACTLY args.steps steps ---------------------------------------------
    warm = max(args.warmup, 3)
    for i in range(warm):
        wl.step()
        if args.sync_steps:
            ctx.synchronize()
            print('step %d ok' % i, file=sys.stderr, flush=True)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    ctx.timer_start()
    for _ in range(args.steps):
        wl.step()
    ms = ctx.timer_stop()
    launches = ctx.launch_count() - launches0
    barrier()
    ms = max_over_ranks(ms, dist)
    value = whole_job_throughput(images_all, args.steps, ms)

    # ---- the same loop sustained (clock sampling window) ----------------------------------------------
    sus_steps = int(min(20000, max(args.steps, math.ceil(args.sustain_ms / max(ms / args.steps, 1e-4)))))
    ctx.timer_start()
    for _ in range(sus_steps):
        wl.step()
    sus_ms = max_over_ranks(ctx.timer_stop(), dist)
    barrier()
    clocks = sampler.finish()
    sustained = {'steps': sus_steps, 'ms_per_step': sus_ms / sus_steps, 'value': whole_job_throughput(images_all, sus_steps, sus_ms)}

    # ---- per-kernel timing of the dominant kernel (separate pass, events around every launch) ---------
    peak, peak_src = load_peaks()
    ctx.profile_enable(True)
    prof_steps = 5
    for _ in range(prof_steps):
        wl.step()
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    roofline = wl.roofline(prof, prof_steps, peak, peak_src)

    # ---- CPU baseline + parity gate on the same sample (rank 0) ---------------------------------------
    cpu_baseline = parity = None
    if rank == 0 and not args.no_cpu:
        wl.step()
        cpu_baseline, n_par, want = wl.cpu_sample()
        parity = wl.parity(n_par, want)
    barrier()

    # ---- end to end through the public API, host buffers ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = wl.e2e(max(3, min(args.steps, 10)), barrier, dist, images_all)

    line = {
        'metric': cfg['metric'], 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': warm, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': scaling,
        'vs_baseline': None, 'dtype': 'f32' if cfg['kind'] == 'decode' else 'f64', 'data': 'synthetic',
        'config': workload_config(args, cfg, batch, getattr(wl, 'cands', None)),
        'l2_policy': wl.l2_note(), 'images_per_rank_per_step': my_batch, 'host_numa': numa,
        'clocks': clocks, 'sustained': sustained, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline,
    }
    if parity is not None:
        line['parity_checked'] = parity
    if cpu_baseline is not None:
        line['cpu_baseline' if world == 1 else 'cpu_baseline_rank0'] = cpu_baseline
    if rank == 0 and world == 1 and not args.no_extra and args.config == 2:
        line['extra'] = extra_numbers(ctx, _lib, peak)
    wl.free()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
# secondary numbers of the headline run (not the metric): candidate-density sweep, loss, VOC matcher
# ---------------------------------------------------------------------------------------------------
def extra_numbers(ctx, _lib, peak):
    out = {}
    lib = ctx.lib
    cfg = dict(CONFIGS[2], id=2)
    for bias in (6.0, 10.0):
        try:
            y, cands = synth_decode_input(dict(cfg, unique=64), 1024, 4321, False, bias)
            d_y = ctx.dev_alloc(y.nbytes)
            ctx.h2d(d_y, y)
            p = decode_params(_lib, cfg)
            for _ in range(3):
                _lib.check(lib.ssdc_decode_submit(ctx.handle, d_y, _lib.F32, 1, 1024, cfg['A'], N_CLASSES, _lib.C.byref(p)))
            ctx.synchronize()
            steps = 20
            ctx.timer_start()
            for _ in range(steps):
                _lib.check(lib.ssdc_decode_submit(ctx.handle, d_y, _lib.F32, 1, 1024, cfg['A'], N_CLASSES, _lib.C.byref(p)))
            ms = ctx.timer_stop()
            ctx.profile_enable(True)
            for _ in range(3):
                _lib.check(lib.ssdc_decode_submit(ctx.handle, d_y, _lib.F32, 1, 1024, cfg['A'], N_CLASSES, _lib.C.byref(p)))
            prof = ctx.profile_read()
            ctx.profile_enable(False)
            out['decode_ssd300_b1024_bias%g' % bias] = {
                'candidates_per_image': round(cands, 1), 'images_per_s': 1024 * steps / (ms / 1e3), 'ms_per_step': ms / steps,
                'step_frac_of_hbm_bound': 1024 * cfg['A'] * W_ROW * 4 / (ms / steps / 1e3) / 1e9 / peak,
                'kernel_ms': {k: round(v[0] / 3, 4) for k, v in prof.items() if v[1]}}
            ctx.dev_free(d_y)
        except Exception as exc:      # secondary number: never take the headline down
            out['decode_ssd300_b1024_bias%g' % bias] = {'error': repr(exc)[:200]}
    try:
        out['ssd_loss_b1024'] = loss_numbers(ctx, _lib, peak)
    except Exception as exc:
        out['ssd_loss_b1024'] = {'error': repr(exc)[:200]}
    try:
        out['voc_match_predictions'] = voc_numbers(ctx)
    except Exception as exc:
        out['voc_match_predictions'] = {'error': repr(exc)[:200]}
    # the other single-GPU BASELINE configurations, device-timed in this process (their full lines - roofline, cpu_baseline,
    # e2e, parity gate - come from `bench.py --config N`; profiles/r02_bench_c*.json)
    for c in (0, 1, 3):
        try:
            out['config%d_device_timed' % c] = other_config_numbers(ctx, _lib, c)
        except Exception as exc:
            out['config%d_device_timed' % c] = {'error': repr(exc)[:200]}
    try:
        out['evaluator_records_from_device'] = records_numbers(ctx, _lib)
    except Exception as exc:
        out['evaluator_records_from_device'] = {'error': repr(exc)[:200]}
    return out


def other_config_numbers(ctx, _lib, c):
    cfg = dict(CONFIGS[c], id=c)

    class A(object):
        bg_bias = None
    wl = WORKLOADS[cfg['kind']](cfg, A(), ctx, _lib, 0, 1, cfg['batch'])
    try:
        for _ in range(5):
            wl.step()
        ctx.synchronize()
        steps = 20
        ctx.timer_start()
        for _ in range(steps):
            wl.step()
        ms = ctx.timer_stop()
        return {'workload': cfg['workload'], 'images_per_s': cfg['batch'] * steps / (ms / 1e3), 'ms_per_step': ms / steps,
                'l2_policy': wl.l2_note()}
    finally:
        ctx.synchronize()
        wl.free()


def records_numbers(ctx, _lib):
    """the Evaluator's result records (inverse transforms + rounding, average_precision_evaluator.py
    :402-422) taken from the device-resident rows of a 1024-image decode (`Evaluator.add_decoded_batch`) against the
    reference's per-detection Python loop (oracle port) on the same detections."""
    from oracle import eval_prep_oracle as ep
    from jpeg_detection_resnet_ssd_b200.data_generator import object_detection_2d_misc_utils as mu
    from jpeg_detection_resnet_ssd_b200.eval_utils.average_precision_evaluator import Evaluator
    from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_output_decoder import decode_detections
    cfg = dict(CONFIGS[2], id=2)
    B = 1024
    y, _ = synth_decode_input(dict(cfg, unique=64), B, 777, True, 8.0)
    dec = decode_detections(y, cfg['conf'], cfg['iou'], cfg['top_k'], 'centroids', True, 300, 300)

    class DS(object):
        pass
    ds = DS()
    ds.image_ids, ds.eval_neutral, ds.labels = ['%06d' % i for i in range(B)], None, [np.zeros((0, 5)) for _ in range(B)]
    inv = [[mu.ResizeInverter(375, 500, 300, 300)] for _ in range(B)]
    ev = Evaluator(model=None, n_classes=20, data_generator=ds)
    ev.add_decoded_batch(ds.image_ids, inverse_transforms=inv)
    ev.reset_predictions()
    t0 = time.perf_counter()
    n = ev.add_decoded_batch(ds.image_ids, inverse_transforms=inv)
    t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    rows = ep.apply_inverse_transforms(dec, [[ep.resize_inverter(375, 500, 300, 300)] for _ in range(B)])
    rec = ep.evaluation_records(rows, False)
    t_cpu = time.perf_counter() - t0
    same = bool(np.array_equal(rec[1], ev._acc['cls'][0]) and np.array_equal(rec[2], ev._acc['conf'][0]) and np.array_equal(rec[3], ev._acc['box'][0]))
    return {'detections': int(n), 'device_path_ms': t_dev * 1e3, 'detections_per_s': n / t_dev, 'cpu_port_ms': t_cpu * 1e3,
            'cpu_port_detections_per_s': n / t_cpu, 'identical_records': same,
            'note': 'device path = Evaluator.add_decoded_batch on the rows decode_detections left in HBM (Resize inverter, rounding, float32 '
                    'records, D2H); cpu port = apply_inverse_transforms + the per-detection loop of the reference on the host results'}


def loss_numbers(ctx, _lib, peak):
    """`ssdc_ssd_loss` with both tensors resident on the device: y_true = the encoder's float64 output,
    y_pred = synthetic float32 predictions (64 unique images tiled)."""
    import synth
    lib = ctx.lib
    B, A, W = 1024, 8732, W_ROW
    enc = make_encoder('ssd300')
    _, h = enc._encoder()
    gt = synth.synth_ground_truth(300, 300, 20, B, seed=79)
    flat, offs = synth.flatten_ground_truth(gt)
    d_true = ctx.dev_alloc(B * A * W * 8)
    _lib.check(lib.ssdc_encode(h, _lib.ptr(flat), _lib.ptr(offs), B, 1, d_true, None, None))
    uniq = 64
    anchors, variances = anchors_and_variances('ssd300')
    base = synth.synth_y_pred(anchors, variances, N_CLASSES, uniq, 99, bg_bias=3.0, hot=40)
    d_pred = ctx.dev_alloc(B * A * W * 4)
    for i in range(0, B, uniq):
        _lib.check(lib.ssdc_memcpy_h2d(ctx.handle, 0, _lib.C.c_void_p(d_pred.value + i * A * W * 4), _lib.ptr(base), base.nbytes))
    out = np.empty(B, np.float32)

    def step():
        _lib.check(lib.ssdc_ssd_loss(ctx.handle, d_true, _lib.F64, d_pred, 1, B, A, N_CLASSES, 3, 0, 1.0, _lib.ptr(out)))
    for _ in range(3):
        step()
    steps = 10
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps          # (the call returns the per-image losses to the host: wall clock)
    nbytes = B * A * W * 12
    ctx.dev_free(d_true)
    ctx.dev_free(d_pred)
    return {'images_per_s': B / dt, 'ms_per_step': dt * 1e3, 'read_GBps': nbytes / dt / 1e9, 'frac_of_hbm_peak': nbytes / dt / 1e9 / peak,
            'mean_loss': float(out.mean()), 'note': 'wall clock of ssdc_ssd_loss (device-resident float64 y_true + float32 y_pred, 3.46 MB/image read; '
                                                    'one host round trip for the mining count, (B,) floats copied back)'}


def voc_numbers(ctx):
    """`Evaluator.match_predictions` on a synthetic dataset (1000 images, 20 classes, 60 detections per image):
    the public call with the reference's list-of-tuples structure, the device part alone, and the oracle's
    per-prediction Python loop on the same data."""
    from oracle import cases
    from oracle import voc_eval_oracle as voc
    from jpeg_detection_resnet_ssd_b200.eval_utils.average_precision_evaluator import Evaluator

    case = dict(name='bench', seed=2024, n_images=1000, n_classes=20, dets_per_image=60, kwargs=dict())
    inp = cases.build_voc_input(case)

    class DS(object):
        pass
    ds = DS()
    ds.labels, ds.image_ids, ds.eval_neutral = inp['labels'], inp['image_ids'], None
    ev = Evaluator(model=None, n_classes=20, data_generator=ds)
    ev.set_predictions(inp['prediction_results'])
    n_pred = sum(len(p) for p in inp['prediction_results'])
    ev.match_predictions(sorting_algorithm='mergesort')
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        ev.match_predictions(sorting_algorithm='mergesort')
    t_api = (time.perf_counter() - t0) / reps
    ctx.profile_enable(True)
    ev.match_predictions(sorting_algorithm='mergesort')
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    dev_ms = prof['thin'][0]
    t0 = time.perf_counter()
    voc.match_predictions(inp['prediction_results'], inp['labels'], inp['image_ids'], None, 20, sorting_algorithm='mergesort')
    t_cpu = time.perf_counter() - t0
    return {'predictions': n_pred, 'api_predictions_per_s': n_pred / t_api, 'api_ms': t_api * 1e3,
            'device_kernels_ms': dev_ms, 'device_predictions_per_s': n_pred / (dev_ms / 1e3),
            'cpu_port_predictions_per_s': n_pred / t_cpu, 'cpu_port_s': t_cpu,
            'note': 'api = Evaluator.match_predictions incl. flattening the list-of-tuples input in Python and the copies'}


if __name__ == '__main__':
    main()
