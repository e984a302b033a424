#!/usr/bin/env python
"""bench.py - headline benchmark of the SSD box codec hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): decoded images/sec (decode + NMS, SSD300 VOC).  A "step" is one pass of
`decode_detections` (conf 0.01, IoU 0.45, top_k 200) over one batch of synthetic SSD300 raw
predictions (8732 anchors x 33 floats per image, float32).  Per-GPU batch is fixed (weak scaling,
batch shards are independent; no data-path collective).

  value      images/s with the batch already resident in HBM (device-timed, CUDA events on the
             library's own stream, max over ranks)
  e2e        the same through the public Python call `decode_detections(y_pred, ...)` with the batch
             in pinned HOST memory: H2D copy + kernels + D2H of the results, every step
  roofline   D1 (decode + filter) kernel: algorithmic bytes = B * A * (C+12) * 4 per launch over its
             event-timed duration, against the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the numpy oracle (a restatement of the reference codec with the same cost structure)
             on one host core, bounded sample of the same workload

`--impl reference` times the reference's CPU implementation (the numpy oracle port; the reference
itself is Python and does not travel to the GPU box) on all host cores.
"""
from __future__ import division

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

A_SSD300 = 8732
N_CLASSES = 21
CONF, IOU, TOPK = 0.01, 0.45, 200
METRIC = 'decoded images/sec (decode+NMS, SSD300 VOC)'
UNIT = 'images/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=1024, help='images per GPU per step')
    ap.add_argument('--bg-bias', type=float, default=8.0, help='background logit bias: candidate density')
    ap.add_argument('--unique', type=int, default=256, help='distinct synthetic images (tiled to the batch)')
    ap.add_argument('--no-extra', action='store_true', help='skip the secondary (encode, density sweep) numbers')
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def make_workload(batch, unique, bg_bias, seed, pinned):
    """Synthetic SSD300 y_pred (batch, 8732, 33) float32.  `unique` distinct images tiled."""
    import synth
    from jpeg_detection_resnet_ssd_b200 import pinned_empty
    from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
    enc = synth.make_encoder(SSDInputEncoder, 'ssd300')
    anchors = synth.anchors_of(enc)
    unique = min(unique, batch)
    base = synth.synth_y_pred(anchors, enc.variances, N_CLASSES, unique, seed, bg_bias=bg_bias, hot=40)
    y = pinned_empty((batch, A_SSD300, N_CLASSES + 12), np.float32) if pinned else np.empty((batch, A_SSD300, N_CLASSES + 12), np.float32)
    for i in range(0, batch, unique):
        n = min(unique, batch - i)
        y[i:i + n] = base[:n]
    cands = float((base[:, :, 1:N_CLASSES] > CONF).sum()) / unique
    return y, cands, enc


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
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
            self._stop_evt.wait(0.0003)          # (the timed region of the default run is only ~5 ms long)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': len(self.samples)}


def decode_params(_lib):
    p = _lib.DecodeParams()
    p.mode = _lib.MODE_PER_CLASS
    p.input_coords = _lib.COORDS['centroids']
    p.normalize = 1
    p.border_pixels = _lib.BORDER['half']
    p.top_k = TOPK
    p.nms_cap = 0
    p.log_wh = 1
    p.do_nms = 1
    p.conf_thresh = CONF
    p.iou_thresh = IOU
    p.img_h = 300.0
    p.img_w = 300.0
    return p


def oracle_decode_chunk(args):
    """Worker of the CPU arms: decodes a slice of images with the numpy oracle."""
    y, = args
    from oracle import ssd_codec_oracle as orc
    out = orc.decode_detections(y, CONF, IOU, TOPK, 'centroids', True, 300, 300)
    return sum(int(np.size(o) // 6) for o in out)


def cpu_baseline_one_core(y, budget_s=12.0, max_images=128):
    """The oracle on ONE host core over a bounded sample sized for ~budget_s of CPU work."""
    from oracle import ssd_codec_oracle as orc
    n0 = min(4, y.shape[0])
    t = time.perf_counter()
    orc.decode_detections(y[:n0], CONF, IOU, TOPK, 'centroids', True, 300, 300)
    per = (time.perf_counter() - t) / n0
    n = int(max(n0, min(max_images, y.shape[0], budget_s / max(per, 1e-6))))
    t = time.perf_counter()
    orc.decode_detections(y[:n], CONF, IOU, TOPK, 'centroids', True, 300, 300)
    dt = time.perf_counter() - t
    return {'value': n / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
            'sample': 'numpy oracle (restatement of the reference decode_detections), first %d images of the '
                      'workload, 1 thread, %.2f s' % (n, dt)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (numpy oracle port, the
    reference being Python that cannot travel to the GPU box) on all host cores."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_step = max(cores, min(64, 4 * cores))
    y, cands, _ = make_workload(per_step, min(args.unique, per_step), args.bg_bias, seed=1234, pinned=False)
    chunks = [(np.ascontiguousarray(c),) for c in np.array_split(y, cores) if c.shape[0]]
    with mp.get_context('fork').Pool(cores) as pool:
        for _ in range(max(1, min(args.warmup, 2))):
            pool.map(oracle_decode_chunk, chunks)
        t = time.perf_counter()
        for _ in range(args.steps):
            pool.map(oracle_decode_chunk, chunks)
        dt = time.perf_counter() - t
    value = per_step * args.steps / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, per_step, cands, note='bounded sample: %d images per step' % per_step),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': 'numpy oracle over %d processes, %d images per step' % (cores, per_step)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, batch, cands, note=None):
    cfg = {'workload': 'SSD300 decode+NMS (configs[2]): decode_detections conf %.2f / IoU %.2f / top_k %d on synthetic '
                       'y_pred (8732 anchors, 21 VOC classes, float32)' % (CONF, IOU, TOPK),
           'batch_per_gpu': batch, 'anchors': A_SSD300, 'classes': N_CLASSES,
           'candidates_per_image': round(cands, 1), 'bg_bias': args.bg_bias,
           'unique_images': min(args.unique, batch),
           'l2': ('input per step (%.0f MB) exceeds the 126 MB L2; no flush needed' % (batch * A_SSD300 * 33 * 4 / 1e6))
                 if batch * A_SSD300 * 33 * 4 > 126e6 else 'n/a (host arm)',
           'parallelism': 'batch shards, one process per GPU, no collective'}
    if note:
        cfg['note'] = note
    return cfg


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
    """Device-timed durations are combined as the MAX over ranks."""
    if dist is None:
        return float(x)
    import torch
    dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_throughput(world, batch_per_rank, steps, max_ms):
    """images/s of the whole job: every rank processed batch_per_rank * steps images (weak scaling)."""
    return world * batch_per_rank * steps / (max_ms / 1e3)


def shard_range(total, n, i):
    """Contiguous batch shards of ceil(total / n) images (same rule as libssdcodec's contexts)."""
    per = (total + n - 1) // n
    return min(total, per * i), min(total, per * (i + 1))


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    dist = init_dist(world, local_rank)

    from __graft_entry__ import build
    if rank == 0:
        build()
    if dist is not None:
        dist.barrier()

    from jpeg_detection_resnet_ssd_b200 import _lib, Context, set_context
    from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_output_decoder import decode_detections
    ctx = Context([local_rank])
    set_context(ctx)
    lib = ctx.lib

    B = args.batch
    y, cands, enc = make_workload(B, args.unique, args.bg_bias, seed=1234 + rank, pinned=True)
    d_y = ctx.dev_alloc(y.nbytes)
    ctx.h2d(d_y, y)
    p = decode_params(_lib)

    def step_device():
        _lib.check(lib.ssdc_decode_submit(ctx.handle, d_y, _lib.F32, 1, B, A_SSD300, N_CLASSES, _lib.C.byref(p)))

    def barrier():
        ctx.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- device-resident timing -------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    ctx.timer_start()
    for _ in range(args.steps):
        step_device()
    ms = ctx.timer_stop()
    launches = ctx.launch_count() - launches0
    barrier()
    clocks = sampler.finish()
    ms = max_over_ranks(ms, dist)
    value = whole_job_throughput(world, B, args.steps, ms)

    # sanity: the step really produced detections
    counts = np.zeros(B, np.int32)
    total = _lib.C.c_int64(0)
    rows = np.empty((B * TOPK, 6))
    idx = np.empty(B * TOPK, np.int32)
    _lib.check(lib.ssdc_decode_collect(ctx.handle, _lib.ptr(rows), B * TOPK, _lib.ptr(counts), _lib.ptr(idx), _lib.C.byref(total)))
    assert total.value > 0 and counts.max() <= TOPK

    # ---- per-kernel timing of the dominant kernel (separate pass, events around every launch) --
    ctx.profile_enable(True)
    prof_steps = 5
    for _ in range(prof_steps):
        step_device()
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    d1_ms = prof['decode_filter'][0] / max(prof['decode_filter'][1], 1)
    alg_bytes = B * A_SSD300 * (N_CLASSES + 12) * 4
    peak, peak_src = load_peaks()
    achieved = alg_bytes / (d1_ms / 1e3) / 1e9
    step_ms_prof = sum(v[0] for v in prof.values()) / prof_steps
    shares = {k: round(v[0] / prof_steps / step_ms_prof, 4) for k, v in prof.items() if v[1]}
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            tj = json.load(fh)['decode_filter_tma_kernel']
        if B == 1024 and abs(args.bg_bias - 8.0) < 1e-9:
            traffic = tj['dram_bytes_per_launch']
    except Exception:
        pass
    roofline = {'bound': 'hbm', 'kernel': 'decode_filter_tma_kernel<float,false> (D1)', 'achieved': achieved, 'peak': peak,
                'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                'kernel_ms': d1_ms, 'algorithmic_bytes_per_launch': alg_bytes,
                'kernel_ms_per_step': {k: round(v[0] / prof_steps, 4) for k, v in prof.items() if v[1]},
                'share_of_step': shares}

    # ---- end to end through the public API, batch in pinned host memory ---------------------
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        out = decode_detections(y, CONF, IOU, TOPK, 'centroids', True, 300, 300)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = decode_detections(y, CONF, IOU, TOPK, 'centroids', True, 300, 300)
    dt = time.perf_counter() - t0
    dt = max_over_ranks(dt, dist)
    n_rows = sum(o.shape[0] for o in out if o.size)
    e2e = {'value': world * B * e2e_steps / dt, 'unit': UNIT, 'h2d_bytes_per_step': int(y.nbytes),
           'd2h_bytes_per_step': int(n_rows * 52 + B * 4 + 8), 'steps': e2e_steps,
           'api': 'ssd_output_decoder.decode_detections(y_pred pinned host ndarray) -> list of ndarrays'}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, B, cands),
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline,
    }

    if rank == 0 and world == 1:
        line['cpu_baseline'] = cpu_baseline_one_core(y)
        if not args.no_extra:
            line['extra'] = extra_numbers(ctx, _lib, enc, peak)
    ctx.dev_free(d_y)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def extra_numbers(ctx, _lib, enc, peak):
    """Secondary numbers (not the headline): the encoder at BASELINE configs[1] (B=32) and at a large
    batch, device-timed, with the write kernel's roofline."""
    import synth
    lib = ctx.lib
    out = {}
    ctx2, h = enc._encoder()
    for B in (32, 1024):
        gt = synth.synth_ground_truth(300, 300, 20, B, seed=77)
        flat, offs = synth.flatten_ground_truth(gt)
        nbytes = B * A_SSD300 * 33 * 8
        d_out = ctx.dev_alloc(nbytes)
        for _ in range(3):
            _lib.check(lib.ssdc_encode(h, _lib.ptr(flat), _lib.ptr(offs), B, 1, d_out, None, None))
        ctx.synchronize()
        steps = 10
        ctx.timer_start()
        for _ in range(steps):
            _lib.check(lib.ssdc_encode(h, _lib.ptr(flat), _lib.ptr(offs), B, 1, d_out, None, None))
        ms = ctx.timer_stop()
        ctx.profile_enable(True)
        for _ in range(3):
            _lib.check(lib.ssdc_encode(h, _lib.ptr(flat), _lib.ptr(offs), B, 1, d_out, None, None))
        prof = ctx.profile_read()
        ctx.profile_enable(False)
        w_ms = prof['enc_write'][0] / max(prof['enc_write'][1], 1)
        ips = B * steps / (ms / 1e3)
        out['encode_b%d' % B] = {
            'images_per_s': ips, 'ms_per_step': ms / steps,
            # whole encode step against the HBM roofline of its algorithmic bytes (y_encoded written once)
            'step_GBps': nbytes / (ms / steps / 1e3) / 1e9, 'step_frac_of_hbm_peak': nbytes / (ms / steps / 1e3) / 1e9 / peak,
            # the y_encoded write-out kernel (template TMA stream) timed alone
            'write_kernel_ms': w_ms, 'write_kernel_GBps': nbytes / (w_ms / 1e3) / 1e9,
            'write_kernel_frac_of_hbm_peak': nbytes / (w_ms / 1e3) / 1e9 / peak,
            'kernel_ms': {k: round(v[0] / 3, 4) for k, v in prof.items() if v[1]},
            'note': 'kernel_ms from a serialised profiling pass; in the timed step the write-out stream overlaps the matching kernels',
        }
        ctx.dev_free(d_out)

    # BASELINE configs[4] in miniature on one GPU: encode -> decode_detections_fast round trip, the
    # float64 y_encoded never leaves the device (every positive confidence is exactly 1.0)
    B = 512
    gt = synth.synth_ground_truth(300, 300, 20, B, seed=78)
    flat, offs = synth.flatten_ground_truth(gt)
    d_enc = ctx.dev_alloc(B * A_SSD300 * 33 * 8)
    pf = _lib.DecodeParams()
    pf.mode, pf.input_coords, pf.normalize, pf.border_pixels = _lib.MODE_FAST, 0, 1, 0
    pf.top_k, pf.nms_cap, pf.log_wh, pf.do_nms = 0, 0, 1, 1
    pf.conf_thresh, pf.iou_thresh, pf.img_h, pf.img_w = 0.5, 0.45, 300.0, 300.0

    def roundtrip():
        _lib.check(lib.ssdc_encode(h, _lib.ptr(flat), _lib.ptr(offs), B, 1, d_enc, None, None))
        _lib.check(lib.ssdc_decode_submit(ctx.handle, d_enc, _lib.F64, 1, B, A_SSD300, N_CLASSES, _lib.C.byref(pf)))
    for _ in range(3):
        roundtrip()
    ctx.synchronize()
    steps = 5
    ctx.timer_start()
    for _ in range(steps):
        roundtrip()
    ms = ctx.timer_stop()
    counts = np.zeros(B, np.int32)
    total = _lib.C.c_int64(0)
    rc = lib.ssdc_decode_collect(ctx.handle, None, 0, _lib.ptr(counts), None, _lib.C.byref(total))
    n_gt = int(offs[-1])
    out['roundtrip_encode_decode_fast_b512'] = {
        'images_per_s': B * steps / (ms / 1e3), 'ms_per_step': ms / steps,
        'ground_truth_boxes': n_gt, 'decoded_boxes': int(total.value),
        'note': 'device-resident float64 y_encoded; top_k=all so rows are emitted at collect time (not timed)'}
    ctx.dev_free(d_enc)

    # BASELINE configs[3]: SSD512 layout (24564 anchors), conf 0.001, dense candidates, B = 512
    from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
    enc512 = synth.make_encoder(SSDInputEncoder, 'ssd512')
    a512 = synth.anchors_of(enc512)
    A = a512.shape[0]
    uniq = 32
    base = synth.synth_y_pred(a512, enc512.variances, N_CLASSES, uniq, 4321, bg_bias=6.0, hot=40)
    B = 512
    d_y = ctx.dev_alloc(B * A * 33 * 4)
    for i in range(0, B, uniq):
        _lib.check(lib.ssdc_memcpy_h2d(ctx.handle, 0, _lib.C.c_void_p(d_y.value + i * A * 33 * 4), _lib.ptr(base), base.nbytes))
    p5 = decode_params(_lib)
    p5.conf_thresh, p5.img_h, p5.img_w = 0.001, 512.0, 512.0
    for _ in range(3):
        _lib.check(lib.ssdc_decode_submit(ctx.handle, d_y, _lib.F32, 1, B, A, N_CLASSES, _lib.C.byref(p5)))
    ctx.synchronize()
    ctx.timer_start()
    for _ in range(steps):
        _lib.check(lib.ssdc_decode_submit(ctx.handle, d_y, _lib.F32, 1, B, A, N_CLASSES, _lib.C.byref(p5)))
    ms = ctx.timer_stop()
    ctx.profile_enable(True)
    for _ in range(3):
        _lib.check(lib.ssdc_decode_submit(ctx.handle, d_y, _lib.F32, 1, B, A, N_CLASSES, _lib.C.byref(p5)))
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    d1 = prof['decode_filter'][0] / max(prof['decode_filter'][1], 1)
    out['decode_ssd512_dense_b512'] = {
        'images_per_s': B * steps / (ms / 1e3), 'ms_per_step': ms / steps,
        'candidates_per_image': float((base[:, :, 1:N_CLASSES] > 0.001).sum()) / uniq,
        'decode_filter_ms': d1, 'decode_filter_frac_of_hbm_peak': B * A * 33 * 4 / (d1 / 1e3) / 1e9 / peak,
        'kernel_ms': {k: round(v[0] / 3, 4) for k, v in prof.items() if v[1]}}
    ctx.dev_free(d_y)

    # SSD loss on the device-resident y_encoded (float64) + a float32 prediction
    try:
        out['ssd_loss_b1024'] = loss_numbers(ctx, _lib, enc, peak)
    except Exception as exc:
        out['ssd_loss_b1024'] = {'error': repr(exc)[:200]}

    # the VOC matching core of the Evaluator (the consumer of the decoder's output)
    try:
        out['voc_match_predictions'] = voc_numbers(ctx)
    except Exception as exc:      # secondary number: never take the headline down
        out['voc_match_predictions'] = {'error': repr(exc)[:200]}
    return out


def loss_numbers(ctx, _lib, enc, peak):
    """`ssdc_ssd_loss` with both tensors resident on the device: y_true = the encoder's float64 output,
    y_pred = synthetic float32 predictions (64 unique images tiled)."""
    import synth
    lib = ctx.lib
    B, A, W = 1024, A_SSD300, 33
    ctx2, h = enc._encoder()
    gt = synth.synth_ground_truth(300, 300, 20, B, seed=79)
    flat, offs = synth.flatten_ground_truth(gt)
    d_true = ctx.dev_alloc(B * A * W * 8)
    _lib.check(lib.ssdc_encode(h, _lib.ptr(flat), _lib.ptr(offs), B, 1, d_true, None, None))
    uniq = 64
    base = synth.synth_y_pred(synth.anchors_of(enc), enc.variances, N_CLASSES, uniq, 99, bg_bias=3.0, hot=40)
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
