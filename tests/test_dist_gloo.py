"""CPU tests of the multi-rank host logic (gloo, world_size 2): the path shards by batch with no
data-path collective; the only cross-rank traffic is the barrier and the max-over-ranks of the
device-timed duration that bench.py reports."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import bench
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
dist = bench.init_dist(world, int(os.environ['LOCAL_RANK']), backend='gloo')
ms = bench.max_over_ranks(10.0 + 5.0 * rank, dist)          # rank 1 is the slow one
dist.barrier()
value = bench.whole_job_throughput(world * 1024, 4, ms)          # weak scaling: every rank brings 1024 images
lo, hi = bench.shard_range(1000, world, rank)
if rank == 0:
    print(json.dumps({'ms': ms, 'value': value, 'shard': [lo, hi]}))
else:
    assert (lo, hi) == (500, 1000)
dist.destroy_process_group()
'''


def free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def torchrun(nproc, script_args, timeout=300):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(nproc),
           '--master-addr', '127.0.0.1', '--master-port', str(free_port())] + script_args
    env = dict(os.environ, OMP_NUM_THREADS='1')
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)


def test_max_over_ranks_and_sharding_world2(tmp_path):
    worker = tmp_path / 'worker.py'
    worker.write_text(WORKER % {'root': ROOT})
    r = torchrun(2, [str(worker)])
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith('{')][-1]
    out = json.loads(line)
    assert out['ms'] == 15.0                                   # max over ranks, not rank 0's own 10 ms
    assert abs(out['value'] - 2 * 1024 * 4 / 0.015) < 1e-6      # whole-job images/s over both ranks
    assert out['shard'] == [0, 500]


def test_shard_rule_matches_library():
    import bench
    for total in (0, 1, 7, 8, 1000, 1024):
        for n in (1, 2, 4, 8):
            spans = [bench.shard_range(total, n, i) for i in range(n)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) <= (total + n - 1) // n


def test_reference_arm_under_torchrun_world2():
    """`bench.py --impl reference` under torchrun: rank 0 alone runs and prints ONE json line, the
    other rank exits 0 without work."""
    r = torchrun(2, ['bench.py', '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '1',
                     '--bg-bias', '9.0'], timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    out = json.loads(lines[0])
    assert out['impl'] == 'reference' and out['unit'] == 'images/s' and out['value'] > 0
    assert out['cpu_baseline']['kind'] == 'port' and out['cpu_baseline']['cores'] >= 1
    assert out['e2e']['h2d_bytes_per_step'] == 0 and out['n_gpus'] == 2
    assert out['config']['config_index'] == 2 and out['config']['batch_per_gpu'] == 1024


def test_reference_arm_other_configs():
    """Every BASELINE configuration has a CPU arm (encoder and round trip included)."""
    for cfg in ('1', '4'):
        r = subprocess.run([sys.executable, 'bench.py', '--impl', 'reference', '--config', cfg, '--steps', '1', '--warmup', '1'],
                           capture_output=True, text=True, timeout=600, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-2000:]
        out = json.loads([l for l in r.stdout.splitlines() if l.startswith('{')][-1])
        assert out['impl'] == 'reference' and out['value'] > 0 and out['config']['config_index'] == int(cfg)
        assert out['dtype'] == 'f64'
