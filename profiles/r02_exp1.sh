#!/bin/bash
# experiment 1: encode lanes, D1 CTAs per SM at small batches
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -m gpu -x -q 2>&1 | tail -3 > $O/exp1_tests.log
for l in 1 2 3; do
  timeout 200 python bench.py --config 1 --no-extra --no-cpu --no-e2e --steps 200 --warmup 10 --enc-lanes $l 2> $O/exp1_enc_b32_l$l.err | python -c "import json,sys; d=json.load(sys.stdin); print('enc B32 lanes $l', d['ms_per_step'], d['sustained'])" >> $O/exp1.log 2>&1
done
for l in 1 3; do
  timeout 200 python bench.py --config 1 --batch 1024 --no-extra --no-cpu --no-e2e --steps 30 --warmup 5 --enc-lanes $l 2> $O/exp1_enc_b1024_l$l.err | python -c "import json,sys; d=json.load(sys.stdin); print('enc B1024 lanes $l', d['ms_per_step'], d['sustained'])" >> $O/exp1.log 2>&1
done
for b in 128 1024; do
  for c in 2 3; do
    timeout 200 python bench.py --config 2 --batch $b --no-extra --no-cpu --no-e2e --steps 100 --warmup 10 --d1-ctas $c 2> $O/exp1_dec_b${b}_c$c.err | python -c "import json,sys; d=json.load(sys.stdin); print('dec B$b ctas $c', d['ms_per_step'], d['sustained'], d['roofline']['kernel_ms_per_step'])" >> $O/exp1.log 2>&1
  done
done
cat $O/exp1_tests.log $O/exp1.log
