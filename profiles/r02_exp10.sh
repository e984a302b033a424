#!/bin/bash
# experiment 10: D1 tile shapes at small batches (the per-rank batch of an 8-GPU strong-scaling run)
set -u
O=gpurun_out
rm -f $O/exp10.log
run() { # batch warps ctas
  timeout 200 python bench.py --config 2 --batch $1 --no-extra --no-cpu --no-e2e --steps 400 --warmup 50 --d1-warps $2 --d1-ctas $3 2> $O/exp10.err | python -c "import json,sys; d=json.load(sys.stdin); print('B=$1 warps $2 ctas $3', round(d['ms_per_step'],5), round(d['sustained']['ms_per_step'],5), d['roofline']['kernel_ms_per_step'])" >> $O/exp10.log 2>&1
}
for b in 128 256; do
  run $b 8 3
  run $b 6 4
  run $b 4 5
  run $b 4 6
done
cat $O/exp10.log
