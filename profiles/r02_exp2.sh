#!/bin/bash
# experiment 2: whole GPU suite after the lane change; encode lanes 3 vs 4 (long loops); L2 evict_first hint on D1's bulk copies
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/exp2_tests.log
rm -f $O/exp2.log
for l in 3 4; do
  timeout 200 python bench.py --config 1 --no-extra --no-cpu --no-e2e --steps 2000 --warmup 200 --enc-lanes $l 2> $O/exp2_enc_l$l.err | python -c "import json,sys; d=json.load(sys.stdin); print('enc B32 lanes $l', d['ms_per_step'], d['sustained'])" >> $O/exp2.log 2>&1
done
for e in 0 1; do
  timeout 200 python bench.py --config 2 --no-extra --no-cpu --no-e2e --steps 100 --warmup 10 --d1-evict-first $e 2> $O/exp2_c2_e$e.err | python -c "import json,sys; d=json.load(sys.stdin); print('c2 evict $e', d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['kernel_ms_per_step'])" >> $O/exp2.log 2>&1
  timeout 200 python bench.py --config 2 --bg-bias 6 --no-extra --no-cpu --no-e2e --steps 100 --warmup 10 --d1-evict-first $e 2> $O/exp2_c2d_e$e.err | python -c "import json,sys; d=json.load(sys.stdin); print('c2 dense evict $e', d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['kernel_ms_per_step'])" >> $O/exp2.log 2>&1
  timeout 200 python bench.py --config 3 --no-extra --no-cpu --no-e2e --steps 50 --warmup 10 --d1-evict-first $e 2> $O/exp2_c3_e$e.err | python -c "import json,sys; d=json.load(sys.stdin); print('c3 evict $e', d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['kernel_ms_per_step'])" >> $O/exp2.log 2>&1
done
cat $O/exp2_tests.log $O/exp2.log
