#!/bin/bash
# experiment 4: D1 producer loads the next image's floor one tile ahead; decode tests; final lines of all configurations
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -3 > $O/exp4_tests.log
rm -f $O/exp4.log
for r in 1 2; do
timeout 200 python bench.py --config 2 --no-extra --no-cpu --no-e2e --steps 100 --warmup 10 2> $O/exp4_c2.err | python -c "import json,sys; d=json.load(sys.stdin); print('c2', d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['kernel_ms_per_step'])" >> $O/exp4.log 2>&1
done
timeout 200 python bench.py --config 3 --no-extra --no-cpu --no-e2e --steps 50 --warmup 10 2> $O/exp4_c3.err | python -c "import json,sys; d=json.load(sys.stdin); print('c3', d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['kernel_ms_per_step'])" >> $O/exp4.log 2>&1
cat $O/exp4_tests.log $O/exp4.log
bash profiles/r02_bench_all.sh
