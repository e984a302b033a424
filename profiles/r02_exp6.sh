#!/bin/bash
# experiment 6: D1 row-maximum prefilter (FMNMX3) + warp-cooperative class masks
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_properties.py tests/test_gpu_thin.py -m gpu -x -q 2>&1 | tail -3 > $O/exp6_tests.log
rm -f $O/exp6.log
timeout 200 python bench.py --config 2 --no-extra --no-cpu --no-e2e --steps 100 --warmup 10 2> $O/exp6_c2.err | python -c "import json,sys; d=json.load(sys.stdin); print('c2', d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['kernel_ms_per_step'])" >> $O/exp6.log 2>&1
timeout 200 python bench.py --config 2 --bg-bias 6 --no-extra --no-cpu --no-e2e --steps 100 --warmup 10 2> $O/exp6_c2d.err | python -c "import json,sys; d=json.load(sys.stdin); print('c2 dense', d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['kernel_ms_per_step'])" >> $O/exp6.log 2>&1
timeout 200 python bench.py --config 2 --bg-bias 10 --no-extra --no-cpu --no-e2e --steps 100 --warmup 10 2> $O/exp6_c2s.err | python -c "import json,sys; d=json.load(sys.stdin); print('c2 sparse', d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['kernel_ms_per_step'])" >> $O/exp6.log 2>&1
timeout 200 python bench.py --config 3 --no-extra --no-cpu --no-e2e --steps 50 --warmup 10 2> $O/exp6_c3.err | python -c "import json,sys; d=json.load(sys.stdin); print('c3', d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['kernel_ms_per_step'])" >> $O/exp6.log 2>&1
cat $O/exp6_tests.log $O/exp6.log
