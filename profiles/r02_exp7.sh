#!/bin/bash
# experiment 7: D1 with smaller tiles / fewer consumer warps per CTA (room for sweep CTAs beside D1?)
set -u
O=gpurun_out
rm -f $O/exp7.log
run() { # warps ctas
  timeout 200 python bench.py --config 2 --no-extra --no-cpu --no-e2e --steps 100 --warmup 10 --d1-warps $1 --d1-ctas $2 2> $O/exp7_w$1_c$2.err | python -c "import json,sys; d=json.load(sys.stdin); print('c2 warps $1 ctas $2', d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['parity_checked'] if 'parity_checked' in d else '')" >> $O/exp7.log 2>&1
}
run 8 3
run 6 3
run 6 4
run 4 4
run 4 5
timeout 300 python -m pytest tests/test_gpu_decode.py -m gpu -x -q -k "floor or sweep or golden" 2>&1 | tail -2 >> $O/exp7.log
cat $O/exp7.log
