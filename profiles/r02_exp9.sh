#!/bin/bash
# persistent host copy pool for pageable decode inputs: whole suite + e2e of configs[0] and configs[2]
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee $O/exp9_tests.log
for c in 0 2; do
  case $c in 0) sw="--steps 5000 --warmup 5000";; *) sw="--steps 20 --warmup 5";; esac
  timeout 600 python bench.py --config $c $sw --no-extra > $O/exp9_c$c.json 2> $O/exp9_c$c.err
  python -c "import json; d=json.load(open('$O/exp9_c$c.json')); print($c, d['value'], d['ms_per_step'], d['e2e'])"
done
