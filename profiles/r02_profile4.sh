#!/bin/bash
# ncu captures of the final D1 + sweep at the dense workloads (SSD300 bg-bias 6, SSD512 conf 0.001)
set -u
OUT=gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extra --sustain-ms 0 --no-pipeline 1"
cap() {  # name, args...
  local name=$1; shift
  timeout 200 $B "$@" > $OUT/plain_$name.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'decode_filter_tma|sweep_kernel' -s 16 -c 2 -o $OUT/r02b_$name $B "$@" > $OUT/ncu_$name.log 2>&1
  echo "$name rc=$? $(tail -1 $OUT/ncu_$name.log)"
}
cap decode_c2_dense --config 2 --bg-bias 6
cap decode_c3 --config 3
