#!/bin/bash
# Round-2 (second half) ncu evidence after the L2 evict_first hint / floor prefetch / encode lanes: full capture of D1 + sweep at
# the headline workload, launch lists of configs[2] and configs[1].  Outputs in gpurun_out/.
set -u
OUT=gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extra --sustain-ms 0 --no-pipeline 1"
timeout 200 $B --config 2 > $OUT/plain_c2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'decode_filter_tma|sweep_kernel' -s 16 -c 2 -o $OUT/r02b_decode_c2 $B --config 2 > $OUT/ncu_c2.log 2>&1
echo "decode_c2 rc=$? $(tail -1 $OUT/ncu_c2.log)"
for c in 2 1; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 60 --csv --log-file $OUT/r02b_launches_c$c.csv $B --config $c > $OUT/ncu_l$c.log 2>&1
  echo "launch list c$c rc=$?"
done
