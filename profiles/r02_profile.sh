#!/bin/bash
# Round-2 ncu evidence (one B200): full captures of D1 + sweep at the headline, the SSD300 dense (bg-bias 6) and the SSD512
# dense (configs[3]) workloads, the encoder's kernels at configs[1], and launch lists.  Outputs in gpurun_out/.
set -u
OUT=gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extra --sustain-ms 0 --no-pipeline 1"
cap() {  # name, regex, args...
  local name=$1 rx=$2; shift 2
  timeout 200 $B "$@" > $OUT/plain_$name.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 16 -c 2 -o $OUT/r02_$name $B "$@" > $OUT/ncu_$name.log 2>&1
  echo "$name rc=$? $(tail -1 $OUT/ncu_$name.log)"
}
cap decode_c2 'decode_filter_tma|sweep_kernel' --config 2
cap decode_c2_dense 'decode_filter_tma|sweep_kernel' --config 2 --bg-bias 6
cap decode_c3 'decode_filter_tma|sweep_kernel' --config 3
timeout 200 $B --config 1 > $OUT/plain_enc.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'template_tma|seed_kernel|pair_kernel|greedy_kernel|apply_list' -s 15 -c 5 -o $OUT/r02_encode_c1 $B --config 1 > $OUT/ncu_enc.log 2>&1
echo "encode rc=$? $(tail -1 $OUT/ncu_enc.log)"
for c in 2 1 4; do
  timeout 200 $B --config $c > $OUT/plain_l$c.log 2>&1 && \
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 60 --csv --log-file $OUT/r02_launches_c$c.csv $B --config $c > $OUT/ncu_l$c.log 2>&1
  echo "launch list c$c rc=$?"
done
