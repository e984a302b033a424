#!/bin/bash
# Round-2 bench lines of all five BASELINE configurations on ONE B200 (+ the CPU reference arm of the headline).
# Outputs: gpurun_out/r02_bench_c{0..4}.json, r02_ref_c2.json (copied to profiles/ afterwards).
set -u
OUT=gpurun_out
for c in 2 0 1 3 4; do
  # (the small-batch configurations run long loops: a few hundred steps of 0.02 ms end before host and device clocks have settled)
  case $c in 0|1) sw="--steps 5000 --warmup 5000";; 4) sw="--steps 10 --warmup 3";; *) sw="--steps 20 --warmup 5";; esac
  timeout 600 python bench.py --config $c $sw > $OUT/r02_bench_c$c.json 2> $OUT/r02_bench_c$c.err
  echo "config $c rc=$? $(grep -o '"value": [0-9.]*' $OUT/r02_bench_c$c.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' $OUT/r02_bench_c$c.json | head -1)"
done
timeout 600 python bench.py --impl reference --config 2 --steps 20 --warmup 5 > $OUT/r02_ref_c2.json 2>&1
echo "reference arm rc=$?"
