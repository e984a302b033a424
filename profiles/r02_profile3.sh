#!/bin/bash
# ncu capture of the encoder's kernels at B = 1024 (one lane, no pipelining)
set -u
OUT=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra --sustain-ms 0 --enc-lanes 1 --config 1 --batch 1024"
timeout 200 $B > $OUT/plain_enc1024.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'seed_kernel|pair_kernel|greedy_kernel|apply_list' -s 12 -c 4 -o $OUT/r02b_encode_b1024 $B > $OUT/ncu_enc1024.log 2>&1
echo "encode rc=$? $(tail -1 $OUT/ncu_enc1024.log)"
