#!/bin/bash
# Round-2 multi-GPU evidence, run on ONE 8-GPU B200 box (gpurun --gpus 8):
#   1. tests/test_gpu_multidev.py: one process, one context spanning all 8 devices (SURVEY 8e "single process" mode)
#   2. configs[2] STRONG scaling (batch 1024 sharded over N ranks), N = 1, 2, 4, 8
#   3. configs[2] weak scaling at N = 8 (1024 images per GPU), with the e2e leg and the concurrent H2D ceiling
#   4. configs[4] at its stated size: encode -> decode_detections_fast round trip, batch 4096 over 8 GPUs
# Every line lands in gpurun_out/r02_scale_*.json (copied to profiles/ afterwards).
set -u
OUT=gpurun_out
run() {  # n, name, args...
  local n=$1 name=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 400 python bench.py --gpus 1 "$@" > $OUT/$name.json 2> $OUT/$name.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
        bench.py --gpus $n "$@" > $OUT/$name.json 2> $OUT/$name.err
  fi
  echo "$name rc=$? $(grep -o '"value": [0-9.]*' $OUT/$name.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' $OUT/$name.json | head -1)"
}
nvidia-smi -L | wc -l
timeout 400 python -m pytest tests/test_gpu_multidev.py -m gpu -q 2>&1 | tail -2 | tee $OUT/r02_scale_multidev_test.log
for n in 1 2 4 8; do
  extra="--no-cpu"; [ $n = 1 ] && extra=""
  run $n r02_scale_c2_strong_n$n --config 2 --scaling strong --steps 50 --warmup 5 --no-extra $extra
done
run 8 r02_scale_c2_weak_n8 --config 2 --steps 50 --warmup 5 --no-extra
run 8 r02_scale_c4_n8 --config 4 --steps 10 --warmup 3 --no-extra
run 8 r02_scale_c3_weak_n8 --config 3 --steps 20 --warmup 3 --no-extra --no-cpu
