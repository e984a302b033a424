#!/bin/bash
# experiment 3: encode lanes 4 vs 6; L2 evict_first hints on the encoder's template stores (B = 1024, round trip) and the loss
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_loss.py tests/test_gpu_encode.py -m gpu -x -q -k "loss or lanes or back_to_back or template or golden" 2>&1 | tail -3 > $O/exp3_tests.log
rm -f $O/exp3.log
for l in 4 6; do
  timeout 200 python bench.py --config 1 --no-extra --no-cpu --no-e2e --steps 2000 --warmup 200 --enc-lanes $l 2> $O/exp3_enc_l$l.err | python -c "import json,sys; d=json.load(sys.stdin); print('enc B32 lanes $l', d['ms_per_step'], d['sustained'])" >> $O/exp3.log 2>&1
done
for h in 1 0; do
  timeout 200 python bench.py --config 1 --batch 1024 --no-extra --no-cpu --no-e2e --steps 50 --warmup 5 --no-l2-hints $h 2> $O/exp3_enc1024_h$h.err | python -c "import json,sys; d=json.load(sys.stdin); print('enc B1024 no_hints $h', d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['kernel_ms_per_step'])" >> $O/exp3.log 2>&1
  timeout 300 python bench.py --config 4 --no-extra --no-cpu --no-e2e --steps 10 --warmup 3 --no-l2-hints $h 2> $O/exp3_c4_h$h.err | python -c "import json,sys; d=json.load(sys.stdin); print('c4 no_hints $h', d['ms_per_step'], d['sustained']['ms_per_step'], d['roofline']['kernel_ms_per_step'])" >> $O/exp3.log 2>&1
done
timeout 300 python - >> $O/exp3.log 2>&1 <<'P'
import bench, json
from jpeg_detection_resnet_ssd_b200 import _lib
ctx = _lib.get_context()
peak, _ = bench.load_peaks()
print('loss', json.dumps(bench.loss_numbers(ctx, _lib, peak)))
P
cat $O/exp3_tests.log $O/exp3.log
