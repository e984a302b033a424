#!/bin/bash
# experiment 5: small-batch encodes with fewer stream operations per call (template on the call's own stream, stream query
# instead of the ordering event); all encoder tests; lanes 4 vs 6; round trip (lane scratch grows at once)
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py tests/test_gpu_loss.py -m gpu -x -q 2>&1 | tail -3 > $O/exp5_tests.log
rm -f $O/exp5.log
for l in 1 4 6; do
  timeout 200 python bench.py --config 1 --no-extra --no-cpu --no-e2e --steps 5000 --warmup 5000 --enc-lanes $l 2> $O/exp5_enc_l$l.err | python -c "import json,sys; d=json.load(sys.stdin); print('enc B32 lanes $l', d['ms_per_step'], d['sustained'])" >> $O/exp5.log 2>&1
done
timeout 300 python bench.py --config 4 --no-extra --no-cpu --no-e2e --steps 10 --warmup 3 2> $O/exp5_c4.err | python -c "import json,sys; d=json.load(sys.stdin); print('c4', d['ms_per_step'], d['sustained']['ms_per_step'])" >> $O/exp5.log 2>&1
timeout 200 python bench.py --config 1 --batch 1024 --no-extra --no-cpu --no-e2e --steps 50 --warmup 5 2> $O/exp5_enc1024.err | python -c "import json,sys; d=json.load(sys.stdin); print('enc B1024', d['ms_per_step'], d['sustained']['ms_per_step'])" >> $O/exp5.log 2>&1
cat $O/exp5_tests.log $O/exp5.log
