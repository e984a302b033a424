#!/bin/bash
# new tests (tile shapes, six lanes) + configs[1] line with the longer e2e loop
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py tests/test_gpu_decode.py -m gpu -x -q -k "lanes or tile_shapes" 2>&1 | tail -3 | tee $O/exp8_tests.log
timeout 600 python bench.py --config 1 --steps 5000 --warmup 5000 > $O/r02_bench_c1.json 2> $O/r02_bench_c1.err
python -c "import json; d=json.load(open('$O/r02_bench_c1.json')); print(d['value'], d['ms_per_step'], d['sustained'], d['e2e'])"
timeout 600 python bench.py --config 0 --steps 5000 --warmup 5000 > $O/r02_bench_c0.json 2> $O/r02_bench_c0.err
python -c "import json; d=json.load(open('$O/r02_bench_c0.json')); print(d['value'], d['ms_per_step'], d['sustained'], d['e2e'])"
