#!/bin/bash
# Final evidence of round 2 on ONE B200: whole GPU test suite, one bench line per BASELINE configuration (+ reference arm),
# ncu capture of D1 + sweep at the headline workload, launch lists of configs[2] and configs[1].
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $O/r02_final_tests.log
bash profiles/r02_bench_all.sh
bash profiles/r02_profile2.sh
