set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || exit 1
python bench.py --steps 3 --warmup 3 --no-extra > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_e_launches_decode.csv python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/ncu_a.log 2>&1
python scratch/enc_prof.py 1024 3 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_e_launches_encode.csv python scratch/enc_prof.py 1024 3 > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"decode_filter_tma|sweep_kernel" -s 6 -c 2 -o gpurun_out/r01_e_decode_full -f python scratch/d1_prof.py > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"template_tma|pair_kernel|greedy_kernel|seed_kernel|apply_list" -s 10 -c 5 -o gpurun_out/r01_e_encode_full -f python scratch/enc_prof.py 1024 2 > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out | tail -12
python scratch/loss_prof.py || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_e_launches_loss.csv python scratch/loss_prof.py > gpurun_out/ncu_e.log 2>&1
