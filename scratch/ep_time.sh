timeout 300 python -m pytest tests/test_gpu_encode.py -x -q -m gpu 2>&1 | tail -2
for n in 1 2 3; do echo -n "nblk=$n: "; SSDC_ENC_NBLK=$n timeout 60 python scratch/enc_prof.py 1024 20; done
timeout 60 python scratch/enc_prof.py 32 20
