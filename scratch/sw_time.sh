for cfg in "2048 8.0" "1024 6.0" "1024 12.0" "512 8.0" "3072 8.0"; do
  for t in 128 256; do echo -n "B,bias=$cfg threads=$t: "; SSDC_SWEEP_THREADS=$t timeout 200 python scratch/d1_prof.py $cfg 2>&1 | tail -1 | cut -c1-200; done
done
