for v in "X=1" "SSDC_D1_NULL=1" "SSDC_D1_NULL=1 SSDC_D1_CTAS=2" "SSDC_D1_NULL=1 SSDC_D1_CTAS=1" "SSDC_D1_CTAS=2"; do
  echo "== $v"; env $v timeout 120 python scratch/d1_prof.py 2>&1 | tail -1
done
