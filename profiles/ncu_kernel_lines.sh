#!/bin/bash
# Top CUDA source lines (executed warp instructions, stall samples with their reasons) of ONE kernel of an ncu report.
#   bash profiles/ncu_kernel_lines.sh gpurun_out/X.ncu-rep <kernel substring> [N] > profiles/X_<kernel>_lines.txt
rep=$1; kern=$2; n=${3:-30}
tmp=$(mktemp)
ncu -i "$rep" --page source --print-source cuda,sass --csv 2>/dev/null > $tmp.all
python - "$tmp.all" "$kern" "$tmp.k" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
out, keep = [], False
for i, r in enumerate(rows):
    if r and r[0] == 'File Path':
        fn = rows[i + 1][1] if i + 1 < len(rows) and rows[i + 1] and rows[i + 1][0] == 'Function Name' else ''
        keep = sys.argv[2] in fn
    if keep:
        out.append(r)
csv.writer(open(sys.argv[3], 'w')).writerows(out)
PY
python "$(dirname $0)/ncu_lines.py" $tmp.k $n
rm -f $tmp $tmp.all $tmp.k
