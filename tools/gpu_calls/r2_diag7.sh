#!/bin/bash
out=gpurun_out/r2g; mkdir -p $out
step() { local sname=$1; shift; echo "== $sname"; ( timeout "$TMO" "$@" ) > $out/$sname.log 2>&1; echo "$sname exit $?" | tee -a $out/summary.txt; tail -n 3 $out/$sname.log; }
: > $out/summary.txt
TMO=600 step conv_tests python -m pytest tests/test_conv_gpu.py -q --timeout 120
TMO=300 step conv_shapes python tools/conv_shapes.py
NCU="ncu --profile-from-start off --set full --import-source on --clock-control none"
TMO=500 step ncu_halo env COLDDIFF_CONV_HALO=1 $NCU -k regex:conv_tc3 -c 3 -o $out/halo -f python tools/fwd_profile.py
ncu -i $out/halo.ncu-rep --page raw --csv > $out/halo_raw.csv 2>/dev/null
ncu -i $out/halo.ncu-rep --page source --csv > $out/halo_source.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r2g/halo_source.csv')))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
with open('gpurun_out/r2g/halo_source_k0.csv', 'w', newline='') as f:
    csv.writer(f).writerows(rows[:starts[min(1, len(starts) - 1)]])
PY
rm -f $out/halo_source.csv $out/halo.ncu-rep
cat $out/summary.txt
