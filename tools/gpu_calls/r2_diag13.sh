#!/bin/bash
out=gpurun_out/r2o; mkdir -p $out
step() { local sname=$1; shift; echo "== $sname"; ( timeout "$TMO" "$@" ) > $out/$sname.log 2>&1; echo "$sname exit $?" | tee -a $out/summary.txt; tail -n 3 $out/$sname.log; }
: > $out/summary.txt
export COLDDIFF_CONV_HALO=2
NCU="ncu --profile-from-start off --set full --import-source on --clock-control none"
TMO=500 step ncu_tc4 $NCU -k regex:'conv_tc4' -c 5 -o $out/tc4 -f python tools/fwd_profile.py
ncu -i $out/tc4.ncu-rep --page raw --csv > $out/tc4_raw.csv 2>/dev/null
for k in 1 2; do
ncu -i $out/tc4.ncu-rep --page source --csv --launch-skip $k --launch-count 1 > $out/tc4_source_k$k.csv 2>/dev/null
done
rm -f $out/tc4.ncu-rep
du -sh gpurun_out >> $out/summary.txt
cat $out/summary.txt
