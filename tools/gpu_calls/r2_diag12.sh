#!/bin/bash
out=gpurun_out/r2n; mkdir -p $out
step() { local sname=$1; shift; echo "== $sname"; ( timeout "$TMO" "$@" ) > $out/$sname.log 2>&1; echo "$sname exit $?" | tee -a $out/summary.txt; tail -n 3 $out/$sname.log; }
: > $out/summary.txt
TMO=400 step tc4_test python -m pytest tests/test_conv_gpu.py -x -q -m gpu --timeout 120 -k "wide_halo or tc4"
if grep -q "tc4_test exit 0" $out/summary.txt; then
TMO=300 step conv_shapes python tools/conv_shapes.py
cat $out/conv_shapes.log
COLDDIFF_CONV_HALO=2 TMO=300 step op_profile_halo2 python tools/op_profile.py
head -8 $out/op_profile_halo2.log
fi
TMO=200 step snow_test python -m pytest tests/test_more_packages_gpu.py -x -q -m gpu --timeout 120 -k "snow"
cat $out/summary.txt
