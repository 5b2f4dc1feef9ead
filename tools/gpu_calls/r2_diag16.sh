#!/bin/bash
out=gpurun_out/r2r; mkdir -p $out
step() { local sname=$1; shift; echo "== $sname"; ( timeout "$TMO" "$@" ) > $out/$sname.log 2>&1; echo "$sname exit $?" | tee -a $out/summary.txt; tail -n 3 $out/$sname.log; }
: > $out/summary.txt
TMO=600 step conv_tests python -m pytest tests/test_conv_gpu.py tests/test_unet_gpu.py -x -q -m gpu --timeout 120
TMO=300 step conv_shapes python tools/conv_shapes.py
head -6 $out/conv_shapes.log | cut -c1-200; tail -1 $out/conv_shapes.log
TMO=300 step op_profile python tools/op_profile.py
head -8 $out/op_profile.log
cat $out/summary.txt
