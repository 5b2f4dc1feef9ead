#!/bin/bash
out=gpurun_out/r2p; mkdir -p $out
step() { local sname=$1; shift; echo "== $sname"; ( timeout "$TMO" "$@" ) > $out/$sname.log 2>&1; echo "$sname exit $?" | tee -a $out/summary.txt; tail -n 3 $out/$sname.log; }
: > $out/summary.txt
TMO=600 step conv_tests python -m pytest tests/test_conv_gpu.py -x -q -m gpu --timeout 120
if grep -q "conv_tests exit 0" $out/summary.txt; then
TMO=300 step conv_shapes python tools/conv_shapes.py
cat $out/conv_shapes.log


TMO=300 step op_profile python tools/op_profile.py
head -8 $out/op_profile.log
COLDDIFF_CONV_HALO=2 TMO=300 step op_profile_halo2 python tools/op_profile.py
head -8 $out/op_profile_halo2.log
TMO=600 step grads python -m pytest tests/test_unet_gpu.py -x -q -m gpu --timeout 300
fi
cat $out/summary.txt
