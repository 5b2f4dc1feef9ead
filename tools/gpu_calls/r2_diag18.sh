#!/bin/bash
out=gpurun_out/r2u; mkdir -p $out
step() { local sname=$1; shift; echo "== $sname"; ( timeout "$TMO" "$@" ) > $out/$sname.log 2>&1; echo "$sname exit $?" | tee -a $out/summary.txt; tail -n 3 $out/$sname.log; }
: > $out/summary.txt
TMO=600 step wgrad_tests python -m pytest tests/test_conv_gpu.py tests/test_grads_gpu.py -x -q -m gpu --timeout 200 -k "wgrad or grad"
TMO=300 step wgrad_shapes python tools/wgrad_shapes.py 32 model-12k
tail -2 $out/wgrad_shapes.log
TMO=300 step op_profile python tools/op_profile.py
head -6 $out/op_profile.log
cat $out/summary.txt
