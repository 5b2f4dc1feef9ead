#!/bin/bash
out=gpurun_out/r2q; mkdir -p $out
step() { local sname=$1; shift; echo "== $sname"; ( timeout "$TMO" "$@" ) > $out/$sname.log 2>&1; echo "$sname exit $?" | tee -a $out/summary.txt; tail -n 3 $out/$sname.log; }
: > $out/summary.txt
TMO=300 step wgrad_shapes python tools/wgrad_shapes.py 32 model-12k
tail -3 $out/wgrad_shapes.log
TMO=900 step gpu_tests python -m pytest tests -x -q -m gpu --timeout 300
TMO=300 step op_profile python tools/op_profile.py
head -12 $out/op_profile.log
TMO=300 step conv_shapes python tools/conv_shapes.py
TMO=900 step bench python bench.py --steps 10
tail -1 $out/bench.log
cat $out/summary.txt
