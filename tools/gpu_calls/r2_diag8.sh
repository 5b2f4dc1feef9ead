#!/bin/bash
out=gpurun_out/r2h; mkdir -p $out
step() { local sname=$1; shift; echo "== $sname"; ( timeout "$TMO" "$@" ) > $out/$sname.log 2>&1; echo "$sname exit $?" | tee -a $out/summary.txt; tail -n 3 $out/$sname.log; }
: > $out/summary.txt
TMO=600 step conv_tests python -m pytest tests/test_conv_gpu.py -q --timeout 120
TMO=400 step conv_shapes python tools/conv_shapes.py
TMO=120 step dw_bench python tools/dw_bench.py
TMO=120 step dw_bench2 python tools/dw_bench.py 32 64 64 256
TMO=900 step gpu_tests python -m pytest tests -x -q -m gpu --timeout 300
TMO=600 step bench python bench.py --no-others --no-sample --no-cpu-baseline --steps 10
cat $out/summary.txt
