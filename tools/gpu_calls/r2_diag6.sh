#!/bin/bash
out=gpurun_out/r2f; mkdir -p $out
step() { local sname=$1; shift; echo "== $sname"; ( timeout "$TMO" "$@" ) > $out/$sname.log 2>&1; echo "$sname exit $?" | tee -a $out/summary.txt; tail -n 3 $out/$sname.log; }
: > $out/summary.txt
TMO=60 step fma_rate tools/micro/fma_rate.bin
TMO=600 step conv_tests python -m pytest tests/test_conv_gpu.py -q --timeout 120
TMO=900 step gpu_tests python -m pytest tests -x -q -m gpu --timeout 300
TMO=300 step conv_shapes python tools/conv_shapes.py
TMO=300 step op_profile python tools/op_profile.py
TMO=600 step bench python bench.py --no-others --no-sample --steps 10
TMO=600 step bench_halo env COLDDIFF_CONV_HALO=1 python bench.py --no-others --no-sample --no-cpu-baseline --steps 10
TMO=900 step gpu_tests_halo env COLDDIFF_CONV_HALO=1 python -m pytest tests -x -q -m gpu --timeout 300
cat $out/summary.txt
