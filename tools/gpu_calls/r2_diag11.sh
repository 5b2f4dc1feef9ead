#!/bin/bash
out=gpurun_out/r2m; mkdir -p $out
step() { local sname=$1; shift; echo "== $sname"; ( timeout "$TMO" "$@" ) > $out/$sname.log 2>&1; echo "$sname exit $?" | tee -a $out/summary.txt; tail -n 3 $out/$sname.log; }
: > $out/summary.txt
TMO=120 step umma_rate tools/micro/umma_rate.bin
cat $out/umma_rate.log
TMO=300 step dw_test python -m pytest tests/test_helpers_and_variants_gpu.py -x -q -m gpu --timeout 300 -k "depthwise"
TMO=300 step dw_bench python tools/dw_bench.py
cat $out/dw_bench.log
TMO=300 step op_profile python tools/op_profile.py
head -12 $out/op_profile.log
cat $out/summary.txt
