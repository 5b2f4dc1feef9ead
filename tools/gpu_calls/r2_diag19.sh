#!/bin/bash
out=gpurun_out/r2v; mkdir -p $out
step() { local sname=$1; shift; echo "== $sname"; ( timeout "$TMO" "$@" ) > $out/$sname.log 2>&1; echo "$sname exit $?" | tee -a $out/summary.txt; tail -n 4 $out/$sname.log; }
: > $out/summary.txt
CS="compute-sanitizer --tool memcheck --error-exitcode 3"
TMO=900 step memcheck_conv $CS python -m pytest tests/test_conv_gpu.py -x -q -m gpu --timeout 600 -k "wide_halo or (stride1 and tc4)"
TMO=600 step memcheck_dw_snow $CS python -m pytest tests/test_helpers_and_variants_gpu.py tests/test_more_packages_gpu.py -x -q -m gpu --timeout 500 -k "depthwise or snow"
TMO=900 step memcheck_step $CS python tools/one_step.py
cat $out/summary.txt
