#!/bin/bash
out=gpurun_out/r2z; mkdir -p $out
step() { local sname=$1; shift; echo "== $sname"; ( timeout "$TMO" "$@" ) > $out/$sname.log 2>&1; echo "$sname exit $?" | tee -a $out/summary.txt; tail -n 3 $out/$sname.log; }
: > $out/summary.txt
TMO=900 step gpu_tests python -m pytest tests -x -q -m gpu --timeout 300
TMO=300 step smoke python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')"
TMO=900 step bench python bench.py
TMO=600 step bench_ref python bench.py --impl reference --steps 1 --warmup 0
cat $out/summary.txt
