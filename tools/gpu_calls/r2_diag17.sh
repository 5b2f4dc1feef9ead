#!/bin/bash
out=gpurun_out/r2s; mkdir -p $out
step() { local sname=$1; shift; echo "== $sname"; ( timeout "$TMO" "$@" ) > $out/$sname.log 2>&1; echo "$sname exit $?" | tee -a $out/summary.txt; tail -n 3 $out/$sname.log; }
: > $out/summary.txt
TMO=900 step gpu_tests python -m pytest tests -x -q -m gpu --timeout 300
TMO=300 step op_profile python tools/op_profile.py
TMO=300 step conv_shapes python tools/conv_shapes.py
TMO=300 step wgrad_shapes python tools/wgrad_shapes.py 32 model-12k
TMO=900 step bench python bench.py --steps 10
TMO=600 step ncu_launches ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $out/launches_step.csv python tools/one_step.py
NCU="ncu --profile-from-start off --set full --clock-control none"
TMO=500 step ncu_conv $NCU -k regex:'conv_tc' -c 12 -o $out/conv -f python tools/fwd_profile.py
ncu -i $out/conv.ncu-rep --page raw --csv > $out/conv_raw.csv 2>/dev/null; rm -f $out/conv.ncu-rep
TMO=500 step ncu_dw $NCU -k regex:'dwconv7_tma|dwconv7_wgrad_tma|snow' -c 6 -o $out/dw -f python tools/one_step.py
ncu -i $out/dw.ncu-rep --page raw --csv > $out/dw_raw.csv 2>/dev/null; rm -f $out/dw.ncu-rep
du -sh gpurun_out >> $out/summary.txt
cat $out/summary.txt
