#!/bin/bash
# round-2 third GPU call: ncu full captures (raw pages exported on the box, reports kept only while gpurun_out stays small),
# the rewritten bench, the FFMA microbenchmark
out=gpurun_out/r2c; mkdir -p $out
step() { name=$1; shift; echo "== $name"; ( timeout "$TMO" "$@" ) > $out/$name.log 2>&1; echo "$name exit $?" | tee -a $out/summary.txt; tail -n 3 $out/$name.log; }
: > $out/summary.txt
TMO=60 step fma_rate tools/micro/fma_rate.bin
TMO=300 step op_profile python tools/op_profile.py
TMO=300 step conv_shapes python tools/conv_shapes.py
TMO=300 step wgrad_shapes python tools/wgrad_shapes.py
NCU="ncu --profile-from-start off --set full --import-source on --clock-control none"
cap() { name=$1; regex=$2; cnt=$3; script=$4
  TMO=500 step ncu_$name $NCU -k regex:"$regex" -c $cnt -o $out/$name -f python $script
  ncu -i $out/$name.ncu-rep --page raw --csv > $out/${name}_raw.csv 2>/dev/null
  ncu -i $out/$name.ncu-rep --page details --csv > $out/${name}_details.csv 2>/dev/null
}
cap conv_fwd 'conv_tc' 12 tools/fwd_profile.py
ncu -i $out/conv_fwd.ncu-rep --page source --csv --kernel-name regex:conv_tc_kernel --launch-skip 0 --launch-count 1 > $out/conv_fwd_source_k0.csv 2>/dev/null
cap dw_ln 'dwconv7_pipe|layernorm' 4 tools/fwd_profile.py
cap attn_fwd 'ctx_partial|ctx_merge|weff' 3 tools/fwd_profile.py
cap bwd_misc 'attn_bwd|dwconv7_wgrad|layernorm_bwd|colsum' 10 tools/one_step.py
cap wgrad 'wgrad_tc' 5 tools/one_step.py
# keep the reports only while everything fits comfortably under gpurun's 64 MiB limit
for f in $(ls -S $out/*.ncu-rep); do
  if [ $(du -sm gpurun_out | cut -f1) -gt 40 ]; then rm -f $f; echo "dropped $f" >> $out/summary.txt; fi
done
TMO=900 step bench python bench.py
du -sh gpurun_out >> $out/summary.txt
cat $out/summary.txt
