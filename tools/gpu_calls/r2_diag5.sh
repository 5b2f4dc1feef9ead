#!/bin/bash
out=gpurun_out/r2e; mkdir -p $out
step() { local sname=$1; shift; echo "== $sname"; ( timeout "$TMO" "$@" ) > $out/$sname.log 2>&1; echo "$sname exit $?" | tee -a $out/summary.txt; tail -n 3 $out/$sname.log; }
: > $out/summary.txt
TMO=900 step gpu_tests python -m pytest tests -x -q -m gpu --timeout 300
TMO=300 step conv_shapes python tools/conv_shapes.py
TMO=300 step op_profile python tools/op_profile.py
TMO=600 step bench python bench.py --no-others --steps 10
TMO=600 step bench_2cta_narrow env COLDDIFF_2CTA_BN=192 python bench.py --no-others --no-sample --no-cpu-baseline --steps 10
TMO=600 step bench_2cta_128 env COLDDIFF_2CTA_BN=128 python bench.py --no-others --no-sample --no-cpu-baseline --steps 10
NCU="ncu --profile-from-start off --set full --import-source on --clock-control none"
cap() { local name=$1 regex=$2 cnt=$3 script=$4
  TMO=500 step ncu_$name $NCU -k regex:"$regex" -c $cnt -o $out/$name -f python $script
  ncu -i $out/$name.ncu-rep --page raw --csv > $out/${name}_raw.csv 2>/dev/null
}
cap conv_fwd 'conv_tc' 8 tools/fwd_profile.py
ncu -i $out/conv_fwd.ncu-rep --page source --csv > $out/conv_fwd_source.csv 2>/dev/null
python - <<'PY'
# keep only the first two kernels of the source page (the 128x128 N = 64 / N = 128 layers)
import csv
rows = list(csv.reader(open('gpurun_out/r2e/conv_fwd_source.csv')))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
with open('gpurun_out/r2e/conv_fwd_source_k01.csv', 'w', newline='') as f:
    csv.writer(f).writerows(rows[:starts[min(2, len(starts) - 1)]])
PY
rm -f $out/conv_fwd_source.csv $out/conv_fwd.ncu-rep
cap bwd_misc 'attn_bwd|dwconv7_wgrad|layernorm_bwd|colsum' 8 tools/one_step.py
rm -f $out/bwd_misc.ncu-rep
du -sh gpurun_out >> $out/summary.txt
cat $out/summary.txt
