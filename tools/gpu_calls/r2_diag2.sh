#!/bin/bash
# round-2 second GPU call: driver's test command, determinism per layer, in-situ profile, bench, ncu full captures of the hot kernels
out=gpurun_out/r2b; mkdir -p $out
step() { name=$1; shift; echo "== $name"; ( timeout "$TMO" "$@" ) > $out/$name.log 2>&1; echo "$name exit $?" | tee -a $out/summary.txt; tail -n 4 $out/$name.log; }
: > $out/summary.txt
TMO=900 step gpu_tests python -m pytest tests -x -q -m gpu --timeout 300
TMO=200 step det_small python tools/determinism_layers.py small
TMO=300 step det_full python tools/determinism_layers.py full
TMO=300 step det_full_train python tools/determinism_layers.py full train
TMO=300 step op_profile python tools/op_profile.py
TMO=300 step conv_shapes python tools/conv_shapes.py
TMO=300 step wgrad_shapes python tools/wgrad_shapes.py
TMO=500 step bench python bench.py --no-autotune --steps 10 --warmup 3
NCU="ncu --profile-from-start off --set full --import-source on --clock-control none"
TMO=400 step ncu_conv_fwd $NCU -k regex:conv_tc -c 14 -o $out/conv_fwd_full -f python tools/fwd_profile.py
TMO=300 step ncu_dw_ln $NCU -k regex:'dwconv7_pipe|layernorm' -c 6 -o $out/dw_ln_full -f python tools/fwd_profile.py
TMO=300 step ncu_attn $NCU -k regex:'ctx_partial|ctx_merge|weff' -c 4 -o $out/attn_fwd_full -f python tools/fwd_profile.py
TMO=500 step ncu_bwd $NCU -k regex:'attn_bwd|dwconv7_wgrad|layernorm_bwd|colsum' -c 12 -o $out/bwd_misc_full -f python tools/one_step.py
TMO=500 step ncu_wgrad $NCU -k regex:'wgrad_tc' -c 8 -o $out/wgrad_full -f python tools/one_step.py
cat $out/summary.txt
