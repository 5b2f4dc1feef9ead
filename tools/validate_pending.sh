#!/bin/bash
# One gpurun call that runs everything written after the round-1 GPU budget was spent (NOTES.md, "Code that exists but has not
# run on a B200 yet") plus the two measurements round 1 still owes.  Usage (about 25 GPU-minutes):
#   gpurun --timeout 2400 -- 'bash tools/validate_pending.sh'
# Every step has its own timeout and log under gpurun_out/pending/; a failing step does not stop the next one.
mkdir -p gpurun_out/pending
out=gpurun_out/pending
step() { name=$1; shift; echo "== $name"; ( timeout "$TMO" "$@" ) > $out/$name.log 2>&1; echo "$name exit $?" | tee -a $out/summary.txt; tail -n 3 $out/$name.log; }
: > $out/summary.txt
TMO=600 step gpu_tests           python -m pytest tests -q -m gpu
# the guarded start-up selection itself: prints the child's report (which opt-in variants reproduce the default kernels and are faster)
TMO=400 step autotune            python -m cold_diffusion_models_b200.tuning
TMO=300 step model_training      env COLDDIFF_MODEL_TRAINING=1 python -m pytest tests/test_model2_train_gpu.py -q
TMO=200 step wgrad_bias_fusion   env COLDDIFF_EXPERIMENTAL=1 python -m pytest tests/test_conv_gpu.py -q -k fused_bias
TMO=400 step bench               python bench.py
TMO=300 step bench_no_autotune   python bench.py --no-autotune
TMO=300 step eager_comparator    python bench.py --impl reference --reference-device cuda --steps 3 --warmup 2
# wgrad with the bias column sums folded in (one extra N = 32 MMA per tile) against the separate cd_colsum launches
TMO=200 step op_profile_default  python tools/op_profile.py
# one-launch weight repacks (csrc/repack.cu): same bench and op profile with the switch on; flip the default in engine.batched_repack
# when the late GPU test (gpu_tests above: test_batched_repack_matches_the_single_launches) is green and this bench is not slower
TMO=300 step bench_batched_repack      env COLDDIFF_BATCHED_REPACK=1 python bench.py --no-autotune
TMO=200 step op_profile_batched_repack env COLDDIFF_BATCHED_REPACK=1 python tools/op_profile.py
# shared-memory-staged per-(batch element, head) LinearAttention kernels (csrc/linattn_small.cu), alone and with the batched repacks
TMO=200 step op_profile_linattn_staged env COLDDIFF_LINATTN_STAGED=1 python tools/op_profile.py
TMO=300 step bench_all_switches        env COLDDIFF_LINATTN_STAGED=1 COLDDIFF_BATCHED_REPACK=1 python bench.py --no-autotune
# line-coalesced epilogue of the tcgen05 convolution (csrc/conv_epilogue.cuh): per-shape table rows / mode 1 / mode 2, then the bench
TMO=300 step conv_shapes_epilogue      python tools/conv_shapes_epilogue.py
TMO=300 step bench_staged_epilogue     env COLDDIFF_CONV_STAGED_EPILOGUE=1 python bench.py --no-autotune
TMO=300 step bench_everything_on       env COLDDIFF_CONV_STAGED_EPILOGUE=3 COLDDIFF_LINATTN_STAGED=1 COLDDIFF_BATCHED_REPACK=1 COLDDIFF_LAYERNORM_MULTI=4 COLDDIFF_CONV_SIMT_PRELOAD=1 python bench.py --no-autotune
# ncu: launch list with DRAM bytes of one optimizer step with the staged epilogue on every launch, and a full capture of six staged
# convolution launches (source-level: does the epilogue still dominate?).  Copy the summaries into profiles/ (see profiles/README.md)
TMO=400 step ncu_launches_staged env COLDDIFF_CONV_STAGED_EPILOGUE=2 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $out/launches_step_staged.csv python tools/one_step.py
TMO=400 step ncu_full_staged     env COLDDIFF_CONV_STAGED_EPILOGUE=2 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_tc_kernel -c 6 -o $out/conv_staged_full -f python tools/one_step.py
# FP16-operand probe: per-shape time of the kind::f16 instantiation against the TF32 kernel on the same values (NOTES.md, FP16 option)
TMO=300 step conv_f16_probe      python tools/conv_f16_probe.py
grep -h '"metric"' $out/bench.log $out/bench_no_autotune.log $out/eager_comparator.log $out/bench_batched_repack.log $out/bench_all_switches.log $out/bench_staged_epilogue.log $out/bench_everything_on.log > $out/bench_lines.json 2>/dev/null
cat $out/summary.txt
