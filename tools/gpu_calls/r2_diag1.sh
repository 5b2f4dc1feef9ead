#!/bin/bash
# round-2 first GPU call: full GPU suite (no -x, gated tests on), determinism per layer, initcheck, eager comparator
out=gpurun_out/r2a; mkdir -p $out
step() { name=$1; shift; echo "== $name"; ( timeout "$TMO" "$@" ) > $out/$name.log 2>&1; echo "$name exit $?" | tee -a $out/summary.txt; tail -n 4 $out/$name.log; }
: > $out/summary.txt
TMO=900 step gpu_tests env COLDDIFF_MODEL_TRAINING=1 COLDDIFF_EXPERIMENTAL=1 python -m pytest tests -q -m gpu -rA --timeout 300
TMO=200 step det_small python tools/determinism_layers.py small
TMO=300 step det_full python tools/determinism_layers.py full
TMO=200 step det_small_train python tools/determinism_layers.py small train
TMO=300 step det_full_train python tools/determinism_layers.py full train
TMO=300 step det_probe python tools/determinism_probe.py
TMO=500 step initcheck compute-sanitizer --tool initcheck --print-limit 40 python tools/determinism_layers.py small
TMO=400 step eager_comparator python bench.py --impl reference --reference-device cuda --steps 3 --warmup 2
cat $out/summary.txt
