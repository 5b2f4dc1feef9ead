timeout -s KILL 100 python tools/dw_bench.py
timeout -s KILL 100 python tools/dw_bench.py 32 16 16 512
timeout -s KILL 400 python -m pytest tests/test_unet_gpu.py tests/test_grads_gpu.py tests/test_trainer_gpu.py -q 2>&1 | tail -5
