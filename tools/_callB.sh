mkdir -p gpurun_out
( DREAMB200_WGRAD_C64=0 timeout 300 python -m pytest tests/test_gpu_training.py -m gpu -q -k "resnet_training" 2>&1 | grep -E "AssertionError|passed|failed|worst" ) > gpurun_out/B_c64_0.log 2>&1
( DREAMB200_WGRAD_C64=1 timeout 300 python -m pytest tests/test_gpu_training.py -m gpu -q -k "resnet_training" 2>&1 | grep -E "AssertionError|passed|failed|worst" ) > gpurun_out/B_c64_1.log 2>&1
( timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/B_all.log 2>&1
cat gpurun_out/B_c64_0.log gpurun_out/B_c64_1.log gpurun_out/B_all.log
