mkdir -p gpurun_out
DREAMB200_RS_RESIDENT_WIDE=0 timeout 200 python bench.py --no-cpu-baseline --steps 10 --layer-table gpurun_out/C_layers_rw0.json > gpurun_out/C_bench_rw0.json 2> gpurun_out/C_err0.log
DREAMB200_RS_RESIDENT_WIDE=1 timeout 200 python bench.py --no-cpu-baseline --steps 10 --layer-table gpurun_out/C_layers_rw1.json > gpurun_out/C_bench_rw1.json 2> gpurun_out/C_err1.log
( timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_training.py -m gpu -q 2>&1 | tail -5 ) > gpurun_out/C_tests.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_rs --launch-skip 2 -c 1 -f -o gpurun_out/C_rs64 python tools/ncu_targets.py rs64 > gpurun_out/C_ncu.log 2>&1
cat gpurun_out/C_tests.log; cut -c1-200 gpurun_out/C_bench_rw0.json; cut -c1-200 gpurun_out/C_bench_rw1.json; tail -3 gpurun_out/C_ncu.log
