#!/bin/bash
# round 2, GPU call 3 (1 GPU): peaks v2 (integer-pipe conversions) tests + timing + ncu, bench line with parity counts
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -k regex:"peaks_fused" -o gpurun_out/c3_peaks -f python tools/ncu_targets.py peaks > gpurun_out/c3_ncu_peaks.log 2>&1
timeout 300 python bench.py --layer-table gpurun_out/c3_layers_vgg_q_infer.json > gpurun_out/c3_bench_vgg_q_infer.json 2> gpurun_out/c3_bench_err.log
tail -4 gpurun_out/c3_pytest.log; cat gpurun_out/c3_bench_vgg_q_infer.json; tail -3 gpurun_out/c3_bench_err.log
