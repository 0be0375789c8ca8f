#!/bin/bash
# round 2, GPU call 4 (1 GPU): peaks v3 (branch-free conversions), CUDA-graph inference, bench with graph replay + B=1 latency
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernel_variants.py -m gpu -x -q > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -k regex:"peaks_fused" -o gpurun_out/c4_peaks -f python tools/ncu_targets.py peaks > gpurun_out/c4_ncu_peaks.log 2>&1
timeout 300 python bench.py --layer-table gpurun_out/c4_layers_vgg_q_infer.json > gpurun_out/c4_bench_vgg_q_infer.json 2> gpurun_out/c4_bench_err.log
timeout 300 python bench.py --no-graph --no-cpu-baseline > gpurun_out/c4_bench_vgg_q_infer_eager.json 2>> gpurun_out/c4_bench_err.log
tail -4 gpurun_out/c4_pytest.log; cat gpurun_out/c4_bench_vgg_q_infer.json; cat gpurun_out/c4_bench_vgg_q_infer_eager.json | cut -c1-400; tail -3 gpurun_out/c4_bench_err.log
