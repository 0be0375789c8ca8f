#!/bin/bash
# round 2, GPU call 7 (1 GPU): banded peaks kernel, graph pipeline, fixed training / drop-in tests, resnet L2-blocking sweep
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_parity.py tests/test_gpu_training.py tests/test_gpu_dropin.py -m gpu -q > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c7_pytest.log
for mb in 0 40 120; do
  DREAMB200_L2_BLOCK_MB=$mb timeout 300 python bench.py --workload resnet_h_infer --steps 10 --layer-table gpurun_out/c7_layers_resnet_h_l2_$mb.json > gpurun_out/c7_bench_resnet_h_l2_$mb.json 2>> gpurun_out/c7_bench_err.log
done
for mb in 0 40; do
  DREAMB200_L2_BLOCK_MB=$mb timeout 300 python bench.py --workload resnet_f_infer --steps 10 --layer-table gpurun_out/c7_layers_resnet_f_l2_$mb.json > gpurun_out/c7_bench_resnet_f_l2_$mb.json 2>> gpurun_out/c7_bench_err.log
done
tail -12 gpurun_out/c7_pytest.log | cut -c1-300
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/c7_bench_resnet_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'conv', round(d['roofline']['conv_stack']['ms_per_step'],2), d['launch_mode'])
    except Exception as e: print(f,'ERR',e)
P
tail -3 gpurun_out/c7_bench_err.log | cut -c1-300
