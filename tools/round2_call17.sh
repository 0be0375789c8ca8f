#!/bin/bash
# round 2, GPU call 17 (1 GPU): TMA-staged ReLU gate in conv_rs2 + per-map tile height in wgrad3x3_pair (tests + A/B)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_kernel_variants.py -m gpu -q -x > gpurun_out/c17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c17_pytest.log
timeout 600 python tools/wgrad_pair_check.py > gpurun_out/c17_wgrad_pair.txt 2>&1
DREAMB200_WGRAD_TH=16 timeout 600 python tools/wgrad_pair_check.py b256_256_100 b512_512_50 b512_512_25 > gpurun_out/c17_wgrad_pair_th16.txt 2>&1
rm -f gpurun_out/c17_ab.txt
for v in base new base new; do
  if [ $v = base ]; then export DREAMB200_LIB=$PWD/variants/base16.so; else unset DREAMB200_LIB; fi
  timeout 300 python bench.py --workload vgg_q_train --steps 8 --layer-table gpurun_out/c17_layers_train_$v.json > gpurun_out/c17_bench_train_$v.json 2>> gpurun_out/c17_bench_err.log
  python -c "
import json; d=json.loads(open('gpurun_out/c17_bench_train_$v.json').read().strip().splitlines()[-1]); print('vgg_q_train $v', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'conv', round(d['roofline']['conv_stack']['ms_per_step'],2))" >> gpurun_out/c17_ab.txt
done
unset DREAMB200_LIB
tail -6 gpurun_out/c17_pytest.log | cut -c1-300; cat gpurun_out/c17_ab.txt; tail -12 gpurun_out/c17_wgrad_pair.txt | cut -c1-400; tail -4 gpurun_out/c17_wgrad_pair_th16.txt | cut -c1-400
tail -3 gpurun_out/c17_bench_err.log | cut -c1-200
