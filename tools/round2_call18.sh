#!/bin/bash
# round 2, GPU call 18 (1 GPU): conv_rs3 (two output rows per accumulator row) bring-up: bit-identity tests, inference A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernel_variants.py -m gpu -q -x -k "two_row" > gpurun_out/c18_pytest_rs3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c18_pytest_rs3.log
tail -5 gpurun_out/c18_pytest_rs3.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_parity.py tests/test_gpu_training.py -m gpu -q > gpurun_out/c18_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c18_pytest.log
rm -f gpurun_out/c18_ab.txt
for v in 0 3 0 3; do
  DREAMB200_RS3=$v timeout 300 python bench.py --no-cpu-baseline --no-secondary --layer-table gpurun_out/c18_layers_rs3_$v.json > gpurun_out/c18_bench_rs3_$v.json 2>> gpurun_out/c18_bench_err.log
  python -c "
import json; d=json.loads(open('gpurun_out/c18_bench_rs3_$v.json').read().strip().splitlines()[-1]); print('vgg_q_infer RS3=$v', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'])
t=json.load(open('gpurun_out/c18_layers_rs3_$v.json'))
for l in t['layers']:
    if 'Cin64 Cout64' in l['layer']: print('   %-56s %7.3f ms %7.1f TF x%d'%(l['layer'],l['ms'],l['tflops'],l['launches']))" >> gpurun_out/c18_ab.txt
done
timeout 300 python bench.py --workload vgg_q_train --steps 8 --layer-table gpurun_out/c18_layers_train.json > gpurun_out/c18_bench_train.json 2>> gpurun_out/c18_bench_err.log
python -c "
import json; d=json.loads(open('gpurun_out/c18_bench_train.json').read().strip().splitlines()[-1]); print('vgg_q_train', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'conv', round(d['roofline']['conv_stack']['ms_per_step'],2))" >> gpurun_out/c18_ab.txt
tail -6 gpurun_out/c18_pytest.log | cut -c1-300; cat gpurun_out/c18_ab.txt
tail -3 gpurun_out/c18_bench_err.log | cut -c1-200
