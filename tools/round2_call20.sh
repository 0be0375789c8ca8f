#!/bin/bash
# round 2, GPU call 20 (1 GPU): gated conv_rs3 (tests + training A/B), residual ring 4 slots + 2 operand stages A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_training.py -m gpu -q > gpurun_out/c20_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c20_pytest.log
tail -5 gpurun_out/c20_pytest.log | cut -c1-300
rm -f gpurun_out/c20_ab.txt
for w in resnet_h_infer resnet_f_infer; do
for v in 5 4 5 4; do
  DREAMB200_RES_INPLACE_SLOTS=$v timeout 300 python bench.py --workload $w --steps 10 --layer-table gpurun_out/c20_layers_${w}_$v.json > gpurun_out/c20_bench_${w}_$v.json 2>> gpurun_out/c20_bench_err.log
  python -c "
import json; d=json.loads(open('gpurun_out/c20_bench_${w}_$v.json').read().strip().splitlines()[-1]); print('$w slots=$v', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'conv', round(d['roofline']['conv_stack']['ms_per_step'],2), d['clocks']['sm_mhz'], d['clocks']['reasons'])
t=json.load(open('gpurun_out/c20_layers_${w}_$v.json'))
for l in t['layers']:
    if ' T1 ' in l['layer'] and l['ms']>0.2: print('   %-56s %7.3f ms %7.1f TF x%d'%(l['layer'],l['ms'],l['tflops'],l['launches']))" >> gpurun_out/c20_ab.txt
done
done
for v in 3 7 3 7; do
  DREAMB200_RS3=$v timeout 300 python bench.py --workload vgg_q_train --steps 8 --layer-table gpurun_out/c20_layers_train_$v.json > gpurun_out/c20_bench_train_$v.json 2>> gpurun_out/c20_bench_err.log
  python -c "
import json; d=json.loads(open('gpurun_out/c20_bench_train_$v.json').read().strip().splitlines()[-1]); print('vgg_q_train RS3=$v', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'conv', round(d['roofline']['conv_stack']['ms_per_step'],2))
t=json.load(open('gpurun_out/c20_layers_train_$v.json'))
for l in t['layers']:
    if 'Cin64 Cout64' in l['layer']: print('   %-56s %7.3f ms %7.1f TF x%d'%(l['layer'],l['ms'],l['tflops'],l['launches']))" >> gpurun_out/c20_ab.txt
done
cat gpurun_out/c20_ab.txt
tail -3 gpurun_out/c20_bench_err.log | cut -c1-200
