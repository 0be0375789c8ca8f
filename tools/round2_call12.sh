#!/bin/bash
# round 2, GPU call 12 (1 GPU): running column sums in the 64-channel training epilogue (tests + A/B), ncu of the resnet expansion layer
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_multistage.py tests/test_gpu_kernel_variants.py -m gpu -q > gpurun_out/c12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c12_pytest.log
rm -f gpurun_out/c12_ab.txt
for v in base new base new; do
  if [ $v = base ]; then export DREAMB200_LIB=$PWD/variants/res_tma.so; else unset DREAMB200_LIB; fi
  timeout 300 python bench.py --workload vgg_q_train --steps 8 --layer-table gpurun_out/c12_layers_train_$v.json > gpurun_out/c12_bench_train_$v.json 2>> gpurun_out/c12_bench_err.log
  python -c "
import json; d=json.loads(open('gpurun_out/c12_bench_train_$v.json').read().strip().splitlines()[-1]); print('vgg_q_train $v', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'conv', round(d['roofline']['conv_stack']['ms_per_step'],2))" >> gpurun_out/c12_ab.txt
done
unset DREAMB200_LIB
timeout 300 ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -k regex:"conv_tc2" -o gpurun_out/c12_expand -f python tools/ncu_targets.py expand > gpurun_out/c12_ncu_expand.log 2>&1
tail -6 gpurun_out/c12_pytest.log | cut -c1-300; cat gpurun_out/c12_ab.txt
python - <<'P'
import json
for v in ('base','new'):
    d=json.load(open('gpurun_out/c12_layers_train_%s.json'%v))
    print(v, sum(l['ms'] for l in d['layers']))
    for l in d['layers']:
        if 'Cout64 ' in l['layer'] or 'Cout128 ' in l['layer']: print("  %-58s %7.3f ms %7.1f TF  x%d"%(l['layer'],l['ms'],l['tflops'],l['launches']))
P
tail -3 gpurun_out/c12_bench_err.log | cut -c1-200
