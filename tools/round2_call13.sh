#!/bin/bash
# round 2, GPU call 13 (1 GPU): fp32 output through the residual slots + TMA stores (tests, A/B, ncu)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/c13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c13_pytest.log
rm -f gpurun_out/c13_ab.txt
for w in resnet_h_infer resnet_f_infer; do
  for v in 0 1 0 1; do
    DREAMB200_RES_INPLACE=$v timeout 300 python bench.py --workload $w --steps 10 --layer-table gpurun_out/c13_layers_${w}_$v.json > gpurun_out/c13_bench_${w}_$v.json 2>> gpurun_out/c13_bench_err.log
    python -c "
import json; d=json.loads(open('gpurun_out/c13_bench_${w}_$v.json').read().strip().splitlines()[-1]); print('$w RES_INPLACE=$v', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'conv', round(d['roofline']['conv_stack']['ms_per_step'],2))" >> gpurun_out/c13_ab.txt
  done
done
timeout 300 ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -k regex:"conv_tc2" -o gpurun_out/c13_expand -f python tools/ncu_targets.py expand > gpurun_out/c13_ncu_expand.log 2>&1
tail -6 gpurun_out/c13_pytest.log | cut -c1-300; cat gpurun_out/c13_ab.txt
python - <<'P'
import json
for v in '01':
    d=json.load(open('gpurun_out/c13_layers_resnet_h_infer_%s.json'%v))
    print(v, sum(l['ms'] for l in d['layers']))
    for l in d['layers'][:6]: print("  %-58s %7.3f ms %7.1f TF  x%d"%(l['layer'],l['ms'],l['tflops'],l['launches']))
P
tail -3 gpurun_out/c13_bench_err.log | cut -c1-200
