#!/bin/bash
# round 2, GPU call 19 (1 GPU): residual ring fed by its own warp in conv_tc2 (tests, resnet A/B, ncu of the expansion layer)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/c19_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c19_pytest.log
tail -4 gpurun_out/c19_pytest.log | cut -c1-300
rm -f gpurun_out/c19_ab.txt
for w in resnet_h_infer resnet_f_infer; do
for v in base new new99 base new; do
  unset DREAMB200_LIB DREAMB200_RES_INPLACE_MAXKB
  if [ $v = base ]; then export DREAMB200_LIB=$PWD/variants/base19.so; fi
  if [ $v = new99 ]; then export DREAMB200_RES_INPLACE_MAXKB=99; fi
  timeout 300 python bench.py --workload $w --steps 10 --layer-table gpurun_out/c19_layers_${w}_$v.json > gpurun_out/c19_bench_${w}_$v.json 2>> gpurun_out/c19_bench_err.log
  python -c "
import json; d=json.loads(open('gpurun_out/c19_bench_${w}_$v.json').read().strip().splitlines()[-1]); print('$w $v', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'conv', round(d['roofline']['conv_stack']['ms_per_step'],2), d['clocks']['sm_mhz'], d['clocks']['reasons'])" >> gpurun_out/c19_ab.txt
done
done
unset DREAMB200_LIB DREAMB200_RES_INPLACE_MAXKB
timeout 300 ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -k regex:"conv_tc2" -o gpurun_out/c19_expand -f python tools/ncu_targets.py expand > gpurun_out/c19_ncu_expand.log 2>&1
cat gpurun_out/c19_ab.txt
python - <<'P'
import json
for w in ('resnet_h_infer',):
  for v in ('base','new'):
    d=json.load(open('gpurun_out/c19_layers_%s_%s.json'%(w,v)))
    print(w, v, sum(l['ms'] for l in d['layers']))
    for l in d['layers'][:12]: print("  %-58s %7.3f ms %7.1f TF  x%d"%(l['layer'],l['ms'],l['tflops'],l['launches']))
P
tail -3 gpurun_out/c19_bench_err.log | cut -c1-200
