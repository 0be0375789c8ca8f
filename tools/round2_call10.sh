#!/bin/bash
# round 2, GPU call 10 (1 GPU): precise mode parity, residual prefetch A/B (resnet), rs2 defaults, full suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c10_pytest.log
for w in resnet_h_infer resnet_f_infer; do
  for v in base new base new; do
    if [ $v = base ]; then export DREAMB200_LIB=$PWD/variants/base_c10.so; else unset DREAMB200_LIB; fi
    timeout 300 python bench.py --workload $w --steps 10 --layer-table gpurun_out/c10_layers_${w}_$v.json > gpurun_out/c10_bench_${w}_$v.json 2>> gpurun_out/c10_bench_err.log
    python -c "
import json; d=json.loads(open('gpurun_out/c10_bench_${w}_$v.json').read().strip().splitlines()[-1]); print('$w $v', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'conv', round(d['roofline']['conv_stack']['ms_per_step'],2))" >> gpurun_out/c10_ab.txt
  done
done
unset DREAMB200_LIB
timeout 400 python bench.py --layer-table gpurun_out/c10_layers_vgg_q_infer.json > gpurun_out/c10_bench.json 2>> gpurun_out/c10_bench_err.log
tail -8 gpurun_out/c10_pytest.log | cut -c1-300; cat gpurun_out/c10_ab.txt
python - <<'P'
import json
d=json.loads(open('gpurun_out/c10_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['roofline']['whole_step'], d['launch_mode'], d.get('latency_b1'))
for k,v in (d.get('secondary') or {}).items(): print('   ', k, {a:b for a,b in v.items() if a in ('value','ms_per_step','e2e_value','whole_step_frac_of_tensor_peak','launch_mode','error')})
P
grep -h "precise max-abs" gpurun_out/c10_pytest.log | head; tail -3 gpurun_out/c10_bench_err.log | cut -c1-200
