#!/bin/bash
# round 2, GPU call 8 (1 GPU): banded peaks fix, phase groups (+ N = 128 pair kernel), full GPU suite, bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c8_pytest.log
timeout 400 python bench.py --layer-table gpurun_out/c8_layers_vgg_q_infer.json > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench_err.log
DREAMB200_PHASE_GROUPS=0 timeout 300 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/c8_bench_nogroups.json 2>> gpurun_out/c8_bench_err.log
tail -12 gpurun_out/c8_pytest.log | cut -c1-300
python - <<'P'
import json
for f in ('gpurun_out/c8_bench.json','gpurun_out/c8_bench_nogroups.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['roofline']['whole_step'], d['roofline']['conv_stack']['ms_per_step'], d['launch_mode'], d.get('latency_b1'))
        for k,v in (d.get('secondary') or {}).items(): print('   ', k, {a:b for a,b in v.items() if a in ('value','ms_per_step','e2e_value','whole_step_frac_of_tensor_peak','launch_mode','error')})
    except Exception as e: print(f,'ERR',e)
P
tail -3 gpurun_out/c8_bench_err.log | cut -c1-300
