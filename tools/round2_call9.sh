#!/bin/bash
# round 2, GPU call 9 (1 GPU): pool-first epilogue (tests + A/B), rs2 utilisation threshold A/B, flaky-gate fixes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_parity.py tests/test_gpu_training.py -m gpu -q > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c9_pytest.log
timeout 300 python bench.py --no-secondary --no-cpu-baseline --layer-table gpurun_out/c9_layers_a.json > gpurun_out/c9_bench_a.json 2> gpurun_out/c9_bench_err.log
DREAMB200_RS2=7 DREAMB200_RS2_MIN_UTIL=0.7 timeout 300 python bench.py --no-secondary --no-cpu-baseline --layer-table gpurun_out/c9_layers_b.json > gpurun_out/c9_bench_b.json 2>> gpurun_out/c9_bench_err.log
DREAMB200_RS2=15 DREAMB200_RS2_MIN_UTIL=0.7 timeout 300 python bench.py --no-secondary --no-cpu-baseline --layer-table gpurun_out/c9_layers_c.json > gpurun_out/c9_bench_c.json 2>> gpurun_out/c9_bench_err.log
tail -8 gpurun_out/c9_pytest.log | cut -c1-300
python - <<'P'
import json
for t in 'abc':
    try:
        d=json.loads(open('gpurun_out/c9_bench_%s.json'%t).read().strip().splitlines()[-1])
        print(t, round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['whole_step']['frac_of_peak'], round(d['roofline']['conv_stack']['ms_per_step'],3))
        L=json.load(open('gpurun_out/c9_layers_%s.json'%t))
        for l in L['layers']:
            if 'rs' in l['layer'] or 'first' in l['layer'] or 'pool' in l['layer']: print('    %-56s %.3f ms %.0f TF'%(l['layer'], l['ms'], l['tflops']))
    except Exception as e: print(t,'ERR',e)
P
tail -3 gpurun_out/c9_bench_err.log | cut -c1-200
