#!/bin/bash
# round 2, GPU call 14 (1 GPU): the whole GPU suite + smoke + the default bench line (with e2e_u8) + launch list + ncu refresh
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c14_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c14_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c14_smoke.log
timeout 500 python bench.py --layer-table gpurun_out/c14_layers_vgg_q_infer.json > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench_err.log
timeout 300 python bench.py --impl reference --steps 3 > gpurun_out/c14_bench_reference.json 2>> gpurun_out/c14_bench_err.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/c14_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-graph > gpurun_out/c14_launches_stdout.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -k regex:"peaks_banded" -o gpurun_out/c14_peaks_banded -f python tools/ncu_targets.py peaks208 > gpurun_out/c14_ncu_peaks_banded.log 2>&1
tail -6 gpurun_out/c14_pytest.log | cut -c1-300; tail -2 gpurun_out/c14_smoke.log
python - <<'P'
import json
d=json.loads(open('gpurun_out/c14_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'e2e_u8', d.get('e2e_u8'), d['roofline']['whole_step'], d['launch_mode'], d.get('latency_b1'))
print(d['parity'])
for k,v in (d.get('secondary') or {}).items(): print('   ', k, {a:b for a,b in v.items() if a in ('value','ms_per_step','e2e_value','whole_step_frac_of_tensor_peak','launch_mode','error')})
print(open('gpurun_out/c14_bench_reference.json').read()[:600])
P
tail -3 gpurun_out/c14_bench_err.log | cut -c1-200
