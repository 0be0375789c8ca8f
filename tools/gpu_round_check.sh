#!/bin/bash
# One-GPU round check (gpurun -- bash tools/gpu_round_check.sh): whole GPU suite + smoke + default bench line + reference arm + launch list + ncu refresh
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/chk_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/chk_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/chk_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/chk_smoke.log
( time timeout 600 python bench.py --layer-table gpurun_out/chk_layers_vgg_q_infer.json > gpurun_out/chk_bench.json 2> gpurun_out/chk_bench_err.log ) 2> gpurun_out/chk_bench_time.txt
timeout 300 python bench.py --impl reference --steps 3 > gpurun_out/chk_bench_reference.json 2>> gpurun_out/chk_bench_err.log
for w in resnet_h_infer resnet_f_infer vgg_q_train; do
  timeout 300 python bench.py --workload $w --steps 10 --layer-table gpurun_out/chk_layers_$w.json > gpurun_out/chk_bench_$w.json 2>> gpurun_out/chk_bench_err.log
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/chk_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-graph > gpurun_out/chk_launches_stdout.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -k regex:"conv_rs3" -o gpurun_out/chk_rs3_pool -f python tools/ncu_targets.py rs64 > gpurun_out/chk_ncu_rs3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -k regex:"conv_rs2" -o gpurun_out/chk_dgrad64 -f python tools/ncu_targets.py dgrad64 > gpurun_out/chk_ncu_dgrad64.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -k regex:"conv_tc2" -o gpurun_out/chk_expand -f python tools/ncu_targets.py expand > gpurun_out/chk_ncu_expand.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -k regex:"wgrad3x3_pair" -o gpurun_out/chk_wgrad512 -f python tools/ncu_targets.py wgrad512 > gpurun_out/chk_ncu_wgrad512.log 2>&1
tail -6 gpurun_out/chk_pytest.log | cut -c1-300; tail -2 gpurun_out/chk_smoke.log; cat gpurun_out/chk_bench_time.txt
python - <<'P'
import json
d=json.loads(open('gpurun_out/chk_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'e2e_u8', d.get('e2e_u8',{}).get('value'), d['roofline']['whole_step'], d['roofline']['frac'], d['launch_mode'], d.get('latency_b1'), d['clocks'])
print(d['parity'])
for k,v in (d.get('secondary') or {}).items(): print('   ', k, {a:b for a,b in v.items() if a in ('value','ms_per_step','e2e_value','whole_step_frac_of_tensor_peak','launch_mode','error')})
print(open('gpurun_out/chk_bench_reference.json').read()[:400])
for w in ('resnet_h_infer','resnet_f_infer','vgg_q_train'):
    d=json.loads(open('gpurun_out/chk_bench_%s.json'%w).read().strip().splitlines()[-1]); print(w, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['roofline']['whole_step'], d['clocks'])
P
tail -3 gpurun_out/chk_bench_err.log | cut -c1-200
