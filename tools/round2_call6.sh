#!/bin/bash
# round 2, GPU call 6 (1 GPU): training gates (curves, huber), new training tests, default bench line with secondary configs
mkdir -p gpurun_out
timeout 600 python tools/gpu_train_gates.py curve huber > gpurun_out/c6_train_gates.jsonl 2> gpurun_out/c6_train_gates.err
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_kernel_variants.py tests/test_gpu_dropin.py -m gpu -q > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c6_pytest.log
( time timeout 600 python bench.py > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench_err.log ) 2> gpurun_out/c6_bench_time.txt
grep -v "^\s*\(Manip\|Keyp\|Frien\|Archi\|.network\)" gpurun_out/c6_train_gates.jsonl; tail -3 gpurun_out/c6_train_gates.err; tail -15 gpurun_out/c6_pytest.log; cat gpurun_out/c6_bench.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['whole_step'], d['latency_b1'], d['launch_mode'])
print(json.dumps(d.get('secondary'), indent=1)[:3000])
"; cat gpurun_out/c6_bench_time.txt; tail -5 gpurun_out/c6_bench_err.log | cut -c1-300
