#!/bin/bash
# 8-GPU check (gpurun --gpus 8 -- bash tools/gpu_multigpu_check.sh; charged 8x the box time): 2-rank NCCL training test, then the default bench line at N = 8 (configs 2, 4, 5)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multigpu.py -m gpu -q > gpurun_out/chk_pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/chk_pytest_multigpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/chk_bench_n8.json 2> gpurun_out/chk_bench_n8.err
tail -4 gpurun_out/chk_pytest_multigpu.log | cut -c1-200
python - <<'P'
import json
d=json.loads(open('gpurun_out/chk_bench_n8.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'e2e_u8', d.get('e2e_u8'), d['clocks'])
for k,v in (d.get('secondary') or {}).items(): print(k, {a:b for a,b in v.items() if a in ('value','ms_per_step','n_gpus','e2e_value','allreduce')})
P
tail -3 gpurun_out/chk_bench_n8.err | cut -c1-200
