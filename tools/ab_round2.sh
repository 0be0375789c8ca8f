#!/bin/bash
# First GPU call of the next round: validate + time the opt-in CTA-pair kernels prepared at the end of round 1.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/ab_round2.sh'
mkdir -p gpurun_out
timeout 400 python tools/tc2_check.py > gpurun_out/ab2_tc2.log 2>&1
timeout 400 python tools/wgrad_pair_check.py > gpurun_out/ab2_wgrad_pair.log 2>&1
for v in "" "DREAMB200_TC2=1"; do
  echo "infer [$v] $(env $v timeout 120 python bench.py --no-cpu-baseline --steps 20 2>> gpurun_out/ab2_err.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["conv_stack"]["ms_per_step"])')" >> gpurun_out/ab2.txt
done
for v in "" "DREAMB200_TC2=1" "DREAMB200_WGRAD3_2SM=1" "DREAMB200_TC2=1 DREAMB200_WGRAD3_2SM=1"; do
  echo "train [$v] $(env $v timeout 120 python bench.py --workload vgg_q_train --steps 6 2>> gpurun_out/ab2_err.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["conv_stack"]["ms_per_step"])')" >> gpurun_out/ab2.txt
done
python - <<'P'
import json
for f in ("gpurun_out/ab2_tc2.log", "gpurun_out/ab2_wgrad_pair.log"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); a = d.get("tc2") or d.get("pair")
            print(d["name"], d.get("identical", d.get("ok")), d["ref"].get("ms"), a.get("ms"), str(a.get("error", ""))[:200])
        else:
            print(l.strip())
P
cat gpurun_out/ab2.txt
