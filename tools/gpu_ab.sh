#!/bin/bash
# Same-box A/B of one bench workload, alternating base / new (gpurun -- bash tools/gpu_ab.sh WORKLOAD [ENVVAR BASE NEW]):
#   bash tools/gpu_ab.sh vgg_q_train                     # base = variants/base.so (tools/build_variant.py), new = the tree's library
#   bash tools/gpu_ab.sh vgg_q_infer DREAMB200_RS3 0 3   # base / new = two values of a runtime switch
# Prints value, ms per step, e2e and the conv-stack time of every run; per-layer tables land in gpurun_out/.
w=${1:-vgg_q_infer}; var=$2; base=$3; new=$4
mkdir -p gpurun_out; rm -f gpurun_out/ab_$w.txt
for v in base new base new; do
  unset DREAMB200_LIB; [ -n "$var" ] && unset $var
  if [ -n "$var" ]; then if [ $v = base ]; then export $var=$base; else export $var=$new; fi
  elif [ $v = base ]; then export DREAMB200_LIB=$PWD/variants/base.so; fi
  timeout 400 python bench.py --workload $w --steps 10 --no-cpu-baseline --no-secondary --layer-table gpurun_out/ab_layers_${w}_$v.json > gpurun_out/ab_bench_${w}_$v.json 2>> gpurun_out/ab_err.log
  python -c "
import json; d=json.loads(open('gpurun_out/ab_bench_${w}_$v.json').read().strip().splitlines()[-1]); print('$w $v', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'conv', round(d['roofline']['conv_stack']['ms_per_step'],2), d['clocks']['sm_mhz'], d['clocks']['reasons'])" >> gpurun_out/ab_$w.txt
done
cat gpurun_out/ab_$w.txt; tail -3 gpurun_out/ab_err.log | cut -c1-200
