#!/bin/bash
# round 2, GPU call 16 (1 GPU): gate prefetch in the data-gradient epilogue (tests + A/B on the training step),
# two-stream graph replay A/B, launch lists of the training step and of resnet-H
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_kernel_variants.py -m gpu -q -x > gpurun_out/c16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c16_pytest.log
rm -f gpurun_out/c16_ab.txt
for v in base new base new; do
  if [ $v = base ]; then export DREAMB200_LIB=$PWD/variants/base16.so; else unset DREAMB200_LIB; fi
  timeout 300 python bench.py --workload vgg_q_train --steps 8 --layer-table gpurun_out/c16_layers_train_$v.json > gpurun_out/c16_bench_train_$v.json 2>> gpurun_out/c16_bench_err.log
  python -c "
import json; d=json.loads(open('gpurun_out/c16_bench_train_$v.json').read().strip().splitlines()[-1]); print('vgg_q_train $v', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'conv', round(d['roofline']['conv_stack']['ms_per_step'],2))" >> gpurun_out/c16_ab.txt
done
unset DREAMB200_LIB
for s in 1 2 1 2; do
  timeout 300 python bench.py --no-cpu-baseline --no-secondary --streams $s > gpurun_out/c16_bench_streams$s.json 2>> gpurun_out/c16_bench_err.log
  python -c "
import json; d=json.loads(open('gpurun_out/c16_bench_streams$s.json').read().strip().splitlines()[-1]); print('vgg_q_infer streams=$s', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'])" >> gpurun_out/c16_ab.txt
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/c16_launches_train.csv python bench.py --workload vgg_q_train --steps 1 --warmup 3 > gpurun_out/c16_launches_train_stdout.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/c16_launches_resnet_h.csv python bench.py --workload resnet_h_infer --steps 1 --warmup 3 --no-graph > gpurun_out/c16_launches_resnet_h_stdout.log 2>&1
tail -6 gpurun_out/c16_pytest.log | cut -c1-300; cat gpurun_out/c16_ab.txt
tail -3 gpurun_out/c16_bench_err.log | cut -c1-200
