#!/bin/bash
# round 2, GPU call 2: full GPU test suite with the new defaults, bench lines, ncu captures of the fused peaks kernel
# and the pair kernel on the wide layers, launch list of the bench command.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
timeout 300 python bench.py --layer-table gpurun_out/c2_layers_vgg_q_infer.json > gpurun_out/c2_bench_vgg_q_infer.json 2> gpurun_out/c2_bench_err.log
timeout 300 python bench.py --workload vgg_q_train --steps 8 --layer-table gpurun_out/c2_layers_vgg_q_train.json > gpurun_out/c2_bench_vgg_q_train.json 2>> gpurun_out/c2_bench_err.log
for t in peaks conv256; do
  timeout 300 ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -k regex:"peaks_fused|conv_tc2|conv_tc_kernel" -o gpurun_out/c2_$t -f python tools/ncu_targets.py $t > gpurun_out/c2_ncu_$t.log 2>&1
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/c2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c2_launches_stdout.log 2>&1
tail -5 gpurun_out/c2_pytest.log; cat gpurun_out/c2_bench_vgg_q_infer.json gpurun_out/c2_bench_vgg_q_train.json; tail -3 gpurun_out/c2_bench_err.log
