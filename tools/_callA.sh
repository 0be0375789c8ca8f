mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/A_smi.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/A_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/A_pytest.log
timeout 240 python bench.py > gpurun_out/A_bench_infer.json 2> gpurun_out/A_bench_infer.err
timeout 240 python bench.py --workload vgg_q_train --steps 5 --warmup 3 > gpurun_out/A_bench_train.json 2> gpurun_out/A_bench_train.err
tail -3 gpurun_out/A_pytest.log; cat gpurun_out/A_bench_infer.json | cut -c1-400
