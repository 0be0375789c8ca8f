#!/bin/bash
# round 2, GPU call 5 (2 GPUs): NCCL 2-rank gradient test, vgg-Q training bench at N=2 and N=1 (same box), training
# gate calibration (deterministic BN reductions), resnet training tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q > gpurun_out/c5_pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c5_pytest_multigpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload vgg_q_train --steps 8 > gpurun_out/c5_bench_vgg_q_train_n2.json 2> gpurun_out/c5_bench_err.log
timeout 300 python bench.py --workload vgg_q_train --steps 8 > gpurun_out/c5_bench_vgg_q_train_n1.json 2>> gpurun_out/c5_bench_err.log
CUDA_VISIBLE_DEVICES=0 timeout 900 python tools/gpu_train_gates.py > gpurun_out/c5_train_gates.jsonl 2> gpurun_out/c5_train_gates.err &
CUDA_VISIBLE_DEVICES=1 timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -x -q > gpurun_out/c5_pytest_training.log 2>&1
wait
tail -3 gpurun_out/c5_pytest_multigpu.log; tail -3 gpurun_out/c5_pytest_training.log; cut -c1-1500 gpurun_out/c5_bench_vgg_q_train_n2.json; cut -c1-300 gpurun_out/c5_bench_vgg_q_train_n1.json; cat gpurun_out/c5_train_gates.jsonl; tail -5 gpurun_out/c5_train_gates.err; tail -3 gpurun_out/c5_bench_err.log
