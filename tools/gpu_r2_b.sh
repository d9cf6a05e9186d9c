#!/bin/bash
# Round-2 GPU call B (2 GPUs): multi-GPU tests (fused exchange == NCCL == single GPU, sharded reset, late peer), bench at N=2.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_statistics.py -m gpu -x -q -k "multi or gather or spectral_large" > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -4 gpurun_out/r2b_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; echo "bench rc=$?"
head -c 600 gpurun_out/r2b_bench_n2.json; tail -3 gpurun_out/r2b_bench_n2.err
