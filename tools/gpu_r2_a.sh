#!/bin/bash
# Round-2 GPU call A: full GPU test-suite, the new bench line, the reference arm, FD-kernel variants.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; echo "ref rc=$?"
for v in base both scatter; do
  lib=""; [ $v != base ] && lib="$PWD/build/libks_$v.so"
  KS_LIB_PATH=$lib timeout 300 python tools/sweep.py --envs 4096,65536 --ppl 16 --steps 20 > gpurun_out/r2a_sweep_$v.jsonl 2>&1
  KS_LIB_PATH=$lib timeout 300 python tools/sweep.py --envs 4096,16384 --ppl 8 --steps 20 >> gpurun_out/r2a_sweep_$v.jsonl 2>&1
done
tail -n 4 gpurun_out/r2a_sweep_*.jsonl
head -c 1500 gpurun_out/r2a_bench.json
