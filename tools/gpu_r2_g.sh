#!/bin/bash
# Round-2 GPU call G: FD-RK4 kernel micro-variants (ptxas -O1, tap 0 as -u|u|) vs the shipped build: throughput + parity.
mkdir -p gpurun_out
for v in base o1 t0abs t0abs_o1; do
  lib=""; [ $v != base ] && lib="$PWD/build/libks_$v.so"
  KS_LIB_PATH=$lib timeout 300 python tools/sweep.py --envs 4096,65536 --ppl 16 --steps 30 > gpurun_out/r2g_sweep_$v.jsonl 2>&1
  KS_LIB_PATH=$lib timeout 300 python tools/sweep.py --envs 4096 --ppl 8 --N 256 --L 88 --J 8 --steps 30 >> gpurun_out/r2g_sweep_$v.jsonl 2>&1
  KS_LIB_PATH=$lib timeout 300 python tools/sweep.py --envs 4096 --ppl 0 --precision f32 --steps 30 >> gpurun_out/r2g_sweep_$v.jsonl 2>&1
  if [ $v != base ]; then KS_LIB_PATH=$lib timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2g_parity_$v.log 2>&1; echo "$v parity rc=$?"; fi
done
for v in base o1 t0abs t0abs_o1; do echo == $v; python - gpurun_out/r2g_sweep_$v.jsonl <<'PY'
import sys, json
for l in open(sys.argv[1]):
    try:
        d = json.loads(l); print(d.get("envs"), d.get("N"), d.get("precision"), d.get("ppl"), d.get("ms_per_period"), d.get("tflops_alg"), d.get("error"))
    except Exception: print("?", l[:120])
PY
done
