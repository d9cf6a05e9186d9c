#!/bin/bash
# Round-2 GPU call L: block size 32 / 64 / 128 threads (512 / 256 / 128 blocks at 4096 envs): does spreading the 512 warps over
# all 148 SMs instead of 128 change the per-warp time?
mkdir -p gpurun_out
for v in base bt64 bt32; do
  lib=""; [ $v != base ] && lib="$PWD/build/libks_$v.so"
  KS_LIB_PATH=$lib timeout 300 python tools/sweep.py --envs 2048,4096,4736,65536 --ppl 16 --steps 30 > gpurun_out/r2l_sweep_$v.jsonl 2>&1
  KS_LIB_PATH=$lib timeout 300 python tools/sweep.py --solver etdrk4 --dt 0.025 --cfg-steps 10 --envs 4096,65536 --ppl 8 --steps 30 >> gpurun_out/r2l_sweep_$v.jsonl 2>&1
done
for v in base bt64 bt32; do echo == $v; python - gpurun_out/r2l_sweep_$v.jsonl <<'PY'
import sys, json
for l in open(sys.argv[1]):
    try:
        d = json.loads(l); print(d.get("envs"), d.get("solver"), d.get("ppl"), d.get("grid"), d.get("ms_per_period"), d.get("error"))
    except Exception: print("?", l[:120])
PY
done
