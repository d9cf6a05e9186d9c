#!/bin/bash
# Round-2 GPU call C: spectral kernel, 16-lane layout: parity (both layouts) + batch-size sweep of both layouts / min-blocks variants.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spectral.py -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -4 gpurun_out/r2c_pytest.log
ENVS=1024,2048,4096,6144,8192,9472,12288,16384,32768,65536
for v in mb3 mb2 mb4; do
  lib=""; [ $v != mb3 ] && lib="$PWD/build/libks_$v.so"
  KS_LIB_PATH=$lib timeout 300 python tools/sweep.py --solver etdrk4 --dt 0.025 --cfg-steps 10 --envs $ENVS --ppl 4 --steps 30 > gpurun_out/r2c_sweep_etd16_$v.jsonl 2>&1
done
timeout 300 python tools/sweep.py --solver etdrk4 --dt 0.025 --cfg-steps 10 --envs $ENVS --ppl 8 --steps 30 > gpurun_out/r2c_sweep_etd8.jsonl 2>&1
timeout 300 python tools/sweep.py --solver etdrk4 --dt 0.025 --cfg-steps 10 --envs 4096,65536 --ppl 4,8 --steps 30 --precision f32 > gpurun_out/r2c_sweep_etd_f32.jsonl 2>&1
for f in gpurun_out/r2c_sweep_*.jsonl; do echo == $f; python - "$f" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    try:
        d = json.loads(l); print(d.get("envs"), d.get("ppl"), d.get("lanes"), d.get("regs"), d.get("ms_per_period"), d.get("periods_per_s"), d.get("error"))
    except Exception:
        print("?", l[:160])
PY
done
