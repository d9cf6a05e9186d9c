#!/bin/bash
# Round-2 GPU call Q: P = 2 / 3 lane layouts (halo from two lanes): parity + small-batch latency sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_env_api.py -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2q_pytest.log
timeout 300 python tools/sweep.py --envs 1,10,64,256,512,592,768,1024,1184,1536,2048,3072,4096 --ppl 2,4,0 --steps 50 > gpurun_out/r2q_sweep_small.jsonl 2>&1
timeout 300 python tools/sweep.py --envs 10,512,1024 --ppl 2,4 --steps 50 --precision f32 >> gpurun_out/r2q_sweep_small.jsonl 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r2q_sweep_small.jsonl"):
    try:
        d = json.loads(l); print(d.get("envs"), d.get("precision"), d.get("ppl"), d.get("grid"), d.get("ms_per_period"), d.get("error"))
    except Exception: print("?", l[:120])
PY
