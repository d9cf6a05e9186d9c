#!/bin/bash
# Round-2 GPU call H (2 GPUs): the exchange on symmetric memory with NVLS multicast stores: correctness + N=2 timing.
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "symmetric" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -5 gpurun_out/r2h_pytest.log
FAST="--steps 50 --warmup 5 --no-sustained --no-episode"
timeout 300 $T --nproc-per-node 2 --master-port 29561 bench.py --gpus 2 $FAST --gather fused_mc > gpurun_out/r2h_n2_fused_mc.json 2> gpurun_out/r2h_n2_fused_mc.err; echo "mc rc=$?"
timeout 300 $T --nproc-per-node 2 --master-port 29562 bench.py --gpus 2 $FAST --gather fused > gpurun_out/r2h_n2_fused.json 2> gpurun_out/r2h_n2_fused.err; echo "fused rc=$?"
for f in gpurun_out/r2h_n2_*.json; do python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read())
    print(sys.argv[1], round(d["ms_per_step"], 4), d.get("gather_verified"), d.get("gather_mode"), d.get("collective_note"), (d.get("config_65536") or {}).get("ms_per_step"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
tail -5 gpurun_out/r2h_n2_fused_mc.err
