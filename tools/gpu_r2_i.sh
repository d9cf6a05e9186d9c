#!/bin/bash
# Round-2 GPU call I (8 GPUs): NVLS multicast stores vs unicast peer stores for the exchange, 4096 envs per GPU.
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
FAST="--steps 50 --warmup 5 --no-sustained --no-episode"
timeout 300 $T --master-port 29571 bench.py --gpus 8 $FAST --gather fused_mc > gpurun_out/r2i_n8_fused_mc.json 2> gpurun_out/r2i_n8_fused_mc.err; echo "mc rc=$?"
timeout 300 $T --master-port 29572 bench.py --gpus 8 $FAST --gather fused > gpurun_out/r2i_n8_fused.json 2> gpurun_out/r2i_n8_fused.err; echo "fused rc=$?"
KS_GATHER_DEBUG=nofence,nosignal timeout 300 $T --master-port 29573 bench.py --gpus 8 $FAST --no-config-65536 --gather fused_mc > gpurun_out/r2i_n8_mc_stores_only.json 2> gpurun_out/r2i_n8_mc_stores_only.err
timeout 300 $T --master-port 29574 bench.py --gpus 8 $FAST --no-config-65536 --gather none > gpurun_out/r2i_n8_none.json 2> gpurun_out/r2i_n8_none.err
for f in gpurun_out/r2i_n8_*.json; do python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read())
    print(sys.argv[1], round(d["ms_per_step"], 4), d.get("gather_verified"), d.get("gather_mode"), d.get("collective_note"), (d.get("config_65536") or {}).get("ms_per_step"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
