#!/bin/bash
# Round-2 GPU call J (2 GPUs): refactored bench (default --gather fused -> fused_mc with verified fallback), multi-GPU tests.
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $T --nproc-per-node 2 --master-port 29581 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err; echo "n2 rc=$?"
timeout 300 $T --nproc-per-node 2 --master-port 29582 bench.py --gpus 2 --steps 20 --warmup 5 --no-sustained --no-episode --no-config-65536 --gather fused_ipc > gpurun_out/r2j_n2_ipc.json 2> gpurun_out/r2j_n2_ipc.err; echo "ipc rc=$?"
timeout 300 $T --nproc-per-node 2 --master-port 29583 bench.py --gpus 2 --steps 20 --warmup 5 --no-sustained --no-episode --no-config-65536 --solver etdrk4 > gpurun_out/r2j_n2_etd.json 2> gpurun_out/r2j_n2_etd.err; echo "etd rc=$?"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -3 gpurun_out/r2j_pytest.log
for f in gpurun_out/r2j_*.json; do python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read())
    print(sys.argv[1], round(d["ms_per_step"], 4), d.get("gather_verified"), d.get("gather_mode"), d.get("collective_note"), (d.get("config_65536") or {}).get("ms_per_step"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
