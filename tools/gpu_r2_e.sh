#!/bin/bash
# Round-2 GPU call E (8 GPUs): fused exchange correctness at 4 and 8 GPUs, bench lines at N=8 and N=4 (all legs),
# exchange cost breakdown at N=8 (no exchange / stores only / + fence / full / NCCL).
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "4gpus or 8gpus" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -3 gpurun_out/r2e_pytest.log
timeout 600 $T --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2e_bench_n8.json 2> gpurun_out/r2e_bench_n8.err; echo "n8 rc=$?"
FAST="--steps 50 --warmup 5 --no-sustained --no-episode --no-config-65536"
timeout 300 $T --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 $FAST --gather none > gpurun_out/r2e_n8_none.json 2> gpurun_out/r2e_n8_none.err
KS_GATHER_DEBUG=nofence,nosignal timeout 300 $T --nproc-per-node 8 --master-port 29543 bench.py --gpus 8 $FAST > gpurun_out/r2e_n8_stores_only.json 2> gpurun_out/r2e_n8_stores_only.err
KS_GATHER_DEBUG=nosignal timeout 300 $T --nproc-per-node 8 --master-port 29544 bench.py --gpus 8 $FAST > gpurun_out/r2e_n8_stores_fence.json 2> gpurun_out/r2e_n8_stores_fence.err
KS_GATHER_DEBUG=nofence timeout 300 $T --nproc-per-node 8 --master-port 29545 bench.py --gpus 8 $FAST > gpurun_out/r2e_n8_nofence.json 2> gpurun_out/r2e_n8_nofence.err
timeout 300 $T --nproc-per-node 8 --master-port 29546 bench.py --gpus 8 $FAST > gpurun_out/r2e_n8_fused.json 2> gpurun_out/r2e_n8_fused.err
timeout 300 $T --nproc-per-node 8 --master-port 29547 bench.py --gpus 8 $FAST --gather nccl > gpurun_out/r2e_n8_nccl.json 2> gpurun_out/r2e_n8_nccl.err
timeout 600 $T --nproc-per-node 4 --master-port 29548 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2e_bench_n4.json 2> gpurun_out/r2e_bench_n4.err; echo "n4 rc=$?"
for f in gpurun_out/r2e_*.json; do python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read())
    print(sys.argv[1], d["n_gpus"], round(d["ms_per_step"], 4), round(d["value"] / 1e6, 2), d.get("gather_verified"), d.get("gather_mode"),
          (d.get("config_65536") or {}).get("value"), (d.get("sustained") or {}).get("value"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
