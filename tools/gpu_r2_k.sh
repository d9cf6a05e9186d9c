#!/bin/bash
# Round-2 GPU call K (8 GPUs): final bench line at N=8 with the default exchange (multicast form), 8-GPU exchange test.
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $T --nproc-per-node 8 --master-port 29591 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2k_bench_n8.json 2> gpurun_out/r2k_bench_n8.err; echo "n8 rc=$?"
timeout 600 $T --nproc-per-node 4 --master-port 29592 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2k_bench_n4.json 2> gpurun_out/r2k_bench_n4.err; echo "n4 rc=$?"
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "8gpus" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -3 gpurun_out/r2k_pytest.log
for f in gpurun_out/r2k_bench_n*.json; do python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read())
    print(sys.argv[1], d["n_gpus"], round(d["ms_per_step"], 4), round(d["value"] / 1e6, 2), d.get("gather_verified"), d.get("gather_mode"), d.get("collective_note"),
          (d.get("config_65536") or {}).get("value"), (d.get("sustained") or {}).get("value"), d["e2e"]["value"], d["e2e_episode_amortised"]["value"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
