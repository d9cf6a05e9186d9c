#!/bin/bash
# Round-2 GPU call F: full GPU test-suite on the final build, 4096-vs-4736-env sweep, final N=1 bench line + reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -4 gpurun_out/r2f_pytest.log
timeout 300 python tools/sweep.py --envs 4096,4736,8192,9472 --ppl 16,8 --steps 30 > gpurun_out/r2f_sweep_fd.jsonl 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err; echo "ref rc=$?"
cat gpurun_out/r2f_sweep_fd.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print(d.get('envs'), d.get('ppl'), d.get('ms_per_period'), d.get('tflops_alg'))
    except Exception: print('?', l[:100])"
tail -3 gpurun_out/r2f_smoke.log
