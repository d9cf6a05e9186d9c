#!/bin/bash
# Round-2 GPU call R: final validation of the shipped build: full GPU test-suite, smoke, N=1 bench line, reference arm, chooser spot checks.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
tail -3 gpurun_out/r2r_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2r_smoke.log
timeout 300 python tools/sweep.py --envs 10,592,1024,1536,2048,4096 --ppl 0 --steps 50 > gpurun_out/r2r_sweep_auto.jsonl 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2r_ref.json 2> gpurun_out/r2r_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2r_sweep_auto.jsonl"):
    try:
        d = json.loads(l); print(d.get("envs"), d.get("ppl"), d.get("ms_per_period"))
    except Exception: print("?", l[:100])
d = json.loads(open("gpurun_out/r2r_bench_n1.json").read())
print(d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["sustained"]["value"], d["config_65536"]["value"], d["large_domain"]["f64"]["value"], d["spectral_mode"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["value"])
PY
