#!/bin/bash
# Round-2 GPU call O: host-visible completion word in ks_step_host (spin on the pinned block) vs cudaStreamSynchronize: e2e A/B + tests.
mkdir -p gpurun_out
F="--steps 200 --warmup 10 --no-sustained --no-episode --no-config-65536 --no-large-domain --no-spectral --no-cpu-baseline"
for i in 1 2; do
python bench.py $F > gpurun_out/r2o_flag_$i.json 2> gpurun_out/r2o_flag_$i.err; echo "flag rc=$?"
KS_HOST_SYNC=stream python bench.py $F > gpurun_out/r2o_stream_$i.json 2> gpurun_out/r2o_stream_$i.err; echo "stream rc=$?"
done
for f in gpurun_out/r2o_*.json; do python -c "
import json,sys; d=json.loads(open('$f').read()); print('$f', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), round(d['e2e']['value']/1e6,3))"; done
timeout 900 python -m pytest tests/test_gpu_env_api.py tests/test_gpu_parity.py tests/test_gpu_spectral.py tests/test_gpu_single_env.py tests/test_gpu_reference_stack.py -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2o_pytest.log
