#!/bin/bash
# Round-2 GPU call T (N GPUs): flag rendezvous (ks_gather_barrier) as the rank re-alignment between timed steps: effect on `value`.
N=${1:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
FAST="--steps 50 --warmup 5 --no-episode --no-config-65536 --sustained-s 1.0"
timeout 300 $T --master-port 29601 bench.py --gpus $N $FAST > gpurun_out/r2t_n${N}_flagbarrier.json 2> gpurun_out/r2t_n${N}_flagbarrier.err; echo "rc=$?"
timeout 300 $T --master-port 29602 bench.py --gpus $N $FAST --no-flag-barrier > gpurun_out/r2t_n${N}_allreduce_only.json 2> gpurun_out/r2t_n${N}_allreduce_only.err; echo "rc=$?"
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "2gpus" > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2t_pytest.log; fi
for f in gpurun_out/r2t_n${N}_*.json; do python -c "
import json; d=json.loads(open('$f').read()); print('$f', round(d['ms_per_step'],4), round(d['value']/1e6,2), d.get('gather_verified'), d.get('gather_mode'), round(d['sustained']['ms_per_step'],4))"; done
