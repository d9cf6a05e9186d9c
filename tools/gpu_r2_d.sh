#!/bin/bash
# Round-2 GPU call D: ncu launch list of the bench + ncu --set full captures of the period kernels (FD-RK4: 4096 / 65536 envs,
# N=256; spectral: 8-lane 4096 / 65536, 16-lane 2048), spectral per-launch overhead experiment.
mkdir -p gpurun_out
S="python tools/sweep.py --steps 3"
E="--solver etdrk4 --dt 0.025 --cfg-steps 10"
# fixed per-launch cost of the spectral kernel: same physical period split into 10 / 20 / 40 ETDRK4 steps
for cs in 10 20 40; do
  python tools/sweep.py --solver etdrk4 --dt $(python -c "print(0.25/$cs)") --cfg-steps $cs --envs 2048,4096 --ppl 4,8 --steps 30 >> gpurun_out/r2d_etd_steps.jsonl 2>&1
done
python tools/sweep.py --solver etdrk4 --dt 0.025 --cfg-steps 10 --envs 2048,4096 --ppl 4,8 --steps 30 --rollout 30 >> gpurun_out/r2d_etd_steps.jsonl 2>&1
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-s 0.2"
$B > gpurun_out/r2d_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2d_launches_bench.csv $B > gpurun_out/r2d_ncu_bench.log 2>&1
echo "launch list rc=$?"
cap () {  # name, kernel regex, sweep args...
  name=$1; shift; rx=$1; shift
  $S "$@" > gpurun_out/r2d_plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 2 -f -o gpurun_out/r2d_$name $S "$@" > gpurun_out/r2d_ncu_$name.log 2>&1
  echo "$name rc=$?"
}
cap fd_4096 ks_period --envs 4096 --ppl 0
cap fd_65536 ks_period --envs 65536 --ppl 0
cap fd_n256 ks_period --envs 4096 --ppl 0 --N 256 --L 88 --J 8
cap fd_n256_f32 ks_period --envs 4096 --ppl 0 --N 256 --L 88 --J 8 --precision f32
cap etd8_4096 ks_etd $E --envs 4096 --ppl 8
cap etd8_65536 ks_etd $E --envs 65536 --ppl 8
cap etd16_2048 ks_etd16 $E --envs 2048 --ppl 4
ls -la gpurun_out/r2d_*.ncu-rep
cat gpurun_out/r2d_etd_steps.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print(d.get('envs'), d.get('ppl'), d.get('cfg_steps'), d.get('ms_per_period'))
    except Exception: print('?', l[:100])"
