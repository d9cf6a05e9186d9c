#!/bin/bash
# Round-2 GPU call D2: ncu launch list of the bench + ncu --set full captures of the period kernels; the reports are
# exported to raw CSV on the box (gpurun copies back at most 64 MiB), only two .ncu-rep files are kept for the source page.
mkdir -p gpurun_out /tmp/ncu
S="python tools/sweep.py --steps 3"
E="--solver etdrk4 --dt 0.025 --cfg-steps 10"
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-s 0.2"
$B > gpurun_out/r2d_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2d_launches_bench.csv $B > gpurun_out/r2d_ncu_bench.log 2>&1
echo "launch list rc=$?"
cap () {  # name, kernel regex, sweep args...
  name=$1; shift; rx=$1; shift
  $S "$@" > gpurun_out/r2d_plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 2 -f -o /tmp/ncu/r2d_$name $S "$@" > gpurun_out/r2d_ncu_$name.log 2>&1
  echo "$name rc=$?"
  ncu -i /tmp/ncu/r2d_$name.ncu-rep --page raw --csv > gpurun_out/r2d_raw_$name.csv 2>/dev/null
}
cap fd_4096 ks_period --envs 4096 --ppl 0
cap fd_65536 ks_period --envs 65536 --ppl 0
cap fd_n256 ks_period --envs 4096 --ppl 0 --N 256 --L 88 --J 8
cap fd_n256_f32 ks_period --envs 4096 --ppl 0 --N 256 --L 88 --J 8 --precision f32
cap etd8_4096 ks_etd $E --envs 4096 --ppl 8
cap etd8_65536 ks_etd $E --envs 65536 --ppl 8
cap etd16_2048 ks_etd16 $E --envs 2048 --ppl 4
cp /tmp/ncu/r2d_fd_4096.ncu-rep /tmp/ncu/r2d_etd8_65536.ncu-rep gpurun_out/
du -sh gpurun_out
