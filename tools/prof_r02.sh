#!/bin/bash
# GPU-side profiling driver (run under gpurun): steady-state launch list + full captures of the three stage kernels
TAG=$1
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 21"
export MM_STREAMS=1
$B > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
cat gpurun_out/plain_$TAG.log | tail -1 | cut -c1-300
# steady-state launch list: skip the first 21 steps (~75 launches each)
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 260 --csv --log-file gpurun_out/launches_ss_$TAG.csv $B > gpurun_out/ncu1_$TAG.log 2>&1
echo "launch list rc=$?"
for K in k_stage_c k_stage_a k_convex; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 365 -c 1 -f -o gpurun_out/r02_${K}_$TAG $B > gpurun_out/ncu_${K}_$TAG.log 2>&1
  echo "$K rc=$?"
done
ls -la gpurun_out/*.ncu-rep
