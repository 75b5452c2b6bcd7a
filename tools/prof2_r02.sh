#!/bin/bash
# steady-state launch list + full captures of chosen kernels: tools/prof2_r02.sh TAG VARIANT "k_stage_c k_stage_a"
TAG=$1; V=$2; KS=$3
mkdir -p gpurun_out
if [ "$V" != new ]; then export MM_LIB_PATH=$PWD/mujoco_manip_b200/_C/variants/libmm_$V.so; fi
export MM_STREAMS=1
B="python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 21"
$B > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 260 --csv --log-file gpurun_out/launches_ss_$TAG.csv $B > gpurun_out/ncu1_$TAG.log 2>&1
echo "launch list rc=$?"
for K in $KS; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 365 -c 1 -f -o gpurun_out/r02_${K}_$TAG $B > gpurun_out/ncu_${K}_$TAG.log 2>&1
  echo "$K rc=$?"
done
