#!/bin/bash
# GPU-side A/B driver (run under gpurun): tools/sweep2_r02.sh TAG "variant ..." "envs ..." ["ENV=V ENV=V" ...]
# variant `new` = the in-tree build, any other NAME = mujoco_manip_b200/_C/variants/libmm_NAME.so
TAG=$1; VARIANTS=$2; ENVS=$3; shift 3
mkdir -p gpurun_out
OUT=gpurun_out/sweep2_$TAG.log
: > $OUT
[ $# -eq 0 ] && set -- "MM_STREAMS=2"
for v in $VARIANTS; do
  for n in $ENVS; do
    for cfg in "$@"; do
      if [ $v = new ]; then unset MM_LIB_PATH; else export MM_LIB_PATH=$PWD/mujoco_manip_b200/_C/variants/libmm_$v.so; fi
      steps=20; [ $n -gt 20000 ] && steps=8
      r=$(env $cfg python bench.py --envs $n --steps $steps --warmup 5 --no-cpu-baseline --no-e2e $BENCH_EXTRA 2>>gpurun_out/sweep2_$TAG.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2))")
      echo "$v $n [$cfg] $r" >> $OUT
    done
  done
done
cat $OUT
