#!/bin/bash
# A/B of two builds of the CUDA library on the GPU box (interleaved runs): tools/ab_libs.sh "base new" "4096 16384"
# `new` = the in-tree build, any other NAME = mujoco_manip_b200/_C/variants/libmm_NAME.so
for rep in 1 2; do
  for v in $1; do
    for n in $2; do
      if [ $v = new ]; then unset MM_LIB_PATH; else export MM_LIB_PATH=$PWD/mujoco_manip_b200/_C/variants/libmm_$v.so; fi
      r=$(python bench.py --envs $n --steps 40 --warmup 8 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2))")
      echo "$v $n $r"
    done
  done
done
