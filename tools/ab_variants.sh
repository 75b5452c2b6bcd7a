#!/bin/bash
# A/B of step-kernel CTA shapes on the GPU box: tools/ab_variants.sh "w8b1:0 w5b2:100000000" "4096 16384"
# (variants come from tools/build_variant.sh; NAME:T sets the MM_SMALL_BATCH threshold: 0 = always the large-CTA
# variant, a huge value = always the short-CTA variant)
for spec in default $1; do
  v=${spec%%:*}; t=${spec##*:}
  for n in $2; do
    if [ $v = default ]; then unset MM_LIB_PATH MM_SMALL_BATCH; else export MM_LIB_PATH=$PWD/mujoco_manip_b200/_C/variants/libmm_$v.so MM_SMALL_BATCH=$t; fi
    r=$(python bench.py --envs $n --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2))")
    echo "$v $n $r"
  done
done
