#!/bin/bash
# GPU-side: device-timed value and end-to-end (host buffers) value for a list of launch-plan settings
#   tools/e2e_r02.sh TAG ENVS "ENV=V ENV=V" ...
TAG=$1; N=$2; shift 2
OUT=gpurun_out/e2e_$TAG.log
: > $OUT
for cfg in "$@"; do
  r=$(env $cfg python bench.py --envs $N --steps 20 --warmup 5 --no-cpu-baseline 2>>gpurun_out/e2e_$TAG.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2))")
  echo "$N [$cfg] value e2e ms: $r" >> $OUT
done
cat $OUT
