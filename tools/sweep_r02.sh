#!/bin/bash
# GPU-side experiment driver (run under gpurun): tests, then bench variants -> gpurun_out/
mkdir -p gpurun_out
OUT=gpurun_out/sweep_$1.log
: > $OUT
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_$1.log 2>&1; echo "gputest rc=$?" >> $OUT
tail -3 gpurun_out/gputest_$1.log >> $OUT
B="python bench.py --no-cpu-baseline --no-e2e --steps 20 --warmup 5"
run() { echo "## $*" >> $OUT; env "$@" $B $EXTRA 2>>gpurun_out/sweep_$1.err | python -c "import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print(json.dumps({k:j[k] for k in ('value','ms_per_step','gpu_launches')}), j['roofline']['kernel_ms_per_launch'])" >> $OUT; }
EXTRA=""
run MM_STREAMS=1
run MM_STREAMS=2
run MM_STREAMS=4
run MM_STREAMS=2 MM_CHUNK=1024
run MM_STREAMS=4 MM_CHUNK=512
run MM_STREAMS=2 MM_LOAD_BALANCE=0
run MM_STREAMS=1 MM_LOAD_BALANCE=0
EXTRA="--envs 16384"
run MM_STREAMS=2
run MM_STREAMS=4 MM_CHUNK=2048
EXTRA="--envs 65536 --steps 8"
run MM_STREAMS=4
# per-kernel durations of the default configuration
MM_STREAMS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$1.csv python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_$1.log 2>&1
echo "ncu rc=$?" >> $OUT
cat $OUT
