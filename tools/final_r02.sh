#!/bin/bash
# Round-2 measurement set (run under gpurun on one B200): every file lands in gpurun_out/final/
O=gpurun_out/final; mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/bench_r02_default.json 2> $O/err.log
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_r02_reference.json 2>> $O/err.log
python bench.py --envs 16384 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_r02_16k.json 2>> $O/err.log
python bench.py --envs 65536 --steps 12 --warmup 3 --no-cpu-baseline > $O/bench_r02_64k.json 2>> $O/err.log
python bench.py --envs 65536 --mode ee_pos_rot6d_g_rel --randomize --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_r02_rot6d64k.json 2>> $O/err.log
python bench.py --workload fsm --envs 16384 --steps 150 --warmup 5 > $O/bench_r02_fsm16k.json 2>> $O/err.log
python bench.py --workload mixed --envs 32768 --steps 150 --warmup 5 > $O/bench_r02_mixed32k.json 2>> $O/err.log
python tools/success_parity.py --episodes 2000 > $O/success_parity_r02.json 2>> $O/err.log
python tools/overflow_probe.py --envs 16384 --steps 500 > $O/overflow_r02.json 2>> $O/err.log
for f in $O/*.json; do echo "$f: $(head -c 260 $f)"; done
tail -5 $O/err.log
