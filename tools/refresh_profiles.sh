#!/bin/bash
# After `tools/final_r02.sh` + `tools/prof2_r02.sh final new "k_stage_c k_stage_a k_convex"` ran on the GPU box:
# copy the bench lines and regenerate the text summaries under profiles/ from the reports in gpurun_out/.
set -e
cd "$(dirname "$0")/.."
cp gpurun_out/final/*.json profiles/
OBJ=mujoco_manip_b200/_C/obj/mm_inst_f64_32.o
for k in stage_c stage_a convex; do python tools/ncu_brief.py gpurun_out/r02_k_${k}_final.ncu-rep > profiles/r02_k_${k}_summary.txt; done
python tools/ncu_hotspots.py gpurun_out/r02_k_stage_c_final.ncu-rep $OBJ k_stage_c --sym k_stage_cIdLi32ELi1ELb0 --outer --top 40 --stall long_sb > profiles/r02_k_stage_c_hotspots.txt 2>&1
python tools/ncu_hotspots.py gpurun_out/r02_k_stage_a_final.ncu-rep $OBJ k_stage_a --outer --top 40 --stall long_sb > profiles/r02_k_stage_a_hotspots.txt 2>&1
python tools/ncu_hotspots.py gpurun_out/r02_k_convex_final.ncu-rep $OBJ k_convex --outer --top 40 --stall no_inst > profiles/r02_k_convex_hotspots.txt 2>&1
(echo "# ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 260: steady-state launches of \`python bench.py --steps 3 --warmup 21\`"
 echo "# (MM_STREAMS=1: one chunk of 4,096 envs, kernels serialised, replayed from the captured step graph; cold-cache timings - compare SHARES, not absolutes)"
 python tools/launch_list.py gpurun_out/launches_ss_final.csv) > profiles/r02_launches_steady_state.txt
grep -h "DRAM\|^duration\|^kernel " profiles/r02_k_*_summary.txt
