#!/bin/bash
# Build a variant of the FP64 / 32-lane kernel unit with extra -D flags and link it with the other (default) objects:
#   tools/build_variant.sh NAME -DMM_WA=1 -DMM_MINB_A=16
# -> mujoco_manip_b200/_C/variants/libmm_NAME.so   (use with MM_LIB_PATH=... python bench.py)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
obj=$root/mujoco_manip_b200/_C/obj
out=$root/mujoco_manip_b200/_C/variants
mkdir -p $out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -diag-suppress=170,128 "$@" \
  -Xptxas -v -c -o $out/mm_inst_f64_32_$name.o $root/mujoco_manip_b200/csrc/mm_inst_f64_32.cu 2> $out/ptxas_$name.log
others=$(ls $obj/*.o | grep -v mm_inst_f64_32.o)
nvcc --shared -cudart shared -gencode arch=compute_100a,code=sm_100a -o $out/libmm_$name.so $out/mm_inst_f64_32_$name.o $others
grep -A2 "Compiling entry function" $out/ptxas_$name.log | grep -o "k_[a-z_]*I\|Used [0-9]* registers\|[0-9]* bytes spill stores" | paste -sd' '
