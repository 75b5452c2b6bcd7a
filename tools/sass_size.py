#!/usr/bin/env python3
"""Static SASS size of a kernel by source function (how much instruction-cache each piece of the source costs).
    python tools/sass_size.py mujoco_manip_b200/_C/obj/mm_inst_f64_32.o k_stage_cIdLi32ELi1ELb0 [--top 25] [--outer]
--outer charges inlined helpers (shuffles, mm_group.h) to the calling function."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_hotspots as H

obj, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
srcdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mujoco_manip_b200", "csrc")
tab = H.line_table(obj, kernel, outer="--outer" in sys.argv)
ranges = H.function_ranges(srcdir)
agg = collections.Counter()
for off, loc, ins in tab:
    agg[H.func_of(ranges, loc[0], loc[1])] += 1
print(f"{kernel}: {len(tab)} SASS instructions = {len(tab) * 16 / 1024:.0f} KB")
for k, v in agg.most_common(top):
    print(f"{v:7d}  {100 * v / len(tab):5.1f}%  {k}")
