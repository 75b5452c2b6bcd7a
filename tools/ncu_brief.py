#!/usr/bin/env python3
"""Compact text summary of an `ncu --set full` report (one kernel launch): duration, instruction count, issue and
occupancy figures, stall reasons per issued instruction, memory traffic.  Runs on the CPU box (`ncu -i`).

    python tools/ncu_brief.py gpurun_out/r02_k_stage_c_final.ncu-rep > profiles/r02_k_stage_c_summary.txt
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def g(name, default="n/a"):
    return m.get(name, (default, ""))[0]


def f(name):
    try:
        return float(g(name).replace(",", ""))
    except ValueError:
        return float("nan")


print(f"report            {rep}")
print(f"kernel            {g('Kernel Name')}   grid {g('launch__grid_size')} x block {g('launch__block_size')}")
print(f"duration          {f('gpu__time_duration.sum') / (1e3 if m['gpu__time_duration.sum'][1] == 'ns' else 1):.1f} {'us' if m['gpu__time_duration.sum'][1] == 'ns' else m['gpu__time_duration.sum'][1]}")
print(f"registers/thread  {g('launch__registers_per_thread')}   dyn smem/CTA {g('launch__shared_mem_per_block_dynamic')} {m.get('launch__shared_mem_per_block_dynamic', ('', ''))[1]}"
      f"   occupancy limits: regs {g('launch__occupancy_limit_registers')} smem {g('launch__occupancy_limit_shared_mem')} CTAs/SM")
print(f"warp instructions {f('smsp__inst_executed.sum'):,.0f}   active lanes / instruction {f('smsp__thread_inst_executed_per_inst_executed.ratio'):.1f}")
print(f"issue active      {f('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} %   warps active {f('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} % of 64/SM"
      f"   FP64 pipe {f('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):.1f} %   SM throughput {f('sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} %")
print("stall cycles per issued instruction:")
st = []
for k in hdr:
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
        st.append((f(k), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
for v, k in sorted(st, reverse=True)[:9]:
    print(f"    {k:22s} {v:6.2f}")
rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
ur, uw = m["dram__bytes_read.sum"][1], m["dram__bytes_write.sum"][1]
print(f"DRAM              read {rd:.2f} {ur}   write {wr:.2f} {uw}   L2 hit rate {f('lts__t_sector_hit_rate.pct'):.1f} %")
print(f"local memory      load sectors {f('l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum'):,.0f}   store sectors {f('l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum'):,.0f}")
print(f"shared memory     wavefronts {f('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'):,.0f}   bank conflicts {f('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'):,.0f}")
