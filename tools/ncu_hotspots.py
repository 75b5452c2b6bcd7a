#!/usr/bin/env python3
"""Join an ncu report's per-SASS-instruction counters with the line table of the cubin, and print
where a kernel spends its instructions / stall samples per SOURCE FUNCTION and per source line.

    python tools/ncu_hotspots.py gpurun_out/prof.ncu-rep mujoco_manip_b200/_C/obj/mm_inst_f64_32.o k_step [--top 40]

Needs ncu, cuobjdump and nvdisasm (all in the CUDA toolkit); runs on the CPU box.
"""
import argparse
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    res, hdr, take = [], None, False
    for r in rows:
        if not r:
            continue
        if r[0] == "Kernel Name":
            take = kernel in r[1]
            hdr = None
            continue
        if r[0] == "Address":
            hdr = r
            continue
        if take and hdr and len(r) >= len(hdr) - 2 and r[0].startswith("0x"):
            res.append(dict(zip(hdr, r)))
        if take and res and r[0] == "Kernel Name":
            break
    return res


def line_table(obj, kernel, outer=False):
    """(offset, (file, line), text) per SASS instruction.  outer=True: instructions of inlined helpers (warp intrinsics,
    mm_group.h) are charged to the nearest calling frame in mm_core.h / mm_env.h / mm_ccd.h (nvdisasm -gi)."""
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi" if outer else "-g", "-c", os.path.join(d, cub)], capture_output=True,
                         text=True).stdout.splitlines()
    table, cur, on, chain, fresh = [], ("?", 0), False, [], True
    for ln in dis:
        if ln.startswith(".text."):
            on = kernel in ln
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            if fresh:
                chain, fresh = [], False
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
        if m:
            if chain and not fresh:
                cur = chain[0]
                if outer:
                    cur = next((c for c in chain if c[0] in ("mm_core.h", "mm_env.h", "mm_ccd.h")), chain[0])
            fresh = True
            table.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return table


def function_ranges(srcdir):
    """(file, first line, name) of every function-like definition in the kernel headers."""
    out = {}
    pat = re.compile(r"^(?:template.*\n)?\s*(?:MM_HDN|MM_HD|__global__|static|inline)[^;{]*?\b([A-Za-z_0-9]+)\s*\(", re.M)
    for f in os.listdir(srcdir):
        if not f.endswith((".h", ".cuh", ".cu")):
            continue
        txt = open(os.path.join(srcdir, f)).read()
        lst = []
        for m in pat.finditer(txt):
            line = txt.count("\n", 0, m.start(1)) + 1
            lst.append((line, m.group(1)))
        out[f] = sorted(lst)
    return out


_FR = {}


def function_ranges_cache(srcdir):
    if srcdir not in _FR:
        _FR[srcdir] = function_ranges(srcdir)
    return _FR[srcdir]


def func_of(ranges, file, line):
    best = "?"
    for ln, name in ranges.get(file, []):
        if ln <= line:
            best = name
        else:
            break
    return f"{file}:{best}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("obj")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=30)
    ap.add_argument("--sym", default=None, help="substring of the (mangled) cubin symbol when it differs from `kernel`")
    ap.add_argument("--srcdir", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                     "mujoco_manip_b200", "csrc"))
    ap.add_argument("--stall", default=None, help="also list the top source lines by one stall reason (long_sb, no_inst, wait, short_sb ...)")
    ap.add_argument("--outer", action="store_true", help="charge inlined helpers (warp intrinsics, mm_group.h) to their callers")
    a = ap.parse_args()
    rows = sass_rows(a.rep, a.kernel)
    tab = line_table(a.obj, a.sym or a.kernel, outer=a.outer)
    if len(rows) != len(tab):
        print(f"warning: {len(rows)} profiled SASS rows vs {len(tab)} disassembled instructions", file=sys.stderr)
    base = int(rows[0]["Address"], 16)
    by_off = {off: (loc, ins) for off, loc, ins in tab}
    ranges = function_ranges(a.srcdir)
    agg_loc = collections.defaultdict(lambda: [0, 0])
    agg_f = collections.defaultdict(lambda: [0, 0, 0])
    agg_l = collections.defaultdict(lambda: [0, 0, 0])
    agg_s = collections.Counter()
    tot = [0, 0, 0]
    for r in rows:
        off = int(r["Address"], 16) - base
        loc, ins = by_off.get(off, (("?", 0), ""))
        if ins.startswith(("LDL", "STL")) or " LDL" in ins[:12] or " STL" in ins[:12]:
            k = f"{loc[0]}:{loc[1]}"
            agg_loc[k][0] += int(r.get("Instructions Executed", "0") or 0)
            agg_loc[func_of(function_ranges_cache(a.srcdir), loc[0], loc[1])][1] += int(r.get("Instructions Executed", "0") or 0)
        ie = int(r.get("Instructions Executed", "0") or 0)
        te = int(r.get("Thread Instructions Executed", "0") or 0)
        ss = int(r.get("Warp Stall Sampling (All Samples)", "0") or 0)
        if a.stall:
            agg_s[f"{loc[0]}:{loc[1]}"] += int(r.get("stall_" + a.stall, "0") or 0)
        for agg, key in ((agg_f, func_of(ranges, loc[0], loc[1])), (agg_l, f"{loc[0]}:{loc[1]}")):
            agg[key][0] += ie
            agg[key][1] += te
            agg[key][2] += ss
        tot[0] += ie
        tot[1] += te
        tot[2] += ss
    print(f"kernel {a.kernel}: {tot[0]:,} warp instructions, {tot[1]:,} thread instructions "
          f"({tot[1] / max(1, tot[0]):.1f} active lanes), {tot[2]:,} stall samples, {len(rows)} SASS instructions")
    print("\n== by source function (share of warp instructions | share of stall samples | active lanes) ==")
    for k, v in sorted(agg_f.items(), key=lambda kv: -kv[1][2])[: a.top]:
        print(f"{100 * v[0] / tot[0]:6.2f}% inst  {100 * v[2] / max(1, tot[2]):6.2f}% samples  {v[1] / max(1, v[0]):5.1f} lanes  {k}")
    print("\n== local-memory (LDL/STL) warp instructions by function ==")
    tl = sum(v[1] for v in agg_loc.values())
    for k, v in sorted(agg_loc.items(), key=lambda kv: -kv[1][1])[:20]:
        if v[1]:
            print(f"{v[1]:14,d}  {100 * v[1] / max(1, tl):5.1f}%  {k}")
    print("\n== local-memory warp instructions by line ==")
    for k, v in sorted(agg_loc.items(), key=lambda kv: -kv[1][0])[:25]:
        if v[0]:
            print(f"{v[0]:14,d}  {k}")
    print("\n== by source line ==")
    for k, v in sorted(agg_l.items(), key=lambda kv: -kv[1][2])[: a.top]:
        print(f"{100 * v[0] / tot[0]:6.2f}% inst  {100 * v[2] / max(1, tot[2]):6.2f}% samples  {v[1] / max(1, v[0]):5.1f} lanes  {k}")
    if a.stall:
        ts = sum(agg_s.values())
        print(f"\n== source lines by stall_{a.stall} samples ({ts:,} = {100 * ts / max(1, tot[2]):.1f} % of all samples) ==")
        src = {}
        for k, v in agg_s.most_common(a.top):
            f, ln = k.rsplit(":", 1)
            if f not in src:
                path = os.path.join(a.srcdir, f)
                src[f] = open(path).read().splitlines() if os.path.exists(path) else []
            text = src[f][int(ln) - 1].strip()[:100] if 0 < int(ln) <= len(src[f]) else ""
            print(f"{v:7d}  {100 * v / max(1, ts):5.1f}%  {k:22s} {text}")


if __name__ == "__main__":
    main()
