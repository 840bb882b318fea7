#!/usr/bin/env python
"""Turns what tools/ncu_r02.sh brought back in gpurun_out/ into the committed evidence under profiles/ (no GPU needed):
  profiles/roofline_traffic.json          DRAM bytes of one clip-kernel launch per workload (bench.py's roofline.traffic)
  profiles/r02_clip_ws_ncu_summary.txt     selected metrics of the --set full captures
  profiles/r02_launches_c4.csv             launch list of the default bench (kernel name, grid, duration)
  profiles/r02_sass_clip_ws.txt            SASS evidence of the hot kernel from the in-tree libdips_b200.so"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload table only)

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.avg.per_second',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio']


def rows_of(path):
    with open(path, newline="") as f:
        return [r for r in csv.reader(line for line in f if line.startswith('"'))]


def traffic():
    out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the clip kernel at the workload's full size on one GPU "
                       "(ncu single-pass metrics, tools/ncu_r02.sh run_dram); bench.py copies dram_bytes_per_launch into roofline.traffic"}
    for name, wl in bench.WORKLOADS.items():
        path = os.path.join(OUT, f"dram_{name}.csv")
        if not os.path.exists(path):
            continue
        vals = {r[12]: (float(r[14].replace(",", "")), r[4]) for r in rows_of(path)[1:]}
        _, w, h, fmt, mode, tau, total = wl
        fb = w * h * (3 if fmt in (0, 2) else 4)
        rd, wr = vals["dram__bytes_read.sum"][0], vals["dram__bytes_write.sum"][0]
        alg = total * fb + w * h * 10
        out[name] = {"dram_bytes_per_launch": int(rd + wr), "read": int(rd), "write": int(wr), "algorithmic": alg,
                     "ratio": round((rd + wr) / alg, 4), "kernel": re.sub(r"\(unnamed>::KParams\)", "", vals["dram__bytes_read.sum"][1]).replace("void unnamed>::", ""),
                     "ncu_duration_ms": round(vals["gpu__time_duration.sum"][0] / 1e6, 3), "source": f"gpurun_out/dram_{name}.csv (round 2)"}
    with open(os.path.join(PROF, "roofline_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    return out


def full_summaries():
    tags = [t for t in ("c4shard", "c4pshard", "c5oshard") if os.path.exists(os.path.join(OUT, f"prof_{t}_raw.csv"))]
    cols = {}
    for t in tags:
        rows = rows_of(os.path.join(OUT, f"prof_{t}_raw.csv"))
        hdr, units, vals = rows[0], rows[1], rows[2]
        for k in KEYS + ["Kernel Name"]:
            if k in hdr:
                i = hdr.index(k)
                cols.setdefault(k, {})[t] = (vals[i], units[i])
    lines = ["ncu --set full --clock-control none, one launch each of clip_kernel_ws on the shard shapes one of 8 GPUs runs (tools/ncu_r02.sh):",
             "  c4shard  = 450 frames of 3840x2160 RGBx8, overall     c4pshard = the same, per-frame mode     c5oshard = 150 frames of 7680x4320 RGB8, overall",
             "(times under ncu are serialised / cold-cache: shares and ratios, not absolutes)", ""]
    lines.append("%-92s" % "metric" + "".join("%18s" % t for t in tags))
    for k in ["Kernel Name"] + KEYS:
        if k in cols:
            u = next(iter(cols[k].values()))[1]
            lines.append("%-92s" % (k[-78:] + (" [" + u + "]" if u else "")) + "".join("%18s" % cols[k].get(t, ("-",))[0][-17:] for t in tags))
    with open(os.path.join(PROF, "r02_clip_ws_ncu_summary.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    return lines


def launches():
    src = os.path.join(OUT, "r02_launches_c4.csv")
    if not os.path.exists(src):
        return
    rows = rows_of(src)
    agg = collections.OrderedDict()
    with open(os.path.join(PROF, "r02_launches_c4.csv"), "w") as f:
        f.write("# python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-stream --sustained-s 0 --rows c3 under ncu --metrics gpu__time_duration.sum\n")
        f.write("id,kernel,block,grid,duration_us\n")
        for r in rows[1:]:
            name = re.sub(r"\(.*", "", r[4]).replace("void unnamed>::", "").replace("unnamed>::", "")
            us = float(r[14].replace(",", "")) / 1e3
            f.write(f"{r[0]},{name},{r[7].replace(',', ' ')},{r[8].replace(',', ' ')},{us:.1f}\n")
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1; a[1] += us
        tot = sum(v[1] for v in agg.values())
        f.write("# share of the summed kernel time per kernel:\n")
        for k, v in agg.items():
            f.write(f"#   {k}: {v[0]} launches, {v[1]:.1f} us, {100 * v[1] / tot:.1f} %\n")


def sass():
    so = os.path.join(ROOT, "dips_b200", "libdips_b200.so")
    names = subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout
    out = ["SASS evidence for the hot kernel of libdips_b200.so (cuobjdump -sass of the in-tree build, sm_100a); regenerate: python tools/ncu_collect.py", ""]
    whole = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", whole)))
    n_fun = len(re.findall(r"Function : ", whole))
    out.append(f"cubin arch: {', '.join(arch)}; {n_fun} kernels; local-memory instructions (LDL/STL) in the whole library: {len(re.findall(r'\\b(LDL|STL)', whole))}")
    per = collections.Counter()
    for fn, body in re.findall(r"Function : (\S+)(.*?)(?=Function : |\Z)", whole, flags=re.S):
        if "clip_kernel_ws" in fn:
            per[fn] = len(re.findall(r"\b(LDL|STL)", body))
    out.append(f"clip_kernel_ws instantiations: {len(per)}; with LDL/STL: {sum(1 for v in per.values() if v)}")
    out.append("")
    for pat, title in (("clip_kernel_wsILi4ELin1ELi0ELi3E", "clip_kernel_ws<4 B/px, all channels, overall, 3 stages>  (C4: 3840x2160 RGBx8)"),
                       ("clip_kernel_wsILi3ELin1ELi0ELi4E", "clip_kernel_ws<3 B/px, all channels, overall, 4 stages>  (C2 / C5: RGB8)")):
        fn = sorted(set(re.findall(r"(_ZN5dipsb\w*" + re.escape(pat) + r"\w*)", names)), key=len)[0]
        text = subprocess.run(["cuobjdump", "-sass", "-fun", fn, so], capture_output=True, text=True).stdout
        ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", text)]
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for _, t in ins)
        out.append(f"== {title}: {len(ins)} instructions")
        out.append("opcode histogram (whole kernel): " + ", ".join(f"{k} {v}" for k, v in ops.most_common(40)))
        idx = [k for k, (_, t) in enumerate(ins) if "REDUX" in t]
        # the stage-unrolled trip loop: the first run of evenly spaced REDUX (one per frame)
        lo = idx[0] - 140 if idx else 0
        first_wait = next((k for k in range(max(lo, 0), idx[0]) if "SYNCS.PHASECHK" in ins[k][1]), max(lo, 0))
        out.append(f"one frame of the trip loop (from the full-barrier wait to the per-frame scalar store; {idx[1] - idx[0]} instructions between two frames' REDUX):")
        for k in range(first_wait, idx[0] + 6):
            out.append(f"    /*{ins[k][0]:04x}*/  {ins[k][1]}")
        out.append("producer warp (TMA bulk copy per frame):")
        for k, (a, t) in enumerate(ins):
            if "UBLKCP" in t:
                for j in range(max(0, k - 4), k + 2):
                    out.append(f"    /*{ins[j][0]:04x}*/  {ins[j][1]}")
                break
        out.append("")
    with open(os.path.join(PROF, "r02_sass_clip_ws.txt"), "w") as f:
        f.write("\n".join(out) + "\n")


if __name__ == "__main__":
    t = traffic()
    print(json.dumps({k: v["ratio"] for k, v in t.items() if k != "_comment"}))
    print("\n".join(full_summaries()[:12]))
    launches()
    sass()
