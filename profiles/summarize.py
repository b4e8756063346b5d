#!/usr/bin/env python
"""ncu report -> markdown summary (one table per distinct kernel, last captured launch of each);
with a third argument also a JSON file of the numbers bench.py quotes (DRAM traffic, pipe view).
usage: summarize.py <report.ncu-rep> <title> [metrics.json]"""
import csv
import io
import json
import subprocess
import sys

rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_branch.sum", "sass__inst_executed_register_spilling",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]
STALL = [h for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
print(f"# {title}\n")
print("`ncu --set full --clock-control none --import-source on`, run after the same command exited 0 "
      "without ncu (profiles/collect.sh full).  Times under ncu are cold-cache and serialised.\n")
last = {}
for r in rows[2:]:
    last[r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")] = r
for name, r in last.items():
    print(f"## `{name}`\n\n| metric | value | unit |\n|---|---|---|")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"| {w} | {r[i]} | {units[i]} |")
    st = [(h, float(r[hdr.index(h)] or 0)) for h in STALL]
    tot = sum(v for _, v in st) or 1
    top = sorted(st, key=lambda kv: -kv[1])[:5]
    print("| top stall reasons (pc sampling) | " +
          ", ".join(f"{h.split('stalled_')[1]} {100 * v / tot:.0f}%" for h, v in top) + " | |\n")

if len(sys.argv) > 3:
    def num(r, key):
        try:
            return float(r[hdr.index(key)])
        except (ValueError, IndexError):
            return None

    def to_bytes(r, key):
        v, u = num(r, key), units[hdr.index(key)].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        return None if v is None else v * scale

    def short(name):
        # check_culled_kernel<stats, warps, obstacle in global memory, coplanar, one-triangle robot leaves>
        base = name.split("<")[0].replace("rc::", "")
        if base == "check_culled_kernel" and "<" in name:
            args = [a.strip() for a in name.split("<", 1)[1].rstrip(">").split(",")]
            if len(args) >= 3 and args[2] in ("1", "true", "(bool)1"):
                base += "_gobs"
        return base

    out = {}
    for name, r in last.items():
        conf, wave = num(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"), num(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
        out[short(name)] = {
            "kernel": name,
            "inst_executed": num(r, "smsp__inst_executed.sum"),
            "bank_conflict_share": (conf / wave) if conf is not None and wave else None,
            "lsu_pipe_pct": num(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
            "occupancy_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "duration_ms_under_ncu": num(r, "gpu__time_duration.sum"),
            "dram_bytes": (to_bytes(r, "dram__bytes_read.sum") or 0) + (to_bytes(r, "dram__bytes_write.sum") or 0),
            "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct": num(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "alu_pipe_pct": num(r, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
            "dram_throughput_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "warp_instructions": num(r, "smsp__inst_executed.sum"),
            "lanes_per_instruction": num(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
            "registers": num(r, "launch__registers_per_thread"),
        }
    json.dump(out, open(sys.argv[3], "w"), indent=1)
