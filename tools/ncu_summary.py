"""Summarise an ncu report (--set full, one kernel) into profiles/: python tools/ncu_summary.py rep.ncu-rep out.md [cells]"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
cells = float(sys.argv[3]) if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_inst_executed_op_tma_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
lines = [f"# ncu summary of `{rep.split('/')[-1]}`", "", "Captured with `ncu --set full --clock-control none --import-source on` under gpurun (one B200);",
         "per-launch times are cold-cache and serialised -- compare shares, not absolutes.", ""]
traffic = {}
for d in data:
    lines += ["| metric | value | unit |", "|---|---|---|"]
    rec = dict(zip(hdr, d))
    for k in want:
        if k in rec:
            lines.append(f"| {k} | {rec[k]} | {units[hdr.index(k)]} |")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and "average_warps" in h and "not_issued" not in h:
            try:
                if float(d[i].replace(",", "")) >= 0.25:
                    lines.append(f"| {h} | {d[i]} | {units[i]} |")
            except ValueError:
                pass
    def num(k):
        v = float(rec[k].replace(",", "")); u = units[hdr.index(k)]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
    tot = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    lines.append(f"| dram bytes per launch (read+write) | {tot:.4g} | byte |")
    if cells:
        lines.append(f"| dram bytes per cell-update | {tot / cells:.1f} | byte |")
    name = rec["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0].split("::")[-1]
    traffic[name] = {"dram_bytes_per_launch": tot, "cells": cells}
    lines.append("")
open(out, "w").write("\n".join(lines) + "\n")
print(json.dumps(traffic))
