#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<name>.md + traffic json.
usage: python tools/ncu_summary.py gpurun_out/k1_prof.ncu-rep profiles/r01_k1 [traffic_json]"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed.sum"]
lines = [f"# ncu summary of {rep}", "", "(ncu --set full --clock-control none; per launch)", ""]
traffic = []
for r in data:
    lines.append(f"## {r[hdr.index('Kernel Name')]}  (id {r[0]})")
    lines.append("")
    lines.append("| metric | unit | value |")
    lines.append("|---|---|---|")
    g = lambda k: r[hdr.index(k)] if k in hdr else None
    for k in want:
        if k in hdr:
            lines.append(f"| {k} | {units[hdr.index(k)]} | {g(k)} |")
    def tobytes(k):
        v, u = float(g(k)), units[hdr.index(k)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    tb = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
    tu = units[hdr.index("gpu__time_duration.sum")]
    t = float(g("gpu__time_duration.sum")) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}[tu]
    lines.append(f"| dram bytes read+write | byte | {tb:.0f} |")
    lines.append(f"| dram GB/s (under ncu) | GB/s | {tb / t / 1e9:.1f} |")
    lines.append("")
    traffic.append(tb)
open(out + ".md", "w").write("\n".join(lines) + "\n")
if len(sys.argv) > 3:
    json.dump({"dram_bytes_per_launch": sum(traffic) / len(traffic), "source": rep,
               "launches": len(traffic)}, open(sys.argv[3], "w"))
print("\n".join(lines))
