#!/usr/bin/env python
"""Capture the DRAM traffic of the bench kernels with ncu (run under gpurun, one GPU) and write the
records bench.py reports as `roofline.traffic`:

    gpurun_out/k1_traffic.json, k2_traffic.json, k2_streams8_traffic.json

Each record holds dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the bench's own command
(`bench.py --only ... --kernel-only`), the kernel time under ncu, and the sha of the kernel sources, so
that bench.py only quotes the figure while those sources are byte-identical (bench_common.read_traffic).
Copy the files to profiles/ to commit them."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_common as bc  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
METRICS = "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
JOBS = [
    ("k1_traffic.json", "cloud_tma_kernel", ["--only", "k1"], bc.K1_SOURCES),
    ("k2_traffic.json", "decode_pipe_kernel|decode_kernel", ["--only", "k2"], bc.K2_SOURCES),
    ("k2_streams8_traffic.json", "decode_pipe_kernel|decode_kernel", ["--only", "k2", "--streams-per-gpu", "8"],
     bc.K2_SOURCES),
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


for name, kern, extra, sources in JOBS:
    cmd = ["python", "bench.py", "--kernel-only", "--steps", "3", "--warmup", "3"] + extra
    ncu = ["ncu", "--metrics", METRICS, "--clock-control", "none", "-k", f"regex:{kern}", "-s", "4", "-c", "1",
           "--csv"] + cmd
    r = subprocess.run(ncu, cwd=ROOT, capture_output=True, text=True)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('"')]
    rec = {"cmd": " ".join(cmd), "source_sha": bc.source_sha(sources), "kernel": None}
    vals = {}
    for row in csv.DictReader(io.StringIO("\n".join(lines))):
        rec["kernel"] = row.get("Kernel Name")
        m, unit, v = row.get("Metric Name"), row.get("Metric Unit"), row.get("Metric Value")
        if m and v:
            vals[m] = (v, unit)
    if "dram__bytes_read.sum" in vals and "dram__bytes_write.sum" in vals:
        rd = to_bytes(*vals["dram__bytes_read.sum"])
        wr = to_bytes(*vals["dram__bytes_write.sum"])
        rec.update({"dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr})
        if "gpu__time_duration.sum" in vals:
            t, u = vals["gpu__time_duration.sum"]
            rec["kernel_time_under_ncu_us"] = float(t.replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
    else:
        rec["error"] = (r.stdout[-400:] + r.stderr[-400:])
    json.dump(rec, open(os.path.join(OUT, name), "w"), indent=1)
    print(name, {k: rec.get(k) for k in ("kernel", "dram_bytes_per_launch", "kernel_time_under_ncu_us", "error")})
