#!/usr/bin/env python
"""Raw metrics of one ncu capture -> the small JSON bench.py quotes (`roofline.traffic`).

    python tools/ncu_metrics.py gpurun_out/x.ncu-rep "<how it was captured>" > profiles/ncu_r2_metrics.json
"""
import csv, io, json, subprocess, sys

rep, source = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(io.StringIO(raw)))
names, units, vals = rows[0], rows[1], rows[2]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3}
def get(name):
    i = names.index(name)
    return float(vals[i].replace(",", "")) * scale.get(units[i], 1.0)
out = {
    "source": source,
    "kernel": vals[names.index("Kernel Name")],
    "dram_bytes_read": get("dram__bytes_read.sum"),
    "dram_bytes_write": get("dram__bytes_write.sum"),
    "duration_us_under_ncu": get("gpu__time_duration.sum"),
    "registers_per_thread": get("launch__registers_per_thread"),
    "dynamic_smem_bytes": get("launch__shared_mem_per_block_dynamic"),
    "warp_instructions": get("smsp__inst_executed.sum"),
    "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "avg_active_threads_per_warp_inst": get("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
}
print(json.dumps(out, indent=1))
