#!/usr/bin/env python
"""Per-phase instruction / stall summary of one icp_pairs_kernel capture.

    python tools/ncu_groups.py gpurun_out/x.ncu-rep      (uses the cubin of the product library in icp_flow_b200/)
Groups the per-line attribution of tools/ncu_by_line.py by the `// ----------------` phase markers of icpf_icploop.cuh.
"""
import os, re, subprocess, sys, tempfile
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "icp_flow_b200", "libicpflow_b200.so")
kern = "_ZN4icpf16icp_pairs_kernelILi3ELb0ELb0EEEvNS_7IcpArgsE"
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "icpf_icp", lib], cwd=tmp, stdout=subprocess.DEVNULL)
sass = os.path.join(tmp, "icp.sass")
with open(sass, "w") as f:
    subprocess.check_call(["nvdisasm", "-g", "-c", os.path.join(tmp, "icpf_icp.sm_100a.cubin")], stdout=f, stderr=subprocess.DEVNULL)
src = os.path.join(tmp, "src.csv")
with open(src, "w") as f:
    subprocess.check_call(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=f, stderr=subprocess.DEVNULL)
out = subprocess.check_output([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), src, sass, kern, "5000"], text=True)
loop = open(os.path.join(ROOT, "icp_flow_b200", "csrc", "icpf_icploop.cuh")).read().split("\n")
marks = [(i + 1, l.strip()[:70]) for i, l in enumerate(loop) if "// ----------------" in l]
def phase(f, l):
    if f == "icpf_kabsch.h": return "solve (kabsch)"
    if f != "icpf_icploop.cuh": return f
    cur = "loop: before first marker"
    for ln, t in marks:
        if l >= ln: cur = t
    return cur
g = defaultdict(lambda: [0, 0.0, 0.0])
tot = 0
for ln in out.split("\n"):
    m = re.match(r'(\S+):(\d+)\s+(\d+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+(\d+)\s+(\d+)', ln)
    if not m: continue
    f, l, inst, thr, smp = m.group(1), int(m.group(2)), int(m.group(3)), float(m.group(5)), float(m.group(6))
    k = phase(f, l); g[k][0] += inst; g[k][1] += inst * thr; g[k][2] += smp; tot += inst
print(out.split("\n")[0])
for k, v in sorted(g.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:72s} inst {100*v[0]/tot:5.1f}%  /CTA-iter {v[0]/20480:7.0f}  thr/inst {v[1]/max(v[0],1):5.1f}  samples {v[2]:5.1f}%")
