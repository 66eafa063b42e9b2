#!/usr/bin/env python
"""Print the key fields of one bench.py JSON line read from stdin (helper for sweeps under gpurun)."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline", {})
    print(tag, f"value={d.get('value', 0) / 1e6:.2f}M", f"ms/step={d.get('ms_per_step', 0):.4f}",
          f"kernel_ms={r.get('kernel_ms', 0):.4f}", f"frac={r.get('frac', 0):.4f}",
          f"e2e={d.get('e2e', {}).get('value', 0) / 1e6:.2f}M", d.get("nn_search", ""))
