#!/bin/bash
# One GPU call: the GPU test-suite, smoke(), the default bench line and the reference arm (short).  Logs -> gpurun_out/
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
  echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
  echo "== bench (N=1)"; timeout 900 python bench.py 2>&1 | tail -3
  echo "== bench --impl reference (3 steps)"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2
} 2>&1 | tee gpurun_out/gpu_check.log
