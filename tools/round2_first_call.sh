#!/bin/bash
# The first GPU call of the next round, in one command: everything this round changed after its last GPU measurement
#  gets its parity run, its timing and its ncu evidence.  Run HERE first (builds the variants):
#
#     python tools/ab_kernel.py build nosameh=-DICPF_NO_SAME_H fullscan=-DICPF_GRIDNN_FULL_SCAN noresume=-DICPF_NO_RESUME \
#                                    coop=-DICPF_COOP_SEARCH
#     gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
#
# Outputs land in gpurun_out/ (copy what is to be judged into profiles/).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
  echo "== non-finite inputs on the GPU (first run: under its own timeout)"
  ICPF_RUN_NONFINITE_ON_GPU=1 timeout 120 python -m pytest tests/test_edge_fuzz.py -q -m gpu -k non_finite 2>&1 | tail -5
  echo "== bench (N=1)"; timeout 600 python bench.py 2>&1 | tail -2
  echo "== A/B (product + variants: C2 step, stop / big-cluster / init / hist_icp timings, output hashes)"
  timeout 600 python tools/ab_kernel.py run
  echo "== C3 stages (4096 x 1024, F = 6.666)"; timeout 300 python tools/time_path.py 4096 1024 6.666 2>&1 | tail -6
  echo "== frames"; timeout 300 python tools/time_frame.py 2>&1 | tail -4; timeout 300 python tools/time_frame.py c4 2>&1 | tail -4
} 2>&1 | tee gpurun_out/round2_first_call.log
# launch lists (cold-cache, serialised: shares, not absolutes) and one full capture of the two top kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2_c3.csv \
    python tools/time_path.py 4096 1024 6.666 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_pairs_kernel -s 6 -c 1 -o gpurun_out/icp_pairs_r2 -f \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hist_fused_kernel -s 1 -c 1 -o gpurun_out/hist_fused_r2 -f \
    python tools/time_path.py 1024 1024 6.666 > /dev/null 2>&1
ls -la gpurun_out | tail -12
