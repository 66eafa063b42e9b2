#!/bin/bash
# Round-2 evidence in one GPU call: launch lists (C2 bench, C3 path), one full capture of the two top kernels, bench line.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2_c3.csv \
    python tools/time_path.py 4096 1024 6.666 > /dev/null 2>&1
bash tools/ncu_icp.sh icp_pairs_r2_final
# (launches alternate first-tier / second-tier and first kernel / item kernel: -s 2 is the second call's first one)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hist_score_kernel -s 2 -c 1 -o gpurun_out/hist_score_r2 -f \
    python tools/time_path.py 4096 1024 6.666 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hist_fused_kernel -s 2 -c 1 -o gpurun_out/hist_fused_r2_final -f \
    python tools/time_path.py 4096 1024 6.666 > /dev/null 2>&1
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_r2_n1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_r2_reference_arm.json
ls -la gpurun_out | tail -12
