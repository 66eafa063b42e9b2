#!/bin/bash
# one full ncu capture of the dominant kernel on the C2 bench workload -> gpurun_out/<tag>.ncu-rep
tag=${1:-icp_pairs}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_pairs_kernel -s 6 -c 1 -o gpurun_out/$tag -f \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/$tag.log 2>&1
ls -la gpurun_out/$tag.ncu-rep
