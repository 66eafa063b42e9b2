#!/bin/bash
# TEST INFRASTRUCTURE: run the emulator-engine tests with the AddressSanitizer build of the kernel sources, i.e. a
# memcheck of every global-memory access the kernels make on the parity inputs (edge cases included) without a GPU.
#   tests/simt/memcheck.sh [pytest args]        e.g.  tests/simt/memcheck.sh tests/test_gpu_path.py -k large
# A violation aborts the run with the kernel's source line (icp_flow_b200/csrc/...).
set -e
cd "$(dirname "$0")/../.."
export ICPF_SIMT_ASAN=1
export ASAN_OPTIONS=detect_leaks=0:halt_on_error=1
export LD_PRELOAD="$(/usr/bin/gcc -print-file-name=libasan.so)"
if [ $# -eq 0 ]; then set -- tests/ -k "simt or fuzz or variants"; fi
python -m pytest -q -x -m "not gpu" "$@"
# ... and once more with the fibers of every block handed over in reverse order: same results = no order-dependent reads
SIMT_ORDER=reverse python -m pytest -q -x -m "not gpu" "$@"
