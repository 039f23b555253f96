#!/bin/bash
set -u
out=gpurun_out/r2_c
mkdir -p "$out"
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee "$out/pytest_gpu.txt"
echo "== rabug aug"; DLRA_AUG=1 timeout 200 python tools/perf_pass.py 65536 4096 16 10 rabug snapshot 2>&1 | tee "$out/perf_rabug.txt"
timeout 200 python tools/perf_pass.py 65536 4096 16 10 rabug,bug snapshot lookahead 2>&1 | tee -a "$out/perf_rabug.txt"
echo "== cfg5 shard"; DLRA_PHASES=1 timeout 300 python tools/run_configs.py cfg5 2>&1 | tail -4 | tee "$out/cfg5.txt"
echo "== perf r=32/64"; timeout 200 python tools/perf_pass.py 65536 4096 32 10 bug,ksl delta 2>&1 | tee "$out/perf32.txt"
timeout 200 python tools/perf_pass.py 262144 4096 64 5 bug,ksl delta 2>&1 | tee "$out/perf64.txt"
