#!/bin/bash
set -u
out=gpurun_out/r2_g
mkdir -p "$out"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee "$out/pytest_gpu.txt"
echo "== rabug"; timeout 200 python tools/perf_pass.py 65536 4096 16 20 rabug snapshot 2>&1 | tee "$out/perf_rabug.txt"
DLRA_JACOBI_LEGACY=1 timeout 200 python tools/perf_pass.py 65536 4096 16 20 rabug snapshot 2>&1 | tee -a "$out/perf_rabug.txt"
echo "== cfg1 cfg3 cfg4"; timeout 900 python tools/run_configs.py cfg1,cfg3,cfg4 2>&1 | tail -14 | tee "$out/cfg134.txt"
