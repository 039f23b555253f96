#!/bin/bash
set -u
out=gpurun_out/r2_j
mkdir -p "$out"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee "$out/pytest_gpu.txt"
echo "== perf r=16"; DLRA_PHASES=1 timeout 200 python tools/perf_pass.py 65536 4096 16 20 bug,rabug snapshot lookahead 2>&1 | tee "$out/perf16.txt"
echo "== cfg1 cfg3 cfg4"; timeout 900 python tools/run_configs.py cfg1,cfg3,cfg4 2>&1 | tail -14 | tee "$out/cfg134.txt"
echo "== bench"; timeout 600 python bench.py --no-cfg5 2>&1 | tail -1 | cut -c1-300 | tee "$out/bench.txt"
