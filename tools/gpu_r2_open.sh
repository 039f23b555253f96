#!/bin/bash
# Round-2 (re-entry) opening call: full GPU suite incl. the formerly gated tests and the wide-rank parity tests, then timings.
set -u
out=gpurun_out/r2_open
mkdir -p "$out"
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee "$out/pytest_gpu.txt"
echo "== perf r=16"; timeout 200 python tools/perf_pass.py 65536 4096 16 10 bug,ksl,rabug snapshot lookahead 2>&1 | tee "$out/perf16.txt"
timeout 200 python tools/perf_pass.py 65536 4096 16 10 bug,ksl delta 2>&1 | tee -a "$out/perf16.txt"
echo "== perf r=32"; timeout 200 python tools/perf_pass.py 65536 4096 32 10 bug,ksl delta,snapshot 2>&1 | tee "$out/perf32.txt"
echo "== perf r=64"; timeout 200 python tools/perf_pass.py 262144 4096 64 5 bug,ksl delta 2>&1 | tee "$out/perf64.txt"
echo "== cfg5 shard"; DLRA_DEBUG=1 timeout 300 python tools/run_configs.py cfg5 2>&1 | tail -8 | tee "$out/cfg5.txt"
echo "== bench"; timeout 300 python bench.py 2>&1 | tail -3 | tee "$out/bench.txt"
