#!/bin/bash
# HEAD validation on one B200: full GPU suite, smoke(), default bench
set -u
out=gpurun_out/r2_final3
mkdir -p "$out"
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee "$out/pytest_gpu.txt"
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee "$out/smoke.txt"
echo "== bench (default command)"; timeout 600 python bench.py 2>"$out/bench.err" | tail -1 > "$out/bench_n1.json"; cut -c1-200 "$out/bench_n1.json"
