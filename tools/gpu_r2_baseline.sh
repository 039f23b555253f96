#!/bin/bash
# Round-2 opening call: re-validate, run the tests that were gated in round 1, baseline timings of the wide-rank shapes.
set -u
out=gpurun_out/r2_base
mkdir -p "$out"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > "$out/clocks.csv" &
SMI=$!
echo "== pytest -m gpu"; timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee "$out/pytest_gpu.txt"
echo "== gated tests"; DLRA_UNVALIDATED=1 timeout 300 python -m pytest tests/test_gpu_two_sided_terms.py tests/test_gpu_aug_basis_first.py -q 2>&1 | tail -25 | tee "$out/pytest_optin.txt"
echo "== perf r=16"; timeout 200 python tools/perf_pass.py 65536 4096 16 10 bug,ksl,rabug snapshot lookahead 2>&1 | tee "$out/perf16.txt"
timeout 200 python tools/perf_pass.py 65536 4096 16 10 bug,ksl delta 2>&1 | tee -a "$out/perf16.txt"
echo "== perf r=32/64"; timeout 200 python tools/perf_pass.py 65536 4096 32 10 bug,ksl delta,snapshot 2>&1 | tee "$out/perf32.txt"
timeout 200 python tools/perf_pass.py 262144 4096 64 5 bug,ksl delta 2>&1 | tee "$out/perf64.txt"
echo "== cfg5 shard"; timeout 300 python tools/run_configs.py cfg5 2>&1 | tail -3 | tee "$out/cfg5.txt"
kill $SMI
