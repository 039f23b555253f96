#!/bin/bash
# First GPU call of a round: everything that has to be (re)confirmed on hardware in one box acquisition.
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/round_start.sh'
# Outputs land in gpurun_out/round_start/.
set -u
out=gpurun_out/round_start
mkdir -p "$out"
echo "== pytest -m gpu" ; timeout 200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee "$out/pytest_gpu.txt"
echo "== opt-in tests (two-sided right-hand-side terms, DLRA_AUG_BASIS_FIRST)"
DLRA_UNVALIDATED=1 timeout 200 python -m pytest tests/test_gpu_two_sided_terms.py tests/test_gpu_aug_basis_first.py -q 2>&1 | tail -25 | tee "$out/pytest_optin.txt"
echo "== step timings at 65536 x 4096, r = 16"
timeout 120 python tools/perf_pass.py 65536 4096 16 10 bug,ksl,rabug,greedy,greedy2,normal snapshot lookahead 2>&1 | tee "$out/perf.txt"
DLRA_AUG=1 timeout 60 python tools/perf_pass.py 65536 4096 16 10 rabug snapshot 2>&1 | sed 's/^/[aug-basis-first] /' | tee -a "$out/perf.txt"
echo "== smoke + bench"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4 | tee "$out/smoke.txt"
timeout 300 python bench.py --steps 50 --warmup 5 2> "$out/bench.err" | tee "$out/bench_n1.json"
