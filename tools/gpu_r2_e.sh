#!/bin/bash
set -u
out=gpurun_out/r2_e
mkdir -p "$out"
echo "== pytest"; timeout 1200 python -m pytest tests/test_gpu_data_parity.py tests/test_gpu_wide_rank.py tests/test_gpu_de_parity.py tests/test_gpu_aug_basis_first.py -q -x 2>&1 | tail -5 | tee "$out/pytest_gpu.txt"
echo "== cfg5 shard"; DLRA_PHASES=1 timeout 300 python tools/run_configs.py cfg5 2>&1 | tail -4 | tee "$out/cfg5.txt"
echo "== rabug"; DLRA_PHASES=1 timeout 200 python tools/perf_pass.py 65536 4096 16 20 rabug snapshot 2>&1 | tee "$out/perf_rabug.txt"
echo "== r=32,64"; timeout 200 python tools/perf_pass.py 65536 4096 32 10 bug,ksl,rabug delta 2>&1 | tee "$out/perf32.txt"
timeout 200 python tools/perf_pass.py 262144 4096 64 5 bug,ksl delta 2>&1 | tee "$out/perf64.txt"
echo "== cfg1 cfg3 cfg4"; timeout 900 python tools/run_configs.py cfg1,cfg3,cfg4 2>&1 | tail -14 | tee "$out/cfg134.txt"
