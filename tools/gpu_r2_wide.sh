#!/bin/bash
set -u
out=gpurun_out/r2_wide
mkdir -p "$out"
echo "== parity"; timeout 600 python -m pytest tests/test_gpu_data_parity.py tests/test_gpu_wide_rank.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -15 | tee "$out/pytest.txt"
echo "== perf r=16"; timeout 200 python tools/perf_pass.py 65536 4096 16 10 bug,ksl snapshot lookahead 2>&1 | tee "$out/perf16.txt"
timeout 200 python tools/perf_pass.py 65536 4096 16 10 bug,ksl delta 2>&1 | tee -a "$out/perf16.txt"
echo "== perf r=32"; timeout 200 python tools/perf_pass.py 65536 4096 32 10 bug,ksl delta,snapshot 2>&1 | tee "$out/perf32.txt"
echo "== perf r=64"; timeout 200 python tools/perf_pass.py 262144 4096 64 5 bug,ksl delta 2>&1 | tee "$out/perf64.txt"
echo "== cfg5 shard"; timeout 300 python tools/run_configs.py cfg5 2>&1 | tail -3 | tee "$out/cfg5.txt"
echo "== cfg5 shard KONLY_RT=16"; DLRA_KONLY_RT=16 timeout 300 python tools/run_configs.py cfg5 2>&1 | tail -3 | tee -a "$out/cfg5.txt"
echo "== cfg5 shard MAX_CLUSTER=2"; DLRA_MAX_CLUSTER=2 timeout 300 python tools/run_configs.py cfg5 2>&1 | tail -3 | tee -a "$out/cfg5.txt"
echo "== cfg5 shard MAX_CLUSTER=1"; DLRA_MAX_CLUSTER=1 timeout 300 python tools/run_configs.py cfg5 2>&1 | tail -3 | tee -a "$out/cfg5.txt"
