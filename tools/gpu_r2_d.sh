#!/bin/bash
set -u
out=gpurun_out/r2_d
mkdir -p "$out"
echo "== pytest (qr-heavy subsets)"; timeout 1200 python -m pytest tests/test_gpu_data_parity.py tests/test_gpu_wide_rank.py tests/test_gpu_fullsize.py tests/test_gpu_greedy_hybrid.py -q -x 2>&1 | tail -5 | tee "$out/pytest_gpu.txt"
echo "== perf r=16"; DLRA_PHASES=1 timeout 200 python tools/perf_pass.py 65536 4096 16 20 bug,ksl,rabug snapshot lookahead 2>&1 | tee "$out/perf16.txt"
echo "== cfg5 shard"; DLRA_PHASES=1 timeout 300 python tools/run_configs.py cfg5 2>&1 | tail -4 | tee "$out/cfg5.txt"
echo "== launch list bug r=16"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches_bug16.csv" python tools/perf_pass.py 65536 4096 16 3 bug snapshot lookahead > "$out/launches_bug16.log" 2>&1
echo "== launch list cfg5"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches_cfg5.csv" python tools/run_configs.py cfg5 > "$out/launches_cfg5.log" 2>&1
