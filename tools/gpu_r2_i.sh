#!/bin/bash
set -u
out=gpurun_out/r2_i
mkdir -p "$out"
echo "== pytest (data + fullsize + storage)"; timeout 1500 python -m pytest tests/test_gpu_data_parity.py tests/test_gpu_fullsize.py tests/test_gpu_storage.py tests/test_gpu_de_parity.py -q -x 2>&1 | tail -5 | tee "$out/pytest_gpu.txt"
echo "== perf r=16"; DLRA_PHASES=1 timeout 200 python tools/perf_pass.py 65536 4096 16 20 bug snapshot lookahead 2>&1 | tee "$out/perf16.txt"
echo "== bench"; timeout 600 python bench.py --no-cfg5 2>&1 | tail -1 | cut -c1-300 | tee "$out/bench.txt"
echo "== launch list cfg4"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches_cfg4.csv" python tools/run_configs.py cfg4 > "$out/launches_cfg4.log" 2>&1
tail -3 "$out/launches_cfg4.log"
echo "== launch list cfg3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches_cfg3.csv" python tools/run_configs.py cfg3 > "$out/launches_cfg3.log" 2>&1
tail -3 "$out/launches_cfg3.log"
