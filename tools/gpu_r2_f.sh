#!/bin/bash
set -u
out=gpurun_out/r2_f
mkdir -p "$out"
echo "== cfg5 shard"; DLRA_PHASES=1 timeout 300 python tools/run_configs.py cfg5 2>&1 | tail -4 | tee "$out/cfg5.txt"
echo "== launch list cfg4"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches_cfg4.csv" python tools/run_configs.py cfg4 > "$out/launches_cfg4.log" 2>&1
tail -3 "$out/launches_cfg4.log"
echo "== launch list cfg3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches_cfg3.csv" python tools/run_configs.py cfg3 > "$out/launches_cfg3.log" 2>&1
tail -3 "$out/launches_cfg3.log"
echo "== launch list rabug r=16"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches_rabug16.csv" python tools/perf_pass.py 65536 4096 16 3 rabug snapshot > "$out/launches_rabug16.log" 2>&1
