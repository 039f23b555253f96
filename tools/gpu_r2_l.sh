#!/bin/bash
# QR-preconditioned core SVD, hidden rank read-back, fused gram + core update: validation and A/B timings
set -u
out=gpurun_out/r2_l
mkdir -p "$out"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee "$out/pytest_gpu.txt"
echo "== perf r=16 (new defaults)"; DLRA_PHASES=1 timeout 300 python tools/perf_pass.py 65536 4096 16 20 bug,rabug snapshot lookahead 2>&1 | tee "$out/perf16.txt"
echo "== perf r=16 (old: no fused core, plain Jacobi)"; DLRA_FUSED_CORE=0 DLRA_JACOBI_PRE=0 DLRA_PHASES=1 timeout 300 python tools/perf_pass.py 65536 4096 16 20 bug,rabug snapshot lookahead 2>&1 | tee "$out/perf16_old.txt"
echo "== perf r=16 rabug aug-first"; DLRA_AUG=1 DLRA_PHASES=1 timeout 200 python tools/perf_pass.py 65536 4096 16 20 rabug snapshot 2>&1 | tee "$out/perf16_aug.txt"
echo "== cfg4"; DLRA_PHASES=1 timeout 600 python tools/run_configs.py cfg4 2>&1 | tee "$out/cfg4.txt" | grep -E "^cfg"
echo "== cfg4 aug-first"; CFG4_AUG=1 DLRA_PHASES=1 timeout 600 python tools/run_configs.py cfg4 2>&1 | tee "$out/cfg4_aug.txt" | grep -E "^cfg"
echo "== cfg4 plain Jacobi"; DLRA_JACOBI_PRE=0 timeout 600 python tools/run_configs.py cfg4 2>&1 | tee "$out/cfg4_nopre.txt" | grep -E "^cfg"
echo "== bench"; timeout 600 python bench.py --no-cfg5 2>"$out/bench.err" | tail -1 > "$out/bench_n1.json"; cut -c1-300 "$out/bench_n1.json"
echo "== launch list rabug"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_rabug.csv" python tools/perf_pass.py 65536 4096 16 3 rabug snapshot > "$out/launches_rabug.log" 2>&1
ls -la "$out"
