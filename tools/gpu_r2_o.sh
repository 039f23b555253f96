#!/bin/bash
# one-launch two-level TSQR (tsqr_fused_kernel): validation and A/B timings
set -u
out=gpurun_out/r2_o
mkdir -p "$out"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee "$out/pytest_gpu.txt"
echo "== perf r=16 (fused TSQR)"; DLRA_PHASES=1 timeout 300 python tools/perf_pass.py 65536 4096 16 30 bug,ksl,rabug snapshot lookahead 2>&1 | tee "$out/perf16.txt"
echo "== perf r=16 (three-launch TSQR)"; DLRA_TSQR_FUSED=0 DLRA_PHASES=1 timeout 300 python tools/perf_pass.py 65536 4096 16 30 bug,ksl,rabug snapshot lookahead 2>&1 | tee "$out/perf16_unfused.txt"
echo "== perf r=8 / other shapes"; timeout 300 python tools/perf_pass.py 65536 4096 8 20 bug snapshot lookahead 2>&1 | tee "$out/perf_other.txt"
timeout 300 python tools/perf_pass.py 32768 1024 12 20 bug snapshot lookahead 2>&1 | tee -a "$out/perf_other.txt"
echo "== cfg1, cfg3"; timeout 600 python tools/run_configs.py cfg1,cfg3 2>&1 | grep -E "^cfg" | tee "$out/cfg13.txt"
echo "== cfg1, cfg3 (three-launch TSQR)"; DLRA_TSQR_FUSED=0 timeout 600 python tools/run_configs.py cfg1,cfg3 2>&1 | grep -E "^cfg" | tee "$out/cfg13_unfused.txt"
echo "== bench"; timeout 600 python bench.py --no-cfg5 2>"$out/bench.err" | tail -1 > "$out/bench_n1.json"; cut -c1-300 "$out/bench_n1.json"
echo "== launch list (bench command)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_bench.csv" python bench.py --no-cpu-baseline --no-cfg5 --steps 10 --warmup 3 > "$out/launches_bench.log" 2>&1
ls -la "$out"
