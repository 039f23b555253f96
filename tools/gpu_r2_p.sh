#!/bin/bash
# fused TSQR without nanosleep in its waits: validation and A/B (threshold 8 vs 2 row blocks)
set -u
out=gpurun_out/r2_p
mkdir -p "$out"
echo "== perf r=16 (fused from 8 blocks)"; DLRA_PHASES=1 timeout 300 python tools/perf_pass.py 65536 4096 16 30 bug,ksl,rabug snapshot lookahead 2>&1 | tee "$out/perf16.txt"
echo "== perf r=16 (fused from 2 blocks)"; DLRA_TSQR_FUSED_MIN=2 DLRA_PHASES=1 timeout 300 python tools/perf_pass.py 65536 4096 16 30 bug,ksl,rabug snapshot lookahead 2>&1 | tee "$out/perf16_min2.txt"
echo "== perf r=16 (three-launch TSQR)"; DLRA_TSQR_FUSED=0 DLRA_PHASES=1 timeout 300 python tools/perf_pass.py 65536 4096 16 30 bug,ksl snapshot lookahead 2>&1 | tee "$out/perf16_unfused.txt"
echo "== cfg3"; timeout 600 python tools/run_configs.py cfg3 2>&1 | grep -E "^cfg" | tee "$out/cfg3.txt"
echo "== cfg3 fused from 2"; DLRA_TSQR_FUSED_MIN=2 timeout 600 python tools/run_configs.py cfg3 2>&1 | grep -E "^cfg" | tee "$out/cfg3_min2.txt"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee "$out/pytest_gpu.txt"
echo "== pytest subset with fused from 2"; DLRA_TSQR_FUSED_MIN=2 timeout 900 python -m pytest tests/test_gpu_data_parity.py tests/test_gpu_fullsize.py tests/test_gpu_wide_rank.py -q -m gpu 2>&1 | tail -3 | tee "$out/pytest_gpu_min2.txt"
echo "== bench"; timeout 600 python bench.py --no-cfg5 2>"$out/bench.err" | tail -1 > "$out/bench_n1.json"; cut -c1-200 "$out/bench_n1.json"
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$out/launches_bench.csv" python bench.py --no-cpu-baseline --no-cfg5 --steps 10 --warmup 3 > "$out/launches_bench.log" 2>&1
