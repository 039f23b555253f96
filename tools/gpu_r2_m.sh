#!/bin/bash
# M = U1'U0 on the auxiliary stream beside the pass (third buffer, stream priorities), 16 loads in flight in the last-CTA reductions,
# cheaper Jacobi rotation, preconditioned core SVD only from 64 x 64: validation and A/B timings
set -u
out=gpurun_out/r2_m
mkdir -p "$out"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee "$out/pytest_gpu.txt"
echo "== perf r=16 (new defaults)"; DLRA_PHASES=1 timeout 300 python tools/perf_pass.py 65536 4096 16 20 bug,rabug,ksl snapshot lookahead 2>&1 | tee "$out/perf16.txt"
echo "== perf r=16 (gram M on the main stream)"; DLRA_GRAM_M_AUX=0 DLRA_PHASES=1 timeout 300 python tools/perf_pass.py 65536 4096 16 20 bug snapshot lookahead 2>&1 | tee "$out/perf16_mmain.txt"
echo "== perf other shapes"; timeout 300 python tools/perf_pass.py 131072 2048 16 20 bug snapshot lookahead 2>&1 | tee "$out/perf_other.txt"
timeout 300 python tools/perf_pass.py 65536 4096 8 20 bug snapshot lookahead 2>&1 | tee -a "$out/perf_other.txt"
echo "== cfg4"; DLRA_PHASES=1 timeout 600 python tools/run_configs.py cfg4 2>&1 | tee "$out/cfg4.txt" | grep -E "^cfg"
echo "== bench"; timeout 600 python bench.py --no-cfg5 2>"$out/bench.err" | tail -1 > "$out/bench_n1.json"; cut -c1-300 "$out/bench_n1.json"
echo "== launch list (bench command)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file "$out/launches_bench.csv" python bench.py --no-cfg5 --no-cpu-baseline --steps 10 --warmup 3 > "$out/launches_bench.log" 2>&1
ls -la "$out"
