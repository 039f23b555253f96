#!/bin/bash
# early L-sum (V0*S0' folded into the m-side TSQR) + two-level reduction in the fused Gram/core kernel: validation and A/B timings
set -u
out=gpurun_out/r2_n
mkdir -p "$out"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee "$out/pytest_gpu.txt"
for rep in 1 2; do
echo "== perf r=16 (new defaults) #$rep"; DLRA_PHASES=1 timeout 300 python tools/perf_pass.py 65536 4096 16 30 bug snapshot lookahead 2>&1 | tee -a "$out/perf16.txt"
echo "== perf r=16 (L sum at the start of the next step) #$rep"; DLRA_LSUM_EARLY=0 DLRA_PHASES=1 timeout 300 python tools/perf_pass.py 65536 4096 16 30 bug snapshot lookahead 2>&1 | tee -a "$out/perf16_lsum0.txt"
done
echo "== perf r=16 (all tail changes off)"; DLRA_LSUM_EARLY=0 DLRA_GRAM_M_AUX=0 DLRA_FUSED_CORE=0 DLRA_PHASES=1 timeout 300 python tools/perf_pass.py 65536 4096 16 30 bug snapshot lookahead 2>&1 | tee "$out/perf16_old.txt"
echo "== perf other shapes"; timeout 300 python tools/perf_pass.py 131072 2048 16 20 bug snapshot lookahead 2>&1 | tee "$out/perf_other.txt"
timeout 300 python tools/perf_pass.py 65536 4096 8 20 bug snapshot lookahead 2>&1 | tee -a "$out/perf_other.txt"
timeout 300 python tools/perf_pass.py 32768 1024 12 20 bug snapshot lookahead 2>&1 | tee -a "$out/perf_other.txt"
echo "== bench"; timeout 600 python bench.py --no-cfg5 2>"$out/bench.err" | tail -1 > "$out/bench_n1.json"; cut -c1-300 "$out/bench_n1.json"
echo "== launch list (bench command)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file "$out/launches_bench.csv" python bench.py --no-cpu-baseline --no-cfg5 --steps 10 --warmup 3 > "$out/launches_bench.log" 2>&1
ls -la "$out"
