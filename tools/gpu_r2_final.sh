#!/bin/bash
# end-of-round validation on one B200: full GPU test suite, smoke(), the default bench line (with the cfg-5 single-GPU record), configs, step timings
set -u
out=gpurun_out/r2_final
mkdir -p "$out"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee "$out/pytest_gpu.txt"
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -4 | tee "$out/smoke.txt"
echo "== perf r=16"; DLRA_PHASES=1 timeout 300 python tools/perf_pass.py 65536 4096 16 30 bug,ksl,rabug,greedy,greedy2 snapshot lookahead 2>&1 | tee "$out/perf16.txt"
echo "== configs"; DLRA_PHASES=1 timeout 900 python tools/run_configs.py cfg1,cfg3,cfg4,cfg5 2>&1 | grep -E "^cfg|phases" | tee "$out/configs.txt" | grep -E "^cfg"
echo "== bench (default command)"; /usr/bin/time -v timeout 900 python bench.py 2>"$out/bench.err" | tail -1 > "$out/bench_n1.json"; cut -c1-300 "$out/bench_n1.json"; grep -E "Elapsed|Maximum resident" "$out/bench.err"
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | tee "$out/bench_ref.json" | cut -c1-300
ls -la "$out"
