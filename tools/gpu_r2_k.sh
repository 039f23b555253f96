#!/bin/bash
# Round-2 evidence run at HEAD: full GPU test suite, bench line, launch list of the bench command, full ncu capture of the
# pipelined pass, the other BASELINE configs, step timings with phase marks.
set -u
out=gpurun_out/r2_k
mkdir -p "$out"
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee "$out/pytest_gpu.txt"
echo "== bench"; timeout 600 python bench.py --no-cfg5 2>"$out/bench.err" | tail -1 > "$out/bench_n1.json"; cut -c1-400 "$out/bench_n1.json"
echo "== perf r=16"; DLRA_PHASES=1 timeout 300 python tools/perf_pass.py 65536 4096 16 20 bug,rabug,ksl snapshot lookahead 2>&1 | tee "$out/perf16.txt"
echo "== perf r=16 rabug aug-first"; DLRA_AUG=1 timeout 200 python tools/perf_pass.py 65536 4096 16 20 rabug snapshot 2>&1 | tee "$out/perf16_aug.txt"
echo "== perf r=32"; timeout 300 python tools/perf_pass.py 65536 4096 32 10 bug,ksl delta,snapshot 2>&1 | tee "$out/perf32.txt"
echo "== configs"; DLRA_PHASES=1 timeout 900 python tools/run_configs.py cfg1,cfg3,cfg4,cfg5 2>&1 | tee "$out/configs.txt" | grep -E "^cfg" 
echo "== launch list (bench command)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file "$out/launches_bench.csv" python bench.py --no-cfg5 --no-cpu-baseline --steps 10 --warmup 3 > "$out/launches_bench.log" 2>&1
tail -2 "$out/launches_bench.log" | cut -c1-200
echo "== ncu full: tri_pass"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tri_pass -s 3 -c 1 -o "$out/prof_tri_pass" python tools/perf_pass.py 65536 4096 16 6 bug snapshot lookahead > "$out/prof_tri.log" 2>&1
ncu -i "$out/prof_tri_pass.ncu-rep" --page raw --csv > "$out/prof_tri_pass_raw.csv" 2>/dev/null
ls -la "$out"
