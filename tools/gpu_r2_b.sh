#!/bin/bash
# full-size oracle parity, cfg5 launch list (tail composition), bench with the cfg5 strong-scaling record at N=1 (128 GiB buffer)
set -u
out=gpurun_out/r2_b
mkdir -p "$out"
echo "== fullsize parity"; timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -5 | tee "$out/pytest_fullsize.txt"
echo "== cfg5 launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches_cfg5.csv" python tools/run_configs.py cfg5 > "$out/launches_cfg5.log" 2>&1
tail -2 "$out/launches_cfg5.log"
echo "== bench N=1 with cfg5_strong"; timeout 900 python bench.py --steps 30 --warmup 5 2>&1 | tail -2 | tee "$out/bench_n1.json"
nvidia-smi --query-gpu=memory.total,memory.used --format=csv | tee "$out/mem.txt"
