#!/bin/bash
set -u
out=gpurun_out/r2_prof2
mkdir -p "$out"
# full capture of one fused K+L sweep (cluster of 2, RT=16) and one K-only sweep (RT=32) at the cfg5 shard shape
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pass_kernel -s 3 -c 3 -o "$out/prof_cfg5_pass" python tools/run_configs.py cfg5 > "$out/prof_cfg5.log" 2>&1
ncu -i "$out/prof_cfg5_pass.ncu-rep" --page raw --csv > "$out/prof_cfg5_pass_raw.csv" 2>/dev/null
ls -la "$out"
