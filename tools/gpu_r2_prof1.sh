#!/bin/bash
set -u
out=gpurun_out/r2_prof1
mkdir -p "$out"
export DLRA_FUSED_RT=16 DLRA_MAX_CLUSTER=2
# launch list of cfg5-shard BUG steps (1 warm + 3 timed): kernel durations under serialisation
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches_cfg5.csv" python tools/run_configs.py cfg5 > "$out/launches_cfg5.log" 2>&1
# full capture of one fused K+L sweep and one K-only sweep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pass_kernel -s 2 -c 2 -o "$out/prof_cfg5_pass" python tools/run_configs.py cfg5 > "$out/prof_cfg5.log" 2>&1
ls -la "$out"
