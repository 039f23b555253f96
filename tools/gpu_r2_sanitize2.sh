#!/bin/bash
set -u
out=gpurun_out/r2_sanitize
mkdir -p "$out"
# racecheck again without the streaming-pass kernels (their mbarrier-synchronised L hand-off exhausts the hazard budget of the first run)
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 --kernel-regex-exclude kns=pass_kernel python tools/sanitize_driver.py > "$out/racecheck_nopass.log" 2>&1
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|Error: Race|ok |DONE" "$out/racecheck_nopass.log" | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | head -30
