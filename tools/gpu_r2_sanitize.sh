#!/bin/bash
set -u
out=gpurun_out/r2_sanitize
mkdir -p "$out"
for tool in racecheck synccheck memcheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py > "$out/$tool.log" 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok |DONE|Error|hazard" "$out/$tool.log" | head -30
done
