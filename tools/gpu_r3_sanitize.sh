#!/bin/bash
# compute-sanitizer over the step kernels, including the last round-2 session's kernels (tsqr_fused_kernel, gram_core_kernel).
# The one-launch TSQR hands data between CTAs inside the kernel: if a tool does not keep its <= 65 CTAs co-resident the waits time out
# (dlra_sync then fails, nothing hangs) -- run that configuration separately so that a time-out is not mistaken for a race.
set -u
out=gpurun_out/r3_sanitize
mkdir -p "$out"
for fused in 1 0; do
for tool in memcheck synccheck racecheck; do
  echo "== $tool DLRA_TSQR_FUSED=$fused"
  DLRA_TSQR_FUSED=$fused timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py > "$out/${tool}_fused$fused.log" 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok |DONE|Error|hazard|timed out" "$out/${tool}_fused$fused.log" | head -30
done
done
