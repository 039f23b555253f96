#!/bin/bash
set -u
out=gpurun_out/r2_wide2
mkdir -p "$out"
echo "== parity"; timeout 600 python -m pytest tests/test_gpu_data_parity.py tests/test_gpu_wide_rank.py -x -q 2>&1 | tail -8 | tee "$out/pytest.txt"
echo "== parity FUSED_RT=16 MAX_CLUSTER=4"; DLRA_FUSED_RT=16 DLRA_KONLY_RT=16 timeout 600 python -m pytest tests/test_gpu_wide_rank.py -x -q 2>&1 | tail -4 | tee -a "$out/pytest.txt"
for cfg in "32 32 2" "32 32 4" "16 32 2" "32 32 1" ; do
  set -- $cfg
  echo "== cfg5 shard FUSED_RT=$1 KONLY_RT=$2 MAX_CLUSTER=$3"
  DLRA_DEBUG=1 DLRA_FUSED_RT=$1 DLRA_KONLY_RT=$2 DLRA_MAX_CLUSTER=$3 timeout 300 python tools/run_configs.py cfg5 2>&1 | tail -6 | tee -a "$out/cfg5.txt"
done
echo "== perf r=32"; 
for cfg in "32 1" "16 2"; do set -- $cfg; echo "FUSED_RT=$1 MAX_CLUSTER=$2"; DLRA_FUSED_RT=$1 DLRA_MAX_CLUSTER=$2 timeout 200 python tools/perf_pass.py 65536 4096 32 10 bug,ksl delta,snapshot 2>&1 | tee -a "$out/perf32.txt"; done
echo "== perf r=64 n=2^18"; timeout 200 python tools/perf_pass.py 262144 4096 64 5 bug,ksl delta 2>&1 | tee "$out/perf64.txt"
