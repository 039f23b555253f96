#!/bin/bash
# 2-GPU A/B: membar.sys after the LL pushes (DLRA_LL_FENCE=1) vs none
set -u
N=2
out=gpurun_out/r2_mg3
mkdir -p "$out"
echo "== multi-GPU parity (world 2) with DLRA_LL_FENCE=1"; DLRA_LL_FENCE=1 timeout 900 python -m pytest tests/test_gpu_multi.py -q -k p2p 2>&1 | tail -3 | tee "$out/pytest_multi_fence.txt"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N"
for rep in 1 2; do
for f in 0 1; do
echo "== bench N=2 DLRA_LL_FENCE=$f #$rep"
DLRA_LL_FENCE=$f DLRA_PHASES=1 timeout 600 $RUN --steps 50 --warmup 5 --no-cfg5 --no-cpu-baseline > "$out/bench_f${f}_$rep.json" 2> "$out/bench_f${f}_$rep.err"; tail -1 "$out/bench_f${f}_$rep.json" | cut -c1-160; grep "dlra phases" "$out/bench_f${f}_$rep.err" | head -2
done
done
ls "$out"
