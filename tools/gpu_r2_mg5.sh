#!/bin/bash
# 2-GPU A/B: back-off between failed LL polls (DLRA_LL_BACKOFF_NS) vs a tight volatile-load loop
set -u
N=2
out=gpurun_out/r2_mg5
mkdir -p "$out"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N"
for rep in 1 2; do
for b in 0 200 1000; do
echo "== bench N=2 DLRA_LL_BACKOFF_NS=$b #$rep"
DLRA_LL_BACKOFF_NS=$b DLRA_PHASES=1 timeout 600 $RUN --steps 50 --warmup 5 --no-cfg5 --no-cpu-baseline > "$out/bench_b${b}_$rep.json" 2> "$out/bench_b${b}_$rep.err"; tail -1 "$out/bench_b${b}_$rep.json" | cut -c1-160; grep "dlra phases" "$out/bench_b${b}_$rep.err" | head -2
done
done
echo "== multi-GPU parity with DLRA_LL_BACKOFF_NS=200"; DLRA_LL_BACKOFF_NS=200 timeout 900 python -m pytest tests/test_gpu_multi.py -q -k p2p 2>&1 | tail -3 | tee "$out/pytest_multi_backoff.txt"
