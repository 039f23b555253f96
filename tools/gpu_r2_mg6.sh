#!/bin/bash
# 2-GPU A/B: LL polls through a never-matching 128-bit compare-and-swap (served at the owning L2 slice) vs volatile loads
set -u
N=2
out=gpurun_out/r2_mg6
mkdir -p "$out"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N"
echo "== multi-GPU parity with DLRA_LL_ATOMIC_POLL=1"; DLRA_LL_ATOMIC_POLL=1 timeout 600 python -m pytest tests/test_gpu_multi.py -q -k p2p 2>&1 | tail -3 | tee "$out/pytest_multi_atomic.txt"
for rep in 1 2; do
for a in 1 0; do
echo "== bench N=2 DLRA_LL_ATOMIC_POLL=$a #$rep"
DLRA_LL_ATOMIC_POLL=$a DLRA_PHASES=1 timeout 300 $RUN --steps 50 --warmup 5 --no-cfg5 --no-cpu-baseline > "$out/bench_a${a}_$rep.json" 2> "$out/bench_a${a}_$rep.err"; tail -1 "$out/bench_a${a}_$rep.json" | cut -c1-160; grep "dlra phases" "$out/bench_a${a}_$rep.err" | head -2
done
done
