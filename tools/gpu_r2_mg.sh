#!/bin/bash
# multi-GPU validation + phase breakdown: bash tools/gpu_r2_mg.sh N
set -u
N=${1:-2}
out=gpurun_out/r2_mg$N
mkdir -p "$out"
if [ "${SKIP_TESTS:-0}" != "1" ]; then echo "== multi-GPU parity (world 2)"; timeout 900 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -5 | tee "$out/pytest_multi.txt"; fi
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N"
echo "== bench N=$N (default: L all-reduce on aux, fused collectives)"
DLRA_PHASES=1 DLRA_PHASES_ALWAYS=1 timeout 600 $RUN --steps 50 --warmup 5 > "$out/bench.json" 2> "$out/bench.err"; tail -1 "$out/bench.json" | cut -c1-400; grep "dlra phases" "$out/bench.err" | head -4
echo "== bench N=$N, L all-reduce on the main stream"
DLRA_LFIN_MAIN=1 timeout 600 $RUN --steps 50 --warmup 5 --no-cfg5 > "$out/bench_lfin_main.json" 2> "$out/bench_lfin_main.err"; tail -1 "$out/bench_lfin_main.json" | cut -c1-300
echo "== bench N=$N, NCCL transport"
DLRA_COMM=nccl timeout 600 $RUN --steps 50 --warmup 5 --no-cfg5 > "$out/bench_nccl.json" 2> "$out/bench_nccl.err"; tail -1 "$out/bench_nccl.json" | cut -c1-300
