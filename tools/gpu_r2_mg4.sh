#!/bin/bash
# end-of-round 2-GPU regression: row-sharded parity tests (both transports) + the driver's N=2 bench command
set -u
N=2
out=gpurun_out/r2_mg4
mkdir -p "$out"
echo "== multi-GPU parity (world 2)"; timeout 900 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3 | tee "$out/pytest_multi.txt"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N"
echo "== bench N=2 (driver command)"
time (timeout 900 $RUN --steps 50 --warmup 5 > "$out/bench.json" 2> "$out/bench.err"); tail -1 "$out/bench.json" | cut -c1-300
echo "== bench N=2 with phase marks"
DLRA_PHASES=1 timeout 600 $RUN --steps 50 --warmup 5 --no-cfg5 --no-cpu-baseline > "$out/bench_ph.json" 2> "$out/bench_ph.err"; tail -1 "$out/bench_ph.json" | cut -c1-160; grep "dlra phases" "$out/bench_ph.err" | head -2
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_mg4/bench.json').read().strip().splitlines()[-1])
print("cfg5_strong:", json.dumps(d.get('cfg5_strong'))[:700])
PY
