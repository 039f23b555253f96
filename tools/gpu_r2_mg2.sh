#!/bin/bash
# 2-GPU regression of this session's changes: multi-GPU parity tests (both transports), N=2 bench with phase marks, A/B of M on the auxiliary stream
set -u
N=2
out=gpurun_out/r2_mg2b
mkdir -p "$out"
echo "== multi-GPU parity (world 2)"; timeout 900 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -5 | tee "$out/pytest_multi.txt"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N"
echo "== bench N=2 (defaults, with cfg5 strong)"
DLRA_PHASES=1 timeout 900 $RUN --steps 50 --warmup 5 > "$out/bench.json" 2> "$out/bench.err"; tail -1 "$out/bench.json" | cut -c1-300; grep "dlra phases" "$out/bench.err" | head -4
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_mg2b/bench.json').read().strip().splitlines()[-1])
print("cfg5_strong:", json.dumps(d.get('cfg5_strong'))[:600])
PY
echo "== bench N=2, M on the main stream"
DLRA_GRAM_M_AUX_MULTI=0 DLRA_PHASES=1 timeout 600 $RUN --steps 50 --warmup 5 --no-cfg5 > "$out/bench_mmain.json" 2> "$out/bench_mmain.err"; tail -1 "$out/bench_mmain.json" | cut -c1-200; grep "dlra phases" "$out/bench_mmain.err" | head -2
echo "== reference arm under torchrun"; timeout 600 $RUN --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
ls "$out"
