#!/bin/bash
# the driver's default bench command on one B200 (with the cfg-5 single-GPU record) + launch count check
set -u
out=gpurun_out/r2_final
mkdir -p "$out"
echo "== bench (default command)"; time (timeout 1200 python bench.py 2>"$out/bench.err" | tail -1 > "$out/bench_n1.json"); cut -c1-300 "$out/bench_n1.json"; tail -3 "$out/bench.err"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_final/bench_n1.json'))
print("cfg5_strong:", json.dumps(d.get('cfg5_strong'))[:900])
PY
