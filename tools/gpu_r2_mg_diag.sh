#!/bin/bash
set -u
N=${1:-2}
out=gpurun_out/r2_mgdiag$N
mkdir -p "$out"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N"
for mode in ${MODES:-p2p_aux p2p_main nccl}; do
  case $mode in
    p2p_aux) envs="DLRA_LFIN_AUX=1";;
    p2p_main) envs="";;
    nccl) envs="DLRA_COMM=nccl";;
    nccl_nosampler) envs="DLRA_COMM=nccl DLRA_BENCH_NO_SAMPLER=1";;
    p2p_nosampler) envs="DLRA_BENCH_NO_SAMPLER=1";;
    p2p_noll) envs="DLRA_NO_LL=1";;
  esac
  echo "== $mode"
  env $envs DLRA_PHASES=1 timeout 600 $RUN --steps 50 --warmup 5 --no-cfg5 > "$out/bench_$mode.json" 2> "$out/bench_$mode.err"
  python - "$out/bench_$mode.json" <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith('{')]
d=json.loads(l[-1]); print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'host_enqueue_ms',round(d.get('host_enqueue_ms_per_step',0),4),'launches',d['gpu_launches'])
PY
  grep "dlra phases" "$out/bench_$mode.err" | grep "50 steps" | cut -c1-600
done
