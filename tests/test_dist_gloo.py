"""CPU (gloo, world_size 2) cover of the N>1 path: the row-sharded algebra + collectives the engine uses."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_sharded_algebra_world2_gloo():
    env = dict(os.environ, OMP_NUM_THREADS="2", CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "dist_gloo_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert "GLOO_SHARDED_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_row_shard_partition():
    sys.path.insert(0, os.path.join(ROOT, "lowrankintegrators.jl_b200"))
    from distributed import row_shard
    for n in (1, 7, 64, 65536, 4194304 + 3):
        for w in (1, 2, 4, 8):
            spans = [row_shard(n, w, k) for k in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
