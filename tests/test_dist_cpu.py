"""world_size-2 gloo test of the sharded (multi-GPU) path's host logic on CPU — see tests/dist_worker.py."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_mix_two_ranks_gloo():
    import __graft_entry__ as ge
    ge.build_library()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29671",
                        os.path.join(ROOT, "tests", "dist_worker.py")],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "dist ok: world=2" in r.stdout


def test_track_range_partitions():
    from whitebox_b200 import shard
    for n in (0, 1, 7, 64, 1024, 4097):
        for w in (1, 2, 3, 8):
            edges = [shard.track_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
