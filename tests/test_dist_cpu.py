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


def test_owner_blocks_cover_every_callback_once():
    from whitebox_b200 import shard
    for k in (1, 2, 5, 8, 4096, 4097):
        for w in (1, 2, 3, 8, 16):
            edges = [shard.owner_blocks(k, r, w) for r in range(w)]
            assert edges[0][0] == 0 and max(e[1] for e in edges) == k
            assert all(edges[i][1] == edges[i + 1][0] or edges[i + 1] == (k, k) for i in range(w - 1))
            assert sum(b - a for a, b in edges) == k


def test_reduce_owned_is_rank_ordered_sum_then_clamp():
    import numpy as np
    from whitebox_b200 import shard
    rng = np.random.RandomState(0)
    world, K = 3, 7
    parts = [(rng.rand(K, 2, 16).astype(np.float32) - 0.5) * 1.5 for _ in range(world)]
    got = np.concatenate([shard.reduce_owned(parts, r, world) for r in range(world)], axis=0)
    want = np.clip(((parts[0] + parts[1]).astype(np.float32) + parts[2]).astype(np.float32), -1, 1)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.abs(got).max() == 1.0
