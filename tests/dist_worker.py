"""Worker of tests/test_dist_cpu.py (world_size 2, gloo): the sharded mixing path on CPU. Each rank schedules
ITS track shard with the product's host engine, renders the shard's UNCLAMPED partial bus (numpy statement of
the segment semantics — there is no CPU render path in the product), all-reduces the bus, clamps after the
reduce, and rank 0 checks the result against the reference's golden vector of the unsharded session."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenarios as sc  # noqa: E402
import segment_render as sr  # noqa: E402
from whitebox_b200 import shard  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_tracks, n_blocks, B = 64, 6, 512
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "cfg2_small.npz")))
    lo, hi = shard.track_range(n_tracks, rank, world)
    assert shard.track_range(n_tracks, world - 1, world)[1] == n_tracks
    # the same fixture as scenarios.standard, but this rank only creates its own tracks
    rng = np.random.RandomState(1234)
    eng = sr.ScheduleOnlyEngine(2, B, 48000, 120.0, batched=True, clamp=False)
    frames = int((n_blocks + 4) * B) + 64
    beats = (n_blocks + 2) * B / 48000 * 2.0
    for t in range(n_tracks):
        src = sc._src(rng, 2, frames, n_tracks)  # every rank draws the full stream to stay in step
        if lo <= t < hi:
            p = sc._track_params(t)
            i = eng.add_track(p["volume_db"], p["pan"], False)
            sid = eng.add_sample(src, 48000)
            eng.add_clip(i, sid, 0.0, beats, 0.0, 1.0, float(p["gain"]))
    eng.play()
    part, peaks = eng.process(n_blocks)  # unclamped partial bus [K][C][B], peaks of the shard
    bus = torch.from_numpy(np.array(part, dtype=np.float32, copy=True))  # all_reduce works in place: keep `part`
    dist.all_reduce(bus)  # the single exchange step
    out = shard.clamp_bus(bus.numpy())
    # the same exchange the way the product does it over peer memory: every rank hands every owner its slice of callbacks
    # (here: all_gather of the partial buses stands in for the peer stores), each owner adds the slices in rank order and
    # clamps, rank 0 collects the owners' slices. Must equal the all-reduce form bit for bit at world 2 (one add).
    parts = [torch.zeros_like(torch.from_numpy(part)) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(part)))
    mine = shard.reduce_owned([p.numpy() for p in parts], rank, world)
    lo_k, hi_k = shard.owner_blocks(n_blocks, rank, world)
    assert mine.shape[0] == hi_k - lo_k
    slices = [None] * world
    dist.all_gather_object(slices, mine)
    master = np.concatenate(slices, axis=0)
    assert master.shape == out.shape
    if world == 2:
        assert np.array_equal(master.view(np.uint32), out.view(np.uint32)), "owner-reduced bus != all-reduced bus"
    all_peaks = [torch.zeros((n_blocks, shard.track_range(n_tracks, r, world)[1] - shard.track_range(n_tracks, r, world)[0], 2))
                 for r in range(world)]
    dist.all_gather(all_peaks, torch.from_numpy(np.ascontiguousarray(peaks)))
    if rank == 0:
        pk = np.concatenate([p.numpy() for p in all_peaks], axis=1)
        assert np.array_equal(pk.view(np.uint32), gold["peaks"].view(np.uint32)), "sharded peaks differ"
        peak = np.abs(gold["out"]).max(axis=(1, 2), keepdims=True)
        err = np.abs(out.astype(np.float64) - gold["out"]).max(axis=(1, 2), keepdims=True)
        assert np.all(err <= 1e-5 * peak), "sharded bus outside 1e-5 of block peak: %g" % float((err / peak).max())
        print("dist ok: world=%d max err %.3g of block peak" % (world, float((err / peak).max())))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
