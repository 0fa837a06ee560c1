#!/usr/bin/env python
"""Offline bounce at scale: N stereo tracks x M minutes exported to an interleaved device format through
wbx::Engine::bounce, against the plain batched render of the same audio. Each track plays one 30 s sample as back-to-back
clips (1024 tracks x 10 min of distinct audio would be 236 GB).

    python tools/bounce_bench.py [--tracks 1024] [--minutes 10] [--chunk 1024] [--fmt I24_X8]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=1024)
    ap.add_argument("--minutes", type=float, default=10.0)
    ap.add_argument("--chunk", type=int, default=1024)
    ap.add_argument("--fmt", default="I24_X8")
    ap.add_argument("--check-blocks", type=int, default=8)
    args = ap.parse_args()
    import whitebox_b200 as wb
    import oracle_api as o
    B, rate = 512, 48000
    code = {"I16": wb.FMT_I16, "I24_X8": wb.FMT_I24_X8, "I32": wb.FMT_I32, "F32": wb.FMT_F32}[args.fmt]
    N = args.tracks
    clip_s = 30.0
    frames = int(clip_s * rate)
    beats_per_clip = clip_s * 2.0  # bpm 120
    n_clips = int(np.ceil(args.minutes * 60.0 / clip_s))
    end_beat = args.minutes * 60.0 * 2.0
    rng = np.random.default_rng(99)
    pool = ((rng.random(1 << 25, dtype=np.float32) * 2.0 - 1.0) * np.float32(0.5 / np.sqrt(N))).astype(np.float32)
    eng = wb.Engine(2, B, rate, 120.0, device=0, sum_mode=wb.SUM_EXACT)
    ref = o.Session("reference" if o.have_ref() else "port", 2, B, rate, 120.0)
    check_frames = (args.check_blocks + 2) * B
    for t in range(N):
        offs = [((2 * t + c) * 7919 * 4099) % (pool.size - frames) for c in range(2)]
        x = [pool[a:a + frames] for a in offs]
        vol, pan, gain = -6.0 - (t % 7), -1.0 + 0.2 * (t % 11), float(np.float32(0.5 + 0.001 * (t % 512)))
        eng.add_track(vol, pan, False)
        sid = eng.add_sample_planar(x, rate)
        ref.add_track(vol, pan, False)
        rsid = ref.add_sample(np.stack([x[0][:check_frames], x[1][:check_frames]]), rate)
        for k in range(n_clips):
            eng.add_clip(t, sid, k * beats_per_clip, (k + 1) * beats_per_clip, 0.0, 1.0, gain)
        ref.add_clip(t, rsid, 0.0, beats_per_clip, 0.0, 1.0, gain)
    total_frames = int(np.ceil(end_beat * 0.5 * rate))
    size = {wb.FMT_I16: 2, wb.FMT_I24_X8: 4, wb.FMT_I32: 4, wb.FMT_F32: 4}[code]
    out = np.zeros((total_frames + B) * 2 * size, np.uint8)
    out.fill(1)  # touch every page of the destination: the timed runs should not pay first-touch page faults
    eng.bounce(0.0, 8.0, code, chunk_blocks=args.chunk, out=out)  # warm-up (buffers, kernels)
    t_bounce = 1e9
    for _ in range(3):  # best of three: a single 80 ms run is at the mercy of the host's memory system
        t0 = time.perf_counter()
        got = eng.bounce(0.0, end_beat, code, chunk_blocks=args.chunk, out=out)
        t_bounce = min(t_bounce, time.perf_counter() - t0)
    assert got.size == total_frames * 2 * size
    # the same audio through the plain batched render (bus back as planar f32, no conversion), chunk by chunk
    n_blocks = (total_frames + B - 1) // B
    pinned = wb.PinnedArray((2, args.chunk * B))
    eng.stop()
    eng.set_playhead(0.0)
    eng.play()
    eng.render(min(args.chunk, n_blocks), want_peaks=False, out=pinned.array[:, :min(args.chunk, n_blocks) * B])
    t_render = 1e9
    for _ in range(3):
        eng.stop()
        eng.set_playhead(0.0)
        eng.play()
        t0 = time.perf_counter()
        done = 0
        while done < n_blocks:
            n = min(args.chunk, n_blocks - done)
            eng.render(n, want_peaks=False, out=pinned.array[:, :n * B])
            done += n
        t_render = min(t_render, time.perf_counter() - t0)
    eng.stop()
    # parity spot check: the first callbacks of the export against the reference's process loop + convert_f32_to_interleaved_*
    ref.play()
    r, _ = ref.process(args.check_blocks)
    planar = np.ascontiguousarray(r.transpose(1, 0, 2).reshape(2, args.check_blocks * B))
    want = o.interleave(ref.kind, planar, code)
    same = bool(np.array_equal(got[:want.size], want))
    tf = N * total_frames
    print("bounce: %d tracks x %.1f min -> %s, chunks of %d callbacks" % (N, args.minutes, args.fmt, args.chunk))
    print("  (best of 3 runs each; the bounce copies every chunk from its page-locked slot into a pageable %d MB destination)" % (out.size >> 20))
    print("  bounce  %.3f s  %.3e track-frames/s  (%.0fx realtime)" % (t_bounce, tf / t_bounce, args.minutes * 60.0 / t_bounce))
    print("  render  %.3f s  %.3e track-frames/s  (same audio, planar f32 bus into page-locked channels)" % (t_render, tf / t_render))
    print("  bounce / render rate = %.3f" % (t_render / t_bounce))
    print("  first %d callbacks == reference process loop + convert_f32_to_interleaved: %s" % (args.check_blocks, same))
    if not same:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
