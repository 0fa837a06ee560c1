"""Realtime mode (development tool): per-callback latency of wbx::Engine::render(1) — one 512-frame callback of
N stereo tracks through the C ABI with host buffers (schedule + H2D + mix + D2H + sync) — p50 / p99 in µs."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=1024)
    ap.add_argument("--calls", type=int, default=400)
    ap.add_argument("--mode", default="auto")
    args = ap.parse_args()
    import whitebox_b200 as wb
    mode = {"auto": wb.SUM_AUTO, "exact": wb.SUM_EXACT, "tree": wb.SUM_TREE}[args.mode]
    eng = wb.Engine(2, 512, 48000, 120.0, device=0, sum_mode=mode)
    rng = np.random.default_rng(3)
    frames = (args.calls + 64) * 512
    base = ((rng.random((2, frames), dtype=np.float32) * 2 - 1) * np.float32(0.5 / np.sqrt(args.tracks))).astype(np.float32)
    for t in range(args.tracks):
        eng.add_track(-6.0 - (t % 7), -1.0 + 0.2 * (t % 11), False)
        sid = eng.add_sample(np.roll(base, 13 * t, axis=1), 48000)
        eng.add_clip(t, sid, 0.0, 1e9, 0.0, 1.0, 0.7)
    eng.play()
    out = wb.PinnedArray((2, 512))
    for _ in range(20):
        eng.render(1, want_peaks=False, out=out.array)
    ts = []
    for _ in range(args.calls):
        t0 = time.perf_counter()
        eng.render(1, want_peaks=False, out=out.array)
        ts.append((time.perf_counter() - t0) * 1e6)
    ts = np.array(ts)
    print("realtime callback, %d tracks, mode=%s, kernel=%s: p50 %.1f us  p99 %.1f us  max %.1f us  (budget 10667 us)" %
          (args.tracks, args.mode, eng.dev.last_kernel(), np.percentile(ts, 50), np.percentile(ts, 99), ts.max()), flush=True)


if __name__ == "__main__":
    main()
