"""Development tool: where one render (default: the realtime callback, K = 1) spends its time — host scheduling, wbx_submit, wbx_mix,
fetch — each phase timed on the host with a stream synchronise after it (so phases do not overlap)."""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=1024)
    ap.add_argument("--calls", type=int, default=300)
    ap.add_argument("--blocks", type=int, default=1, help="callbacks per render (1 = realtime callback)")
    ap.add_argument("--restart", type=int, default=0, help="1: stop/play before every render (bounce from the start, as bench.py's e2e)")
    args = ap.parse_args()
    import whitebox_b200 as wb
    eng = wb.Engine(2, 512, 48000, 120.0, device=0)
    rng = np.random.default_rng(3)
    K = args.blocks
    frames = ((2 * args.calls + 64) * K if not args.restart else (K + 8)) * 512
    base = ((rng.random((2, frames), dtype=np.float32) * 2 - 1) * np.float32(0.5 / np.sqrt(args.tracks))).astype(np.float32)
    for t in range(args.tracks):
        eng.add_track(-6.0 - (t % 7), -1.0 + 0.2 * (t % 11), False)
        sid = eng.add_sample(np.roll(base, 13 * t, axis=1), 48000)
        eng.add_clip(t, sid, 0.0, 1e9, 0.0, 1.0, 0.7)
    eng.play()
    dev, L = eng.dev, eng.L
    dev.set_track_count(args.tracks)
    out = wb.PinnedArray((2, 512 * K))
    ptrs = wb._chan_ptrs(out.array)
    segs, cnt, gains = C.c_void_p(), C.c_uint32(), C.c_void_p()
    lv = np.zeros((args.tracks, 2), np.float32)
    T = {k: [] for k in ("schedule", "submit", "mix", "fetch+levels", "one_call")}
    for i in range(args.calls):
        t0 = time.perf_counter()
        if args.restart:
            eng.stop()
            eng.play()
            t0 = time.perf_counter()
        L.wbxh_schedule(eng.h, K, C.byref(segs), C.byref(cnt), C.byref(gains))
        t1 = time.perf_counter()
        assert L.wbx_submit(dev.h, segs, cnt.value, gains, K) == 0
        L.wbx_synchronize(dev.h)
        t2 = time.perf_counter()
        assert L.wbx_mix(dev.h, 0) == 0
        L.wbx_synchronize(dev.h)
        t3 = time.perf_counter()
        assert L.wbx_fetch(dev.h, ptrs, None) == 0
        assert L.wbx_fetch_levels(dev.h, lv.ctypes.data) == 0
        t4 = time.perf_counter()
        if args.restart:
            eng.stop()
            eng.play()
        L.wbxh_schedule(eng.h, K, C.byref(segs), C.byref(cnt), C.byref(gains))
        t5 = time.perf_counter()
        assert L.wbx_render_levels(dev.h, segs, cnt.value, gains, K, ptrs, None, lv.ctypes.data) == 0
        t6 = time.perf_counter()
        if i >= min(20, args.calls // 3):
            for k, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t6 - t5)):
                T[k].append(v * 1e6)
    print("K=%d, %d tracks, kernel %s" % (K, args.tracks, dev.last_kernel()))
    for k, v in T.items():
        print("  %-14s p50 %7.1f us  p99 %7.1f us" % (k, np.percentile(v, 50), np.percentile(v, 99)))


if __name__ == "__main__":
    main()
