"""Kernel-level timing helper (development tool): N stereo tracks x K callbacks resident on the device, times
wbx_mix alone for several tile shapes / sum modes with CUDA events. Not the contract bench (see bench.py)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=1024)
    ap.add_argument("--blocks", type=int, default=512)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--fpl", default="16,8,4")
    ap.add_argument("--rate", type=int, default=48000, help="source sample rate (44100 -> linear resample)")
    ap.add_argument("--offset", type=int, default=0, help="clip start offset in frames (unaligned windows)")
    ap.add_argument("--tree", type=int, default=0)
    ap.add_argument("--poly", type=int, default=0, help="1: polyphase quality mode (with --rate != 48000)")
    ap.add_argument("--fx", type=int, default=0, help="1: every track carries the 4-band EQ + compressor chain (cfg 4)")
    ap.add_argument("--reverb", type=int, default=0, help="taps of a convolution reverb on every track (cfg 5)")
    args = ap.parse_args()
    import torch
    import whitebox_b200 as wb
    N, K, B = args.tracks, args.blocks, 512
    dev = wb.DeviceEngine(0)
    dev.configure(2, B, 48000)
    dev.set_track_count(N)
    speed = args.rate / 48000.0
    frames = int((K + 2) * B * speed) + 64 + args.offset
    rng = np.random.default_rng(1)
    base = ((rng.random((2, frames), dtype=np.float32) * 2 - 1) * np.float32(0.5 / np.sqrt(N))).astype(np.float32)
    segs = np.zeros(N, wb.SEGMENT_DTYPE)
    t0 = time.time()
    for t in range(N):
        sid = dev.sample_upload(np.roll(base, t * 17, axis=1), args.rate)
        segs[t] = (t, 0, K, 0, B, sid, float(args.offset), speed, 0.5 + 0.001 * (t % 512), 2 if args.poly else 0, 0.0, 0.0, 0.0, 0.0)
    gains = np.full((N, 2), 0.7, np.float32)
    print("setup %.1fs, %.2f GiB" % (time.time() - t0, N * 2 * frames * 4 / 2**30), flush=True)
    stream = torch.cuda.Stream()
    dev.set_stream(stream.cuda_stream)
    dev.set_sum_mode(wb.SUM_TREE if args.tree else wb.SUM_EXACT)
    if args.reverb:
        import ctypes as C
        ir = (np.random.default_rng(2).standard_normal(args.reverb) * np.exp(-np.arange(args.reverb) / (args.reverb / 6.0)) * 0.01).astype(np.float32)
        ir[0] = 1.0
        assert wb.lib().wbx_set_impulse_response(dev.h, ir.ctypes.data, ir.size) == 0
        p = wb.effect_params(reverb=True)
        fx = (C.c_uint8 * 256)()
        assert wb.lib().wbx_effects_design(C.byref(p), 48000, fx) == 0
        for t in range(N):
            assert wb.lib().wbx_set_track_effects(dev.h, t, fx) == 0
        with torch.cuda.stream(stream):
            dev.submit(segs, gains, K)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.iters):
                dev.submit(segs, gains, K)
            e1.record(stream)
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        macs = N * 2 * K * B * args.reverb
        print("reverb submit (%s), %d tracks x %d callbacks x %d taps: %.3f ms  %.3e MAC/s (%.1f TFLOP/s direct-form count)  %.2fx realtime" %
              (os.environ.get("WBX_FIR", "auto"), N, K, args.reverb, ms, macs / ms * 1e3, 2 * macs / ms * 1e3 / 1e12,
               K * B / 48000.0 / (ms * 1e-3)), flush=True)
    if args.fx:
        import ctypes as C
        p = wb.effect_params(eq=((120.0, 4.0, 0.7), (800.0, -6.0, 1.2), (2500.0, 3.0, 2.0), (8000.0, 5.0, 0.7)),
                             threshold_db=-30.0, ratio_code=2, attack_ms=2.0, release_ms=60.0, makeup_db=3.0)
        fx = (C.c_uint8 * 256)()
        assert wb.lib().wbx_effects_design(C.byref(p), 48000, fx) == 0
        for t in range(N):
            assert wb.lib().wbx_set_track_effects(dev.h, t, fx) == 0
        with torch.cuda.stream(stream):
            dev.submit(segs, gains, K)  # warm-up (allocations)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.iters):
                dev.submit(segs, gains, K)  # expand + render tracks + effect chains + cell patch
            e1.record(stream)
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        print("fx submit (expand + render_tracks + effects + patch), %d tracks x %d callbacks: %.3f ms  %.3e track-frames/s" %
              (N, K, ms, N * K * B / ms * 1e3), flush=True)
    dev.submit(segs, gains, K)
    bytes_alg = N * K * B * 8 * speed
    for fpl in args.fpl.split(","):
        os.environ["WBX_FPL"] = fpl
        with torch.cuda.stream(stream):
            for _ in range(2):
                dev.mix()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.iters):
                dev.mix()
            e1.record(stream)
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        print("fpl=%s %s: %.3f ms  %.1f GB/s  %.3e track-frames/s" % (fpl, dev.last_kernel(), ms, bytes_alg / ms / 1e6,
                                                                     N * K * B / ms * 1e3), flush=True)


if __name__ == "__main__":
    main()
