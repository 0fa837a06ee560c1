#!/usr/bin/env python
"""bench.py — the mixing hot path on B200 (BASELINE.json metric: mixed stereo samples/sec at N tracks; achieved
HBM GB/s vs roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfgN] [--sub 0|1]

A "step" renders `blocks_per_step` consecutive Engine::process callbacks (512 frames each) of `tracks_per_gpu` stereo
48 kHz f32 tracks per GPU in one device launch. The contract line is BASELINE cfg 2 (1024 tracks, gain/pan + bus sum;
fade = 0, the reference has none); the same line carries short runs of cfg 3 / cfg 4 / cfg 5 under "configs".
  value    whole-job stereo track-frames mixed per second, sources + schedule resident in HBM, timed with CUDA
           events on the launching stream (max over ranks). The K-step loop is repeated until the timed region is
           at least --min-seconds long (sustained clocks), ms_per_step is the mean over all timed steps.
  e2e      the same metric through the host engine API (wbx::Engine::render via the C ABI) with HOST buffers:
           host clip scheduling, H2D of the segment table + gains, schedule expansion, mix, D2H of the clamped
           bus (into page-locked host channels) and of the per-track VU levels all inside the timed region.
           Source samples are resident engine state (uploaded at load time, like wb::Sample objects in the
           reference); `e2e_cold` additionally counts uploading every source sample from host memory each step;
           `e2e_pageable` is e2e into plain (pageable, aligned_alloc-style) caller channels — what an unchanged
           wb::AudioBuffer (core/audio_buffer.h:34) gets.
  roofline achieved = algorithmic bytes (8 B per stereo track-frame) / mean mix-kernel duration (cfg 5: direct-form
           flops / step duration against the tensor peak).
  parity   after the timed region the first callbacks of ALL tracks x ranks are rendered again through the public
           API (the sharded peer-memory path at N > 1) and compared on rank 0 with the CPU checker (the compiled
           reference for cfg 2/3, the C spec for cfg 4, an f64 statement of the spec for cfg 5).
N > 1 (torchrun): tracks shard across ranks (weak scaling: tracks per GPU fixed; cfg 5 strong), each rank mixes its
shard unclamped and the partial buses are summed and clamped (engine.cpp:1627 after the sum) by the exchange fused into
the mix kernel over peer memory (include/wbx.h "sharded render"; --exchange nccl: one NCCL all-reduce instead).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK = 512
RATE = 48000
ALG_BYTES_PER_TRACK_FRAME = 8  # stereo f32 source frame read once (cfg 2)
SEGMENT_BYTES = 80             # sizeof(wbx_segment)
PARITY_TOL = 1e-5              # north_star: 1e-5 relative, stated on the block peak

WORKLOAD_TEXT = {
    "cfg2": "cfg2: %d stereo tracks/GPU, 48 kHz f32, gain/pan + bus sum (fade=0: the reference has none), 512-frame block",
    "cfg3": "cfg3: %d stereo tracks/GPU, 44.1 kHz f32 sources resampled to 48 kHz (2-tap linear, the reference's resampler) + mix, 512-frame block",
    "cfg3p": "cfg3p: %d stereo tracks/GPU, 44.1 kHz f32 sources resampled to 48 kHz with the 128-phase x 16-tap polyphase windowed sinc (BASELINE cfg 3's wording; extension, parity unpinned: the reference only has the linear resampler) + mix, 512-frame block",
    "cfg4": "cfg4: %d stereo tracks/GPU, 48 kHz f32, 4-band biquad EQ + compressor chain on every track (extension, parity unpinned) + mix; a step = chains (wbx_submit) + mix",
    "cfg5": "cfg5: %d stereo tracks/GPU, 48 kHz f32, %d-tap convolution reverb on every track (tensor-core path; extension, parity unpinned) + mix; a step = chains (wbx_submit) + mix",
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON result. Libraries write banners to fd 1 behind Python's back (NCCL
    prints its version there at communicator init), so fd 1 is pointed at stderr for the whole run and the result line
    goes to a private duplicate of the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def peaks_json():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d
    return {}


def hbm_peak():
    d = peaks_json()
    if "hbm_gbs" in d:
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload_config(wl, n_per_gpu, world, blocks, taps):
    """The static description of a workload: identical in our arm and in the reference arm."""
    text = WORKLOAD_TEXT[wl] % ((n_per_gpu, taps) if wl == "cfg5" else (n_per_gpu,))
    src_rate = 44100 if wl in ("cfg3", "cfg3p") else RATE
    gib = n_per_gpu * blocks * BLOCK * ALG_BYTES_PER_TRACK_FRAME * src_rate / RATE / 2**30
    return {"workload": text, "tracks_per_gpu": n_per_gpu, "total_tracks": n_per_gpu * world, "block_frames": BLOCK,
            "blocks_per_step": blocks, "sample_rate": RATE,
            "l2": "inputs larger than L2 (%.2f GiB streamed per step per GPU vs 126 MB)" % gib}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_POOL = {}
PIN_POOL = False


def source_frames(n_blocks, src_rate=RATE):
    return int((n_blocks + 4) * BLOCK * src_rate / RATE) + 64


def make_sources(n_tracks, n_blocks, seed, src_rate=RATE):
    """SURVEY §8(d) fixture at bench size: stereo f32 sources of (n_blocks+4)*512+64 frames, uniform(-1,1) *
    0.5/sqrt(N). Every track/channel is a different window of one 256 MiB MT-seeded random pool (generating
    16 GiB of fresh random numbers per run would dominate the bench's wall time); yields (t, [L, R]) views."""
    frames = source_frames(n_blocks, src_rate)
    pool_len = max(1 << 26, 2 * frames)
    key = (seed, pool_len, n_tracks)
    if key not in _POOL:
        _POOL.clear()
        rng = np.random.default_rng(seed)
        pool = rng.random(pool_len, dtype=np.float32)
        pool *= 2.0
        pool -= 1.0
        pool *= np.float32(0.5 / np.sqrt(n_tracks))
        if PIN_POOL:  # page-locked copy so that streaming sources from host runs at PCIe speed (e2e_cold)
            import whitebox_b200 as wb
            pinned = wb.PinnedArray((pool_len,))
            pinned.array[:] = pool
            pool = pinned.array
            _POOL["_keep"] = pinned
        _POOL[key] = pool
    pool = _POOL[key]
    span = pool_len - frames
    for t in range(n_tracks):
        offs = [((2 * t + c) * 7919 * 4099 + 12345 * c) % span for c in range(2)]
        yield t, [pool[o:o + frames] for o in offs]


# `ncu --set full` captures of the mix kernel at the bench shape (1024 tracks x 4096 callbacks), newest first
NCU_MIX_CAPTURES = {"cfg2": ["r02_ncu_full_mix_cfg2.txt", "r01b_ncu_full_mix_fpl16_K4096.txt", "r01_ncu_full_mix_fpl16_K4096.txt"],
                    "cfg3": ["r02_ncu_full_mix_cfg3.txt"]}


def ncu_traffic(wl, n_tracks, n_blocks):
    """(dram__bytes_read.sum + dram__bytes_write.sum of one mix-kernel launch, file) from the newest committed
    `ncu --set full` capture of this exact shape (profiles/), or (None, None) when no capture of the shape exists."""
    if (n_tracks, n_blocks) != (1024, 4096):
        return None, None
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    for name in NCU_MIX_CAPTURES.get(wl, []):
        try:
            rd = wr = None
            for ln in open(os.path.join(ROOT, "profiles", name)):
                f = ln.split()
                if ln.startswith("dram__bytes_read.sum"):
                    rd = float(f[2]) * scale[f[1]]
                if ln.startswith("dram__bytes_write.sum"):
                    wr = float(f[2]) * scale[f[1]]
            if rd is not None and wr is not None:
                return rd + wr, "profiles/" + name
        except Exception:
            continue
    return None, None


def track_params(t):
    return -6.0 - (t % 7), -1.0 + 0.2 * (t % 11), float(np.float32(0.5 + 0.001 * (t % 512)))


def cfg4_params(wb):
    return wb.effect_params(eq=((120.0, 4.0, 0.7), (800.0, -6.0, 1.2), (2500.0, 3.0, 2.0), (8000.0, 5.0, 0.7)),
                            threshold_db=-30.0, ratio_code=2, attack_ms=2.0, release_ms=60.0, makeup_db=3.0)


def cfg5_ir(taps):
    ir = (np.random.default_rng(2).standard_normal(taps) * np.exp(-np.arange(taps) / (taps / 6.0)) * 0.01).astype(np.float32)
    ir[0] = 1.0
    return ir


# ---------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU engine (oracle/_ref) or the C port, on the host cores
# ---------------------------------------------------------------------------------------------------------
class CpuEngines:
    """The reference CPU engine (oracle/_ref, else the C port) set up once for the bench workload.
    threads == 1 is the faithful single-threaded reference; threads > 1 builds that many independent engine
    instances with n_tracks/threads tracks each (the generous all-cores row)."""

    def __init__(self, kind, n_tracks, n_blocks, threads, seed=1234, src_rate=RATE):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_api as o
        self.n_tracks, self.n_blocks, self.threads = n_tracks, n_blocks, threads
        per = [n_tracks // threads + (1 if i < n_tracks % threads else 0) for i in range(threads)]
        self.sessions = []
        src = make_sources(n_tracks, n_blocks, seed, src_rate)
        for i in range(threads):
            s = o.Session(kind, 2, BLOCK, RATE, 120.0)
            for j in range(per[i]):
                t, x = next(src)
                vol, pan, gain = track_params(t)
                s.add_track(vol, pan, False)
                sid = s.add_sample(np.stack(x), src_rate)
                s.add_clip(j, sid, 0.0, 1e9, 0.0, 1.0, gain)
            self.sessions.append(s)

    def run(self):
        """One pass over n_blocks callbacks from beat 0 -> (track_frames_per_s, seconds)."""
        for s in self.sessions:
            s.stop()
            s.play()
            s.time_process(1)  # warm-up callback (consumes pending parameter messages)
        secs = [0.0] * self.threads

        def work(i):
            secs[i] = self.sessions[i].time_process(self.n_blocks)

        ths = [threading.Thread(target=work, args=(i,)) for i in range(self.threads)]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        wall = time.perf_counter() - t0
        return self.n_tracks * self.n_blocks * BLOCK / wall, wall

    def close(self):
        for s in self.sessions:
            s.close()
        self.sessions = []


def cpu_engine_run(kind, n_tracks, n_blocks, threads, seed=1234, src_rate=RATE):
    eng = CpuEngines(kind, n_tracks, n_blocks, threads, seed, src_rate)
    try:
        return eng.run()
    finally:
        eng.close()


def oracle_kind():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as o
    if o.have_ref():
        return "reference"
    if not o.have_port():
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
    return "port"


def run_reference(args):
    """The reference's own CPU implementation of the path (oracle/_ref when it was built, else the C port) on the box's
    host cores, on the same workload description (`config`) as our arm; each step is a bounded sample of it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = oracle_kind()
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 32))
    wl = args.workload if args.workload in ("cfg2", "cfg3") else "cfg2"
    n_per_gpu = args.tracks or 1024
    K = args.blocks or 4096
    n_tracks = n_per_gpu * args.gpus
    src_rate = 44100 if wl in ("cfg3", "cfg3p") else RATE
    blocks = max(64, args.ref_blocks // args.gpus)  # the bounded sample; keeps the host-memory footprint in check
    vals = []
    engines = CpuEngines(kind, n_tracks, blocks, threads, src_rate=src_rate)
    for i in range(args.warmup + args.steps):
        v, secs = engines.run()
        log("reference step %d: %.3e track-frames/s (%.2fs)" % (i, v, secs))
        if i >= args.warmup:
            vals.append((v, secs))
    engines.close()
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([s for _, s in vals])) * 1e3
    one, _ = cpu_engine_run(kind, min(n_tracks, 1024), max(8, blocks // 4), 1, src_rate=src_rate)
    out = {
        "impl": "reference", "metric": "mixed stereo samples/sec at N tracks", "value": value,
        "unit": "stereo track-frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(wl, n_per_gpu, args.gpus, K, args.taps),
        "cpu_baseline": {"value": value, "unit": "stereo track-frames/s", "cores": threads, "kind": kind,
                         "sample": "each step = the first %d of the workload's %d callbacks, all %d tracks, through Engine::process on %d "
                                   "independent engine instances (tracks split evenly, one thread each; the reference mix itself is "
                                   "single-threaded); single-thread (faithful) figure: %.3e" % (blocks, K, n_tracks, threads, one),
                         "single_thread_value": one},
        "e2e": {"value": value, "unit": "stereo track-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def parity_expected(ctx, wl, N, Kp, src_rate, taps, src_blocks):
    """Rank 0: what the CPU checker says the first Kp callbacks of ALL N * world tracks mix to -> (out [Kp][C][B], kind)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as o
    wb, world = ctx.wb, ctx.world
    need = int((Kp + 1) * BLOCK * src_rate / RATE) + 80
    if wl == "cfg5":
        # f64 statement of the spec (oracle/wb_oracle.c apply_reverb: y[n] = (float) sum_k h[k] x[n-k] in f64), evaluated
        # with an f64 FFT — the direct sum at 65536 taps is ~1e13 MACs for this check; differs from it by ~1e-13 relative
        ir = cfg5_ir(taps).astype(np.float64)
        n_out = Kp * BLOCK
        nfft = 1 << int(np.ceil(np.log2(n_out + taps)))
        H = np.fft.rfft(ir, nfft)
        bus = np.zeros((2, n_out), np.float64)
        for r in range(world):
            for t, x in make_sources(N, src_blocks, 1234 + r, src_rate):
                vol, pan, gain = track_params(r * N + t)
                vol_lin = wb.db_to_linear(vol - 3.0 * np.log2(world))
                pl, pr = wb.panning_coefs(pan)
                for c, pc in ((0, pl), (1, pr)):
                    xin = (x[c][:n_out] * np.float32(gain)).astype(np.float32).astype(np.float64)  # Sampler::stream: src * gain
                    y = np.fft.irfft(np.fft.rfft(xin, nfft) * H, nfft)[:n_out].astype(np.float32)  # chain output (f32)
                    bus[c] += (y * np.float32(np.float32(vol_lin) * np.float32(pc))).astype(np.float32)
        out = np.clip(bus, -1.0, 1.0).astype(np.float32)
        return np.ascontiguousarray(out.reshape(2, Kp, BLOCK).transpose(1, 0, 2)), "f64 FFT statement of the C spec (apply_reverb)"
    kind = "port" if wl in ("cfg4", "cfg3p") else oracle_kind()
    if kind == "port" and not o.have_port():
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
    s = o.Session(kind, 2, BLOCK, RATE, 120.0)
    fxp = cfg4_params(wb) if wl == "cfg4" else None
    if wl == "cfg3p":
        s.set_resampler(1)
    g = 0
    for r in range(world):
        for t, x in make_sources(N, src_blocks, 1234 + r, src_rate):
            vol, pan, gain = track_params(r * N + t)
            s.add_track(vol - 3.0 * np.log2(world), pan, False)
            sid = s.add_sample(np.stack([x[0][:need], x[1][:need]]), src_rate)
            s.add_clip(g, sid, 0.0, 1e9, 0.0, 1.0, gain)
            if fxp is not None:
                s.set_effects(g, fxp)
            g += 1
    s.play()
    out, _ = s.process(Kp)
    s.close()
    return out, ("compiled reference (oracle/_ref)" if kind == "reference" else "C port / spec (oracle/wb_oracle.c)")


def parity_verdict(got, exp, kind, exact_expected):
    """got / exp: [Kp][C][B] f32."""
    bit = bool(np.array_equal(got.view(np.uint32), exp.view(np.uint32)))
    peak = np.maximum(np.abs(exp).max(axis=(1, 2), keepdims=True), 1e-30)
    err = float((np.abs(got.astype(np.float64) - exp.astype(np.float64)) / peak).max())
    return {"kind": kind, "callbacks": int(got.shape[0]), "bit_exact": bit, "max_err_of_block_peak": err,
            "tolerance": PARITY_TOL, "ok": bool(bit or err <= PARITY_TOL), "bit_exact_expected": bool(exact_expected)}


def run_workload(ctx, wl, N, K, steps, warmup, min_seconds, main, args):
    """One workload on every rank -> the result dict on rank 0 (None elsewhere)."""
    torch, dist, wb, shard = ctx.torch, ctx.dist, ctx.wb, ctx.shard
    world, rank, local = ctx.world, ctx.rank, ctx.local
    taps = args.taps
    src_rate = 44100 if wl in ("cfg3", "cfg3p") else RATE
    resubmit = wl in ("cfg4", "cfg5")  # the effect chains run at submit: a step is submit + mix
    Kp = {"cfg2": 4, "cfg3": 4, "cfg3p": 4, "cfg4": 4, "cfg5": 160 if taps > 8192 else 16}[wl]  # callbacks of the parity render
    Kp = max(1, min(Kp, args.parity_blocks)) if args.parity_blocks else Kp
    Kmax = max(K, Kp)
    peak_gbs, peak_src = hbm_peak()
    global PIN_POOL
    cold = bool(args.cold) and main and wl == "cfg2" and world == 1
    PIN_POOL = cold

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- session: N tracks on this rank (global track index rank*N + t) --------------------------------
    t_setup = time.perf_counter()
    # cfg 5 is held to a tolerance by nature (the reverb is not bit-exact to its f64 spec), so its small bus sum may take
    # the re-associated tree order; every other workload is timed in the reference's exact track order
    eng = wb.Engine(2, BLOCK, RATE, 120.0, device=local,
                    sum_mode=wb.SUM_EXACT if (args.exact and wl != "cfg5") else wb.SUM_AUTO)
    stream = ctx.stream
    eng.dev.set_stream(stream.cuda_stream)
    if wl == "cfg3p":
        eng.set_resampler(1)  # polyphase quality mode (extension)
    host_sources = []
    for t, x in make_sources(N, Kmax, 1234 + rank, src_rate):
        vol, pan, gain = track_params(rank * N + t)
        eng.add_track(vol - 3.0 * np.log2(world), pan, False)  # keep the N*world-track bus inside +/-1
        sid = eng.add_sample_planar(x, src_rate)
        eng.add_clip(t, sid, 0.0, 1e9, 0.0, 1.0, gain)
        if cold and rank == 0:
            host_sources.append(x)

    def attach_chains(warm):
        """(Re)attach the workload's effect chains: also clears their running state / reverb histories. The host engine
        hands edited chains to the device at its next render; warm=True does one render now (so that the device-resident
        loop below, which re-submits the schedule itself, finds them) — the parity render attaches without it and so
        starts from silence like the checker."""
        if wl == "cfg4":  # 4-band EQ + compressor on every track (extension, "parity unpinned": include/wbx.h)
            fxp = cfg4_params(wb)
            for t in range(N):
                eng.set_effects(t, fxp)
        if wl == "cfg5":  # convolution reverb on every track (FFT path by default; WBX_FIR=tc: tensor cores)
            for t in range(N):
                eng.set_effects(t, wb.effect_params(reverb=True))
        if warm and wl in ("cfg4", "cfg5"):
            eng.stop()
            eng.play()
            eng.render(1, want_peaks=False, want_bus=(rank == 0 or exchange != "peer"))
            eng.stop()

    if wl == "cfg5":
        eng.set_impulse_response(cfg5_ir(taps))
    if rank == 0:
        log("setup: %s, %d tracks x %d blocks (%.2f GiB of sources per GPU) in %.1fs" %
            (wl, N, K, N * 2 * source_frames(Kmax, src_rate) * 4 / 2**30, time.perf_counter() - t_setup))

    track_frames_per_step = N * K * BLOCK  # per rank
    dev = eng.dev
    dev.set_track_count(N)
    flags = wb.MIX_NO_CLAMP if world > 1 else 0

    # ---- bus exchange across ranks: peer memory fused into the mix (default) or one NCCL all-reduce -----
    exchange = "none"
    if world > 1:
        exchange = "nccl" if args.exchange == "nccl" else "peer"
        if exchange == "peer":
            try:
                handle = dev.shard_init(rank, world, Kmax)
                handles = [None] * world
                dist.all_gather_object(handles, handle)
                dev.shard_connect_ipc(handles)
                ok = 1
            except Exception as ex:  # e.g. CUDA IPC not permitted in this container
                log("rank %d: peer-memory exchange unavailable (%s) - using the NCCL all-reduce" % (rank, ex))
                ok = 0
            t_ok = torch.tensor([ok], device="cuda")
            dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
            if int(t_ok.item()) == 0:
                dev.shard_close()
                exchange = "nccl"
    attach_chains(True)

    def mix_step(ev_pair=None):
        """One step's device work: mix this rank's tracks, sum the bus across ranks, clamp. ev_pair brackets the
        mix kernel alone (cfg 2/3: the roofline kernel) or the whole step (chains + mix)."""
        if ev_pair and resubmit:
            ev_pair[0].record(stream)
        if resubmit:  # render the tracks with a chain + run the chains (wbx_submit), then mix
            dev.submit_raw(segs_ptr, n_segs_dev, gains_ptr, K)
        if ev_pair and not resubmit:
            ev_pair[0].record(stream)
        if exchange == "peer" and ev_pair:  # phased form: the mix kernel can be bracketed alone (warm-up steps only)
            dev.mix_sharded(0)  # mix; tiles -> owners' exchange buffers over NVLink while mixing; arrival signal
            ev_pair[1].record(stream)
            dev.mix_sharded(1)  # wait for all ranks, owner reduce in rank order + clamp -> rank 0's master bus
            dev.mix_sharded(2)  # wait until every slice is in
        elif exchange == "peer":
            # the timed form: mix, then the whole exchange (signal, wait, owner reduce in rank order + clamp -> rank 0's
            # master bus, signal, wait) as ONE more kernel launch (wbx_mix_sharded)
            dev.mix_sharded()
        else:
            dev.mix(flags)
            if ev_pair:
                ev_pair[1].record(stream)
            if world > 1:
                ptr, n = dev.device_bus()
                dist.all_reduce(shard.bus_tensor(dev))  # one NCCL all-reduce of the partial buses (sum, f32)
                dev.clamp_device(ptr, n)

    # ---- (1) device-resident throughput: schedule submitted once, launches of the mix kernel -----------
    eng.play()
    segs, gains = eng.schedule(K)
    segs = np.ascontiguousarray(segs, dtype=wb.SEGMENT_DTYPE)
    gains = np.ascontiguousarray(gains, dtype=np.float32)
    segs_ptr, gains_ptr, n_segs_dev = segs.ctypes.data, gains.ctypes.data, len(segs)
    dev.submit(segs, gains, K)
    dev.synchronize()
    W = max(3, warmup)
    with torch.cuda.stream(stream):
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(W)]
        for i in range(W):
            if i == W - 1:
                w0.record(stream)
            mix_step(wev[i] if exchange == "peer" else None)
        w1.record(stream)
        barrier()
        est_ms = max(w0.elapsed_time(w1), 1e-3)
        # the K-step loop is repeated until the timed region lasts >= min_seconds (same count on every rank)
        reps_t = torch.tensor([max(1.0, np.ceil(min_seconds * 1e3 / (est_ms * steps)))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(reps_t, op=dist.ReduceOp.MAX)
        reps = int(min(reps_t.item(), 200))
        n_timed = reps * steps
        sampler = ClockSampler(local)
        if rank == 0 and main:
            sampler.start()
        launches0 = dev.launch_count()
        n_ev = min(n_timed, 64)  # the kernel alone is bracketed on the first n_ev timed steps
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_ev)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for i in range(n_timed):
            mix_step(ev[i] if (i < n_ev and exchange != "peer") else None)
        e1.record(stream)
        barrier()
        launches = (dev.launch_count() - launches0) / reps  # per K-step loop
        clocks = sampler.stop() if (rank == 0 and main) else None
    total_ms = e0.elapsed_time(e1)
    # the roofline kernel alone: bracketed on the first timed steps; with the peer exchange on the warm-up steps (the timed
    # steps then run the fused two-launch form, which leaves no place for an event between mix and exchange)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in (wev[1:] if exchange == "peer" else ev)]))
    kernel_name = dev.last_kernel()
    out_dev, _ = dev.fetch(False, want_bus=(rank == 0 or exchange != "peer"))

    # ---- (2) end to end through the host engine API with host buffers --------------------------------
    def e2e_step(cold_step, out):
        if cold_step:  # stream every source sample from (page-locked) host memory again, then render
            for t, x in enumerate(host_sources):
                dev.sample_update_planar(t, x)
        eng.stop()
        eng.play()
        if world == 1:  # the public call: host scheduling, H2D table, expand, mix, D2H bus + VU levels
            return eng.render(K, want_peaks=False, out=out)
        if exchange == "peer":  # the public call on every rank; rank 0 receives the master bus
            return eng.render(K, want_peaks=False, out=out if rank == 0 else None, want_bus=(rank == 0))
        segs2, gains2 = eng.schedule(K)
        dev.submit(segs2, gains2, K)
        mix_step()
        if rank == 0:
            dev.L.wbx_fetch(dev.h, wb._chan_ptrs(out), None)
            return out, dev.fetch_levels()
        dev.synchronize()
        return None, None

    pinned_out = wb.PinnedArray((2, Kmax * BLOCK))
    out_host = pinned_out.array[:, :K * BLOCK]
    par_host = pinned_out.array[:, :Kp * BLOCK]
    host_output = "page-locked channels written by the mix kernel" if world == 1 else "copied from rank 0's master bus"
    shm_path = None
    shared = None
    if exchange == "peer":
        # one host output buffer shared by all ranks (/dev/shm segment, registered with CUDA by every process): each owner
        # stores its reduced slice there over its own PCIe link, nobody copies the whole bus
        shm_path = "/dev/shm/wbx_bench_%s_%s" % (os.environ.get("MASTER_PORT", "0"), wl)
        ok = 1
        try:
            if rank == 0:
                np.memmap(shm_path, dtype=np.float32, mode="w+", shape=(2, Kmax * BLOCK)).flush()
        except Exception as ex:
            log("rank 0: cannot create %s (%s)" % (shm_path, ex))
        dist.barrier()
        try:
            shared = np.memmap(shm_path, dtype=np.float32, mode="r+", shape=(2, Kmax * BLOCK))
            if dev.L.wbx_host_register(shared.ctypes.data, shared.nbytes) != 0:
                raise RuntimeError("cudaHostRegister of the shared segment failed")
        except Exception as ex:
            log("rank %d: shared host output unavailable (%s) - rank 0 copies the master bus instead" % (rank, ex))
            ok = 0
        t_ok = torch.tensor([ok], device="cuda")
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        if int(t_ok.item()) == 1:
            dev.shard_set_host_output(shared)
            out_host = shared[:, :K * BLOCK]
            par_host = shared[:, :Kp * BLOCK]
            host_output = "shared page-locked segment, every owner rank stores its slice of the master bus into it"
    e2e_steps = steps if main else max(2, steps // 2)
    with torch.cuda.stream(stream):
        for _ in range(2):
            e2e_step(False, out_host)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            out_e2e, peaks_e2e = e2e_step(False, out_host)
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        same = None
        if rank == 0 and out_e2e is not None and not resubmit:  # chain state carries over between renders
            same = bool(np.array_equal(np.asarray(out_dev).view(np.uint32), np.asarray(out_e2e).view(np.uint32)))
        e2e_cold_s = e2e_page_s = e2e_reg_s = None
        if cold:
            e2e_step(True, out_host)
            barrier()
            t0 = time.perf_counter()
            for _ in range(max(1, steps // 4)):
                e2e_step(True, out_host)
            barrier()
            e2e_cold_s = (time.perf_counter() - t0) / max(1, steps // 4)
        if main and world == 1:  # plain pageable caller channels, as an unchanged wb::AudioBuffer has them
            page_out = np.zeros((2, K * BLOCK + 8), np.float32)[:, :K * BLOCK]
            for _ in range(2):
                e2e_step(False, page_out)
            barrier()
            t0 = time.perf_counter()
            for _ in range(max(2, steps // 2)):
                out_page, _ = e2e_step(False, page_out)
            barrier()
            e2e_page_s = (time.perf_counter() - t0) / max(2, steps // 2)
            if rank == 0 and same is not None:
                same = same and bool(np.array_equal(np.asarray(out_dev).view(np.uint32), np.asarray(out_page).view(np.uint32)))
            # ... and the same unchanged allocation page-locked in place once (wbx_host_register: two calls at start-up in
            # the reference, no change to AudioBuffer): the mix kernel then writes it directly, as in the e2e leg
            e2e_reg_s = None
            base = np.ascontiguousarray(page_out.base if page_out.base is not None else page_out)
            if dev.L.wbx_host_register(base.ctypes.data, base.nbytes) == 0:
                try:
                    for _ in range(2):
                        e2e_step(False, page_out)
                    barrier()
                    t0 = time.perf_counter()
                    for _ in range(max(2, steps // 2)):
                        out_reg, _ = e2e_step(False, page_out)
                    barrier()
                    e2e_reg_s = (time.perf_counter() - t0) / max(2, steps // 2)
                    if rank == 0 and same is not None:
                        same = same and bool(np.array_equal(np.asarray(out_dev).view(np.uint32), np.asarray(out_reg).view(np.uint32)))
                finally:
                    dev.synchronize()
                    dev.L.wbx_host_unregister(base.ctypes.data)

    # ---- (3) parity: the first Kp callbacks of every track of every rank, again, through the public API -----------
    with torch.cuda.stream(stream):
        if resubmit:
            attach_chains(False)  # from silence, like the checker
        eng.stop()
        eng.play()
        if world == 1 or exchange == "peer":
            par_out, _ = eng.render(Kp, want_peaks=False, out=par_host if rank == 0 or world == 1 else None,
                                    want_bus=(rank == 0 or world == 1))
        else:
            segs2, gains2 = eng.schedule(Kp)
            dev.submit(segs2, gains2, Kp)
            dev.mix(flags)
            ptr, n = dev.device_bus()
            dist.all_reduce(shard.bus_tensor(dev))
            dev.clamp_device(ptr, n)
            par_out = None
            if rank == 0:
                dev.L.wbx_fetch(dev.h, wb._chan_ptrs(par_host), None)
                par_out = par_host
            dev.synchronize()
        barrier()
    n_segs = len(segs)
    h2d = n_segs * SEGMENT_BYTES + N * 8
    d2h = 2 * K * BLOCK * 4 + N * 2 * 4  # clamped bus + per-track VU levels (reduced over callbacks on the device)

    # max over ranks
    t = torch.tensor([total_ms, kern_ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_ms, e2e_s = [float(v) for v in t.tolist()]

    res = None
    if rank == 0:
        got = np.ascontiguousarray(np.asarray(par_out).reshape(2, Kp, BLOCK).transpose(1, 0, 2))
        t0 = time.perf_counter()
        exp, pkind = parity_expected(ctx, wl, N, Kp, src_rate, taps, Kmax)
        exact_expected = world == 1 and bool(args.exact) and wl != "cfg5"
        parity = parity_verdict(got, exp, pkind, exact_expected)
        parity["tracks"] = N * world
        parity["path"] = ("wbx::Engine::render through the C ABI with host buffers" +
                          (", tracks sharded over %d ranks: %s" % (world, "peer-memory exchange fused into the mix + shared host output" if exchange == "peer" else "NCCL all-reduce") if world > 1 else ""))
        if not resubmit:  # the device-resident timed run produced the same first callbacks
            dv = np.ascontiguousarray(np.asarray(out_dev)[:, :Kp * BLOCK].reshape(2, Kp, BLOCK).transpose(1, 0, 2))
            parity["device_run_equals_parity_render"] = bool(np.array_equal(dv.view(np.uint32), got.view(np.uint32)))
        parity["checker_seconds"] = round(time.perf_counter() - t0, 2)
        log("%s parity vs %s: bit_exact=%s max_err_of_block_peak=%.3g (%d callbacks x %d tracks, %.1fs)" %
            (wl, pkind, parity["bit_exact"], parity["max_err_of_block_peak"], Kp, N * world, parity["checker_seconds"]))

        ms_per_step = total_ms / n_timed
        value = world * track_frames_per_step / (ms_per_step * 1e-3)
        alg_bytes = track_frames_per_step * ALG_BYTES_PER_TRACK_FRAME * src_rate / RATE  # each source sample read once
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(wl, N, K)
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                    "traffic": traffic, "peak_source": peak_src, "kernel_ms": kern_ms,
                    "traffic_source": (traffic_src + " (ncu --set full, bytes per launch)") if traffic else None,
                    "algorithmic_bytes_per_launch": alg_bytes}
        if wl == "cfg3p":
            roofline["note"] = ("quality mode, not the reference's resampler: 16 taps x 2 channels per frame with a per-frame gather of "
                                "the phase's coefficient row; bound by shared-memory wavefronts (ncu: profiles/"
                                "r02_ncu_full_mix_cfg3_polyphase.txt), not HBM")
        if wl == "cfg4":
            roofline["note"] = ("kernel_ms = the whole step (track render + effect chains + mix); algorithmic bytes = 8 B per "
                                "track-frame (SURVEY.md 8d)")
        if wl == "cfg5":
            pj = peaks_json()
            tf_peak = float(pj.get("bf16_tflops_sustained", pj.get("bf16_tflops", 1590.0)))
            flops = 2.0 * taps * 2 * track_frames_per_step  # direct-form count: 2 * taps per output sample and channel
            tf = flops / (kern_ms * 1e-3) / 1e12
            fir_path = int(wb.lib().wbx_fir_path(dev.h))
            roofline = {"bound": "tensor", "achieved": tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": tf / tf_peak,
                        "traffic": None, "kernel_ms": kern_ms,
                        "path": {0: "direct form, CUDA cores", 1: "direct form as a Toeplitz GEMM on tcgen05 tensor cores (fp16 2-term split)",
                                 2: "partitioned FFT convolution (overlap-save, f32 CUDA cores)"}.get(fir_path, str(fir_path)),
                        "peak_source": "MEASURED_PEAKS.json bf16 (sustained)" if pj else "fallback 1590 TFLOP/s"}
            if fir_path == 1:
                roofline["issued_bf16_tflops"] = tf * ctx.fir_split
                roofline["note"] = ("kernel_ms = the whole step (track render + reverb chain + mix); achieved = DIRECT-FORM flops "
                                    "(2 * taps per output sample and channel, SURVEY.md 8d) / step time; issued_bf16_tflops = x%d, the "
                                    "tensor work actually issued (split-precision products)" % ctx.fir_split)
            else:
                roofline["note"] = ("kernel_ms = the whole step (track render + reverb chain + mix); achieved = DIRECT-FORM flops "
                                    "(2 * taps per output sample and channel, SURVEY.md 8d: counted regardless of the algorithm) / step "
                                    "time. The FFT path performs ~80x fewer multiply-adds than the direct form, which is why the "
                                    "direct-form count can exceed the tensor peak; WBX_FIR=tc selects the tcgen05 direct-form path")
        e2e_value = world * track_frames_per_step / e2e_s
        res = {
            "metric": "mixed stereo samples/sec at N tracks", "value": value, "unit": "stereo track-frames/s",
            "n_gpus": world, "steps": steps, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong" if (wl == "cfg5" and not args.tracks) else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(wl, N, world, K, taps),
            "timed": {"steps_timed": n_timed, "repeats_of_steps_loop": reps, "timed_ms": total_ms,
                      "note": "the --steps loop is repeated back to back until the timed region is >= %.1f s; ms_per_step = mean" % min_seconds},
            "detail": {
                "out_frames_per_s": value / (N * world), "realtime_x": value / (N * world) / RATE,
                "kernel": kernel_name, "parallelism": ("tracks sharded x%d, %s" % (world, "bus exchange over peer memory fused into the mix kernel (tiles stored into the owner rank's buffer, flag barrier, owner reduce + clamp into rank 0)" if exchange == "peer" else "1 NCCL all-reduce of the bus")) if world > 1 else "1 GPU",
                "bus_exchange": exchange, "e2e_host_output": host_output, "e2e_equals_device_run": same,
            },
            "roofline": roofline,
            "parity": parity,
            "e2e": {"value": e2e_value, "unit": "stereo track-frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3,
                    "note": "wbx::Engine::render through the C ABI with host buffers: host clip scheduling + H2D segment table + schedule expansion (+ effect chains) + mix + the clamped bus into page-locked host channels (written by the mix kernel itself at N=1, by the owner ranks at N>1) + VU levels to the host; source samples resident (engine state, as wb::Sample in the reference)"},
            "gpu_launches": int(round(launches)),
        }
        if clocks is not None:
            res["clocks"] = clocks
        if e2e_page_s:
            res["e2e_pageable"] = {"value": track_frames_per_step / e2e_page_s, "unit": "stereo track-frames/s",
                                   "ms_per_step": e2e_page_s * 1e3, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                   "note": "as e2e, into plain pageable caller channels (an unchanged wb::AudioBuffer, core/audio_buffer.h:34)"}
        if e2e_reg_s:
            res["e2e_registered"] = {"value": track_frames_per_step / e2e_reg_s, "unit": "stereo track-frames/s",
                                     "ms_per_step": e2e_reg_s * 1e3, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                     "note": "as e2e_pageable after wbx_host_register of the same (unchanged) allocation"}
        if e2e_cold_s:
            src_bytes = N * 2 * source_frames(Kmax, src_rate) * 4
            res["e2e_cold"] = {"value": track_frames_per_step / e2e_cold_s, "unit": "stereo track-frames/s",
                               "h2d_bytes_per_step": h2d + src_bytes, "d2h_bytes_per_step": d2h,
                               "note": "as e2e, plus streaming every source sample from page-locked host memory over PCIe each step (wbx_sample_update)"}
        if main and world == 1 and wl in ("cfg2", "cfg3"):
            cpu_kind = oracle_kind()
            cpu_val, cpu_secs = cpu_engine_run(cpu_kind, N, args.cpu_blocks, 1, src_rate=src_rate)
            res["cpu_baseline"] = {"value": cpu_val, "unit": "stereo track-frames/s", "cores": 1, "kind": cpu_kind,
                                   "sample": "%d callbacks of the same %d-track workload through Engine::process, 1 thread (the reference mix is single-threaded), %.1fs" % (args.cpu_blocks, N, cpu_secs)}
    # ---- teardown ------------------------------------------------------------------------------------------
    if world > 1:
        dist.barrier()
    if exchange == "peer":
        dev.shard_set_host_output(None)
        dev.synchronize()
        if shared is not None:
            dev.L.wbx_host_unregister(shared.ctypes.data)
        dev.shard_close()
        if world > 1:
            dist.barrier()
        if shm_path and rank == 0 and os.path.exists(shm_path):
            os.remove(shm_path)
    del out_host, par_host, shared
    eng.close()
    del pinned_out
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build_library()
    import whitebox_b200 as wb
    from whitebox_b200 import shard

    ctx = Ctx()
    ctx.torch, ctx.dist, ctx.wb, ctx.shard = torch, dist, wb, shard
    ctx.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.rank = rank = int(os.environ.get("RANK", "0"))
    ctx.local = local = int(os.environ.get("LOCAL_RANK", "0"))
    ctx.fir_split = int(wb.lib().wbx_fir_split_factor())
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit("--gpus %d needs torchrun with %d ranks (WORLD_SIZE=%d)" % (args.gpus, args.gpus, world))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx.stream = torch.cuda.Stream()

    defaults = {"cfg2": (1024, 4096), "cfg3": (1024, 4096), "cfg3p": (1024, 4096), "cfg4": (512, 1024), "cfg5": (max(1, 256 // world), 64)}
    wl = args.workload
    N = args.tracks if args.tracks else defaults[wl][0]
    K = args.blocks if args.blocks else defaults[wl][1]
    res = run_workload(ctx, wl, N, K, args.steps, args.warmup, args.min_seconds, True, args)
    if wl == "cfg2" and args.sub:
        # the other BASELINE configs as short runs on the same box (same harness, fewer steps)
        sub = {}
        sub_shape = {"cfg3": (1024, 4096), "cfg3p": (1024, 4096), "cfg4": (512, 1024), "cfg5": (max(1, 256 // world), 64)}
        for swl in ("cfg3", "cfg3p", "cfg4", "cfg5"):
            try:
                r = run_workload(ctx, swl, sub_shape[swl][0], sub_shape[swl][1], args.sub_steps, 3, args.sub_seconds, False, args)
            except Exception as ex:  # a failing sub-run must not take the contract line with it
                log("sub-run %s failed: %r" % (swl, ex))
                r = {"error": repr(ex)} if rank == 0 else None
                if world > 1:
                    raise
            if rank == 0:
                keep = ("value", "unit", "ms_per_step", "scaling", "config", "timed", "detail", "roofline", "parity", "e2e", "gpu_launches", "error")
                sub[swl] = {k: r[k] for k in keep if k in r}
        if rank == 0:
            res["configs"] = sub
    if rank == 0:
        emit(res)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg3p", "cfg4", "cfg5"],
                    help="BASELINE.json configs[1] (the contract line, default) or configs[2..4] alone")
    ap.add_argument("--tracks", type=int, default=0, help="stereo tracks per GPU (0: 1024; cfg4 512; cfg5 256 / gpus = strong scaling)")
    ap.add_argument("--blocks", type=int, default=0, help="512-frame callbacks per step (0: 4096; cfg4 1024; cfg5 64)")
    ap.add_argument("--taps", type=int, default=65536, help="cfg5: taps of the impulse response")
    ap.add_argument("--ref-blocks", type=int, default=1024, help="callbacks per step of the reference arm's bounded sample (at 1 GPU)")
    ap.add_argument("--cpu-blocks", type=int, default=1024, help="callbacks of the cpu_baseline sample")
    ap.add_argument("--exact", type=int, default=1, help="1: bit-exact sequential track order, 0: auto")
    ap.add_argument("--cold", type=int, default=1, help="also measure e2e_cold (N=1 only)")
    ap.add_argument("--sub", type=int, default=1, help="cfg2 line: also run cfg3 / cfg4 / cfg5 briefly (\"configs\")")
    ap.add_argument("--sub-steps", type=int, default=5)
    ap.add_argument("--sub-seconds", type=float, default=0.25, help="minimum timed region of a sub-run")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="minimum timed region of the contract line")
    ap.add_argument("--parity-blocks", type=int, default=0, help="callbacks of the parity render (0: 4; cfg5 160)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: bus sum over peer memory fused into the mix kernel, or one NCCL all-reduce")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
