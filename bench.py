#!/usr/bin/env python
"""bench.py — the mixing hot path on B200 (BASELINE.json metric: mixed stereo samples/sec at N tracks; achieved
HBM GB/s vs roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--blocks B] [--tracks T]

A "step" renders `--blocks` consecutive Engine::process callbacks (512 frames each) of `--tracks` stereo 48 kHz
f32 tracks per GPU (BASELINE cfg 2: gain/pan + bus sum; fade = 0, the reference has none) in one device launch.
  value    whole-job stereo track-frames mixed per second, sources + schedule resident in HBM, timed with CUDA
           events on the launching stream (max over ranks).
  e2e      the same metric through the host engine API (wbx::Engine::render via the C ABI) with HOST buffers:
           host clip scheduling, H2D of the segment table + gains, schedule expansion, mix, D2H of the clamped
           bus (into page-locked host channels) and of the per-track VU levels all inside the timed region.
           Source samples are resident engine
           state (uploaded at load time, like wb::Sample objects in the reference); `e2e_cold` additionally
           counts uploading every source sample from host memory each step.
  roofline achieved = algorithmic bytes (8 B per stereo track-frame + cells) / mean mix-kernel duration.
N > 1 (torchrun): tracks shard across ranks (weak scaling: --tracks per GPU), each rank mixes its shard
unclamped and the partial buses are summed and clamped (engine.cpp:1627 after the sum) by the exchange fused into
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
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)", 1965.0


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


NCU_MIX_CAPTURES = ["r01b_ncu_full_mix_fpl16_K4096.txt", "r01_ncu_full_mix_fpl16_K4096.txt"]  # newest first


def ncu_traffic(n_tracks, n_blocks):
    """(dram__bytes_read.sum + dram__bytes_write.sum of one mix-kernel launch, file) from the newest committed
    `ncu --set full` capture of this exact shape (profiles/), or (None, None) when no capture of the shape exists."""
    if (n_tracks, n_blocks) != (1024, 4096):
        return None, None
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    for name in NCU_MIX_CAPTURES:
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = oracle_kind()
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 32))
    n_tracks = args.tracks * args.gpus
    blocks = max(64, args.ref_blocks // args.gpus)  # keeps the host-memory footprint of the sample bounded
    vals = []
    engines = CpuEngines(kind, n_tracks, blocks, threads)
    for i in range(args.warmup + args.steps):
        v, secs = engines.run()
        log("reference step %d: %.3e track-frames/s (%.2fs)" % (i, v, secs))
        if i >= args.warmup:
            vals.append((v, secs))
    engines.close()
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([s for _, s in vals])) * 1e3
    one, _ = cpu_engine_run(kind, min(n_tracks, 1024), max(8, blocks // 4), 1)
    out = {
        "impl": "reference", "metric": "mixed stereo samples/sec at N tracks", "value": value,
        "unit": "stereo track-frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "cfg2: %d stereo tracks, 48 kHz f32, gain/pan + bus sum (fade=0), 512-frame block" % n_tracks,
                   "tracks": n_tracks, "block_frames": BLOCK, "blocks_per_step": blocks,
                   "note": "reference CPU engine (Engine::process), %d independent engine instances, tracks split evenly" % threads},
        "cpu_baseline": {"value": value, "unit": "stereo track-frames/s", "cores": threads, "kind": kind,
                         "sample": "%d callbacks of %d tracks per step; single-thread (faithful) figure: %.3e" % (blocks, n_tracks, one),
                         "single_thread_value": one},
        "e2e": {"value": value, "unit": "stereo track-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build_library()
    import whitebox_b200 as wb
    from whitebox_b200 import shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit("--gpus %d needs torchrun with %d ranks (WORLD_SIZE=%d)" % (args.gpus, args.gpus, world))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- workload: BASELINE.json configs[1] by default; configs[2..4] selectable (profiles/, not the contract line) ----
    wl = args.workload
    defaults = {"cfg2": (1024, 4096), "cfg3": (1024, 4096), "cfg4": (512, 1024), "cfg5": (256 // world, 64)}[wl]
    N = args.tracks if args.tracks else defaults[0]
    K = args.blocks if args.blocks else defaults[1]
    src_rate = 44100 if wl == "cfg3" else RATE
    resubmit = wl in ("cfg4", "cfg5")  # the effect chains run at submit: a step is submit + mix
    hbm_peak, peak_src, sm_max = peaks_json()
    global PIN_POOL
    args.cold = args.cold if wl == "cfg2" else 0
    PIN_POOL = bool(args.cold) and world == 1

    # ---- session: N tracks on this rank (global track index rank*N + t) --------------------------------
    t_setup = time.perf_counter()
    eng = wb.Engine(2, BLOCK, RATE, 120.0, device=local, sum_mode=wb.SUM_EXACT if args.exact else wb.SUM_AUTO)
    stream = torch.cuda.Stream()
    eng.dev.set_stream(stream.cuda_stream)
    host_sources = []
    for t, x in make_sources(N, K, 1234 + rank, src_rate):
        vol, pan, gain = track_params(rank * N + t)
        eng.add_track(vol - 3.0 * np.log2(world), pan, False)  # keep the N*world-track bus inside +/-1
        sid = eng.add_sample_planar(x, src_rate)
        eng.add_clip(t, sid, 0.0, 1e9, 0.0, 1.0, gain)
        if args.cold and rank == 0:
            host_sources.append(x)
    if wl == "cfg4":  # 4-band EQ + compressor on every track (extension, "parity unpinned": include/wbx.h)
        fxp = wb.effect_params(eq=((120.0, 4.0, 0.7), (800.0, -6.0, 1.2), (2500.0, 3.0, 2.0), (8000.0, 5.0, 0.7)),
                               threshold_db=-30.0, ratio_code=2, attack_ms=2.0, release_ms=60.0, makeup_db=3.0)
        for t in range(N):
            eng.set_effects(t, fxp)
    if wl == "cfg5":  # one 65536-tap impulse response, convolution reverb on every track (tensor-core path)
        taps = args.taps
        ir = (np.random.default_rng(2).standard_normal(taps) * np.exp(-np.arange(taps) / (taps / 6.0)) * 0.01).astype(np.float32)
        ir[0] = 1.0
        eng.set_impulse_response(ir)
        for t in range(N):
            eng.set_effects(t, wb.effect_params(reverb=True))
    if wl in ("cfg4", "cfg5"):  # the host engine hands edited chains to the device at its next render: do one now
        eng.play()
        eng.render(1, want_peaks=False)
        eng.stop()
    if rank == 0:
        log("setup: %s, %d tracks x %d blocks (%.2f GiB of sources per GPU) in %.1fs" %
            (wl, N, K, N * 2 * source_frames(K, src_rate) * 4 / 2**30, time.perf_counter() - t_setup))

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
                handle = dev.shard_init(rank, world, K)
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

    def mix_step(ev_pair=None):
        """One step's device work: mix this rank's tracks, sum the bus across ranks, clamp. ev_pair brackets the
        mix kernel alone (for the roofline)."""
        if ev_pair and resubmit:
            ev_pair[0].record(stream)
        if resubmit:  # render the tracks with a chain + run the chains (wbx_submit), then mix
            dev.submit_raw(segs_ptr, n_segs_dev, gains_ptr, K)
        if ev_pair and not resubmit:
            ev_pair[0].record(stream)
        if exchange == "peer":
            dev.mix_sharded(0)  # mix; tiles -> owners' exchange buffers over NVLink while mixing; arrival signal
            if ev_pair:
                ev_pair[1].record(stream)
            dev.mix_sharded(1)  # wait for all ranks, owner reduce in rank order + clamp -> rank 0's master bus
            dev.mix_sharded(2)  # wait until every slice is in
        else:
            dev.mix(flags)
            if ev_pair:
                ev_pair[1].record(stream)
            if world > 1:
                ptr, n = dev.device_bus()
                dist.all_reduce(shard.bus_tensor(dev))  # one NCCL all-reduce of the partial buses (sum, f32)
                dev.clamp_device(ptr, n)

    # ---- (1) device-resident throughput: schedule submitted once, K launches of the mix kernel ---------
    eng.play()
    segs, gains = eng.schedule(K)
    segs = np.ascontiguousarray(segs, dtype=wb.SEGMENT_DTYPE)
    gains = np.ascontiguousarray(gains, dtype=np.float32)
    segs_ptr, gains_ptr, n_segs_dev = segs.ctypes.data, gains.ctypes.data, len(segs)
    dev.submit(segs, gains, K)
    dev.synchronize()
    with torch.cuda.stream(stream):
        for _ in range(max(3, args.warmup)):
            mix_step()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        launches0 = dev.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for i in range(args.steps):
            mix_step(ev[i])
        e1.record(stream)
        barrier()
        launches = dev.launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
    total_ms = e0.elapsed_time(e1)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    kernel_name = dev.last_kernel()
    out_dev, _ = dev.fetch(False, want_bus=(rank == 0 or exchange != "peer"))

    # ---- (2) end to end through the host engine API with host buffers --------------------------------
    def e2e_step(cold):
        if cold:  # stream every source sample from (page-locked) host memory again, then render
            for t, x in enumerate(host_sources):
                dev.sample_update_planar(t, x)
        eng.stop()
        eng.play()
        if world == 1:  # the public call: host scheduling, H2D table, expand, mix, D2H bus + VU levels
            return eng.render(K, want_peaks=False, out=out_host)
        if exchange == "peer":  # the public call on every rank; rank 0 receives the master bus
            return eng.render(K, want_peaks=False, out=out_host if rank == 0 else None, want_bus=(rank == 0))
        segs2, gains2 = eng.schedule(K)
        dev.submit(segs2, gains2, K)
        mix_step()
        if rank == 0:
            dev.L.wbx_fetch(dev.h, wb._chan_ptrs(pinned_out.array), None)
            return pinned_out.array, dev.fetch_levels()
        dev.synchronize()
        return None, None

    pinned_out = wb.PinnedArray((2, K * BLOCK))
    out_host = pinned_out.array
    host_output = "page-locked channels written by the mix kernel" if world == 1 else "copied from rank 0's master bus"
    shm_path = None
    if exchange == "peer":
        # one host output buffer shared by all ranks (/dev/shm segment, registered with CUDA by every process): each owner
        # stores its reduced slice there over its own PCIe link, nobody copies the whole bus
        shm_path = "/dev/shm/wbx_bench_%s" % os.environ.get("MASTER_PORT", "0")
        ok = 1
        try:
            if rank == 0:
                np.memmap(shm_path, dtype=np.float32, mode="w+", shape=(2, K * BLOCK)).flush()
        except Exception as ex:
            log("rank 0: cannot create %s (%s)" % (shm_path, ex))
        dist.barrier()
        try:
            shared = np.memmap(shm_path, dtype=np.float32, mode="r+", shape=(2, K * BLOCK))
            if dev.L.wbx_host_register(shared.ctypes.data, shared.nbytes) != 0:
                raise RuntimeError("cudaHostRegister of the shared segment failed")
        except Exception as ex:
            log("rank %d: shared host output unavailable (%s) - rank 0 copies the master bus instead" % (rank, ex))
            ok = 0
        t_ok = torch.tensor([ok], device="cuda")
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        if int(t_ok.item()) == 1:
            dev.shard_set_host_output(shared)
            out_host = shared
            host_output = "shared page-locked segment, every owner rank stores its slice of the master bus into it"
    with torch.cuda.stream(stream):
        for _ in range(2):
            e2e_step(False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out_e2e, peaks_e2e = e2e_step(False)
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e_cold_s = None
        if args.cold and world == 1:
            e2e_step(True)
            barrier()
            t0 = time.perf_counter()
            for _ in range(max(1, args.steps // 4)):
                e2e_step(True)
            barrier()
            e2e_cold_s = (time.perf_counter() - t0) / max(1, args.steps // 4)
    n_segs = len(segs)
    h2d = n_segs * 48 + N * 8
    d2h = 2 * K * BLOCK * 4 + N * 2 * 4  # clamped bus + per-track VU levels (reduced over callbacks on the device)

    # max over ranks
    t = torch.tensor([total_ms, kern_ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_ms, e2e_s = [float(v) for v in t.tolist()]

    if rank == 0:
        same = (bool(np.array_equal(out_dev.view(np.uint32), out_e2e.view(np.uint32)))
                if (out_e2e is not None and not resubmit) else None)  # chain state carries over between renders
        ms_per_step = total_ms / args.steps
        value = world * track_frames_per_step / (ms_per_step * 1e-3)
        alg_bytes = track_frames_per_step * ALG_BYTES_PER_TRACK_FRAME * src_rate / RATE  # each source sample read once
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(N, K) if wl == "cfg2" else (None, None)
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": traffic, "peak_source": peak_src, "kernel_ms": kern_ms,
                    "traffic_source": (traffic_src + " (ncu --set full, bytes per launch)") if traffic else None,
                    "algorithmic_bytes_per_launch": alg_bytes}
        if wl == "cfg4":
            roofline["note"] = ("kernel_ms = the whole step (render tracks + effect chains + mix); the chains are recurrences in time, "
                                "bound by dependent-FMA latency, not by HBM (DESIGN.md 5.5)")
        if wl == "cfg5":
            pj = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
            tf_peak = float(pj.get("bf16_tflops_sustained", pj.get("bf16_tflops", 1590.0)))
            flops = 2.0 * args.taps * 2 * track_frames_per_step  # direct-form count: 2 * taps per output sample and channel
            tf = flops / (kern_ms * 1e-3) / 1e12
            roofline = {"bound": "tensor", "achieved": tf * 6, "peak": tf_peak, "unit": "TFLOP/s", "frac": tf * 6 / tf_peak,
                        "traffic": None, "kernel_ms": kern_ms, "direct_form_tflops": tf,
                        "note": "kernel_ms = the whole step (render tracks + reverb chain + mix); achieved = direct-form flops x 6 "
                                "(3-term bf16 split of both operands, six products: the bf16 tensor work actually issued)",
                        "peak_source": "MEASURED_PEAKS.json bf16 (sustained)" if pj else "fallback 1590 TFLOP/s"}
        e2e_value = world * track_frames_per_step * args.steps / e2e_s
        cpu_kind = oracle_kind()
        cpu_blocks = args.cpu_blocks
        cpu_val, cpu_secs = (cpu_engine_run(cpu_kind, N, cpu_blocks, 1, src_rate=src_rate)
                             if world == 1 and wl in ("cfg2", "cfg3") else (None, None))
        res = {
            "metric": "mixed stereo samples/sec at N tracks", "value": value, "unit": "stereo track-frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong" if (wl == "cfg5" and not args.tracks) else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": {
                    "cfg2": "cfg2: %d stereo tracks/GPU, 48 kHz f32, gain/pan + bus sum (fade=0: the reference has none), 512-frame block",
                    "cfg3": "cfg3: %d stereo tracks/GPU, 44.1 kHz f32 sources resampled to 48 kHz (2-tap linear, the reference's resampler) + mix, 512-frame block",
                    "cfg4": "cfg4: %d stereo tracks/GPU, 48 kHz f32, 4-band biquad EQ + compressor chain on every track (extension, parity unpinned) + mix; a step = chains (wbx_submit) + mix",
                    "cfg5": "cfg5: %d stereo tracks/GPU, 48 kHz f32, " + str(args.taps) + "-tap convolution reverb on every track (tensor-core path; extension, parity unpinned) + mix; a step = chains (wbx_submit) + mix",
                }[wl] % N,
                "tracks_per_gpu": N, "total_tracks": N * world, "block_frames": BLOCK, "blocks_per_step": K,
                "out_frames_per_s": value / (N * world), "realtime_x": value / (N * world) / RATE,
                "l2": "inputs larger than L2 (%.2f GiB streamed per step per GPU vs 126 MB)" % (alg_bytes / 2**30),
                "kernel": kernel_name, "parallelism": ("tracks sharded x%d, %s" % (world, "bus exchange over peer memory fused into the mix kernel (tiles stored into the owner rank's buffer, flag barrier, owner reduce + clamp into rank 0)" if exchange == "peer" else "1 NCCL all-reduce of the bus")) if world > 1 else "1 GPU",
                "bus_exchange": exchange, "e2e_host_output": host_output,
                "e2e_equals_device_run": same,
            },
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": "stereo track-frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3,
                    "note": "wbx::Engine::render through the C ABI with host buffers: host clip scheduling + H2D segment table + schedule expansion (+ effect chains) + mix + the clamped bus into page-locked host channels (written by the mix kernel itself at N=1, copied from rank 0's master bus at N>1) + VU levels to the host; source samples resident (engine state, as wb::Sample in the reference)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if e2e_cold_s:
            src_bytes = N * 2 * source_frames(K, src_rate) * 4
            res["e2e_cold"] = {"value": track_frames_per_step / e2e_cold_s, "unit": "stereo track-frames/s",
                               "h2d_bytes_per_step": h2d + src_bytes, "d2h_bytes_per_step": d2h,
                               "note": "as e2e, plus streaming every source sample from page-locked host memory over PCIe each step (wbx_sample_update)"}
        if cpu_val:
            res["cpu_baseline"] = {"value": cpu_val, "unit": "stereo track-frames/s", "cores": 1, "kind": cpu_kind,
                                   "sample": "%d callbacks of the same %d-track workload through Engine::process, 1 thread (the reference mix is single-threaded), %.1fs" % (cpu_blocks, N, cpu_secs)}
        emit(res)
    if world > 1:
        dist.barrier()
        if shm_path and rank == 0 and os.path.exists(shm_path):
            os.remove(shm_path)
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"],
                    help="BASELINE.json configs[1] (the contract line, default) or configs[2..4] (profiles/)")
    ap.add_argument("--tracks", type=int, default=0, help="stereo tracks per GPU (0: 1024; cfg4 512; cfg5 256 / gpus = strong scaling)")
    ap.add_argument("--blocks", type=int, default=0, help="512-frame callbacks per step (0: 4096; cfg4 1024; cfg5 64)")
    ap.add_argument("--taps", type=int, default=65536, help="cfg5: taps of the impulse response")
    ap.add_argument("--ref-blocks", type=int, default=1024, help="callbacks per step of the reference arm (at 1 GPU)")
    ap.add_argument("--cpu-blocks", type=int, default=1024, help="callbacks of the cpu_baseline sample")
    ap.add_argument("--exact", type=int, default=1, help="1: bit-exact sequential track order, 0: auto")
    ap.add_argument("--cold", type=int, default=1, help="also measure e2e_cold (N=1 only)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: bus sum over peer memory fused into the mix kernel, or one NCCL all-reduce")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        args.tracks = args.tracks or 1024
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
