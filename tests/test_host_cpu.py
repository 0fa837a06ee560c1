"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol the headers
declare, refuses to run without a device (no CPU fallback), and the host scheduler (include/wbx_engine.hpp,
whitebox_b200/csrc/wbx_host.cpp) emits segment tables whose documented semantics (tests/segment_render.py,
numpy) reproduce the reference's golden vectors bit for bit."""
import ctypes
import os
import re

import numpy as np
import pytest

import scenarios as sc
import segment_render as sr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def wb():
    import __graft_entry__ as ge
    ge.build_library()
    import whitebox_b200
    return whitebox_b200


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wbxh?_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(wb):
    L = wb.lib()
    for header, listed in (("wbx.h", wb.WBX_SYMBOLS), ("wbx_host.h", wb.WBXH_SYMBOLS)):
        declared = _declared(header)
        assert declared, header
        assert sorted(listed) == declared, "python symbol list out of date for " + header
        for name in declared:
            assert hasattr(L, name), "%s declared in %s but not exported" % (name, header)
    assert L.wbx_abi_version() == 3


def test_segment_struct_layout(wb):
    assert ctypes.sizeof(wb.Segment) == 80 and wb.SEGMENT_DTYPE.itemsize == 80
    for f in ("track", "block", "n_blocks", "dst_offset", "length", "sample_id", "src_pos", "speed", "gain", "flags",
              "clip_frame", "fade_in_frames", "fade_out_frames", "clip_len_frames"):
        assert getattr(wb.Segment, f).offset == wb.SEGMENT_DTYPE.fields[f][1]


def test_no_cpu_fallback(wb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(wb.WbxError):
        wb.DeviceEngine(0)
    with pytest.raises(wb.WbxError):
        wb.Engine(2, 512, 48000, 120.0, device=0)
    eng = wb.Engine(2, 512, 48000, 120.0, device=-1)  # scheduling-only engine
    eng.add_track(0.0, 0.0, False)
    with pytest.raises(wb.WbxError):  # ... cannot render
        eng.render(1)


def test_host_scalar_math_matches_golden(wb, golden_dir):
    g = np.load(os.path.join(golden_dir, "scalars.npz"))
    pc = np.array([wb.panning_coefs(float(p)) for p in g["pans"]], np.float32)
    dl = np.array([wb.db_to_linear(float(d)) for d in g["dbs"]], np.float32)
    assert np.array_equal(pc.view(np.uint32), g["pan_coefs"].view(np.uint32))
    assert np.array_equal(dl.view(np.uint32), g["db_lin"].view(np.uint32))


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


@pytest.mark.parametrize("batched", [True, False])
@pytest.mark.parametrize("name", sorted(sc.ALL))
def test_host_scheduler_matches_golden(wb, golden_dir, name, batched):
    gold = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    res = sc.ALL[name](lambda C, B, r, bpm: sr.ScheduleOnlyEngine(C, B, r, bpm, batched))
    for k in gold:
        assert _same(res[k], gold[k]), "%s: %s" % (name, k)


@pytest.mark.parametrize("seed", range(4))
def test_host_scheduler_matches_golden_fuzz(wb, golden_dir, seed):
    gold = dict(np.load(os.path.join(golden_dir, "fuzz%d.npz" % seed)))
    res = sc.fuzz(lambda C, B, r, bpm: sr.ScheduleOnlyEngine(C, B, r, bpm, True), seed)
    for k in gold:
        assert _same(res[k], gold[k]), k


def test_host_scheduler_matches_port_fuzz(wb):
    """More random sessions against the C restatement (skipping the reference-UB seeds)."""
    import oracle_api as o
    L = o.lib("port")
    L.wbo_ub_count.restype = ctypes.c_uint64
    ran = 0
    for seed in range(4, 60):
        before = L.wbo_ub_count()
        ref = sc.fuzz(lambda C, B, r, bpm: o.Session("port", C, B, r, bpm), seed)
        if L.wbo_ub_count() != before:
            continue
        res = sc.fuzz(lambda C, B, r, bpm: sr.ScheduleOnlyEngine(C, B, r, bpm, seed % 2 == 0), seed)
        for k in ref:
            assert _same(res[k], ref[k]), "fuzz%d: %s" % (seed, k)
        ran += 1
    assert ran > 35


def test_host_scheduler_matches_port_overlap_fuzz(wb):
    """The product's clip-list editing (overlapping adds) + scheduler against the C restatement on random sessions."""
    import oracle_api as o
    L = o.lib("port")
    L.wbo_ub_count.restype = ctypes.c_uint64
    ran = 0
    for seed in range(80):
        before = L.wbo_ub_count()
        ref = sc.fuzz_overlap(lambda C, B, r, bpm: o.Session("port", C, B, r, bpm), seed)
        if L.wbo_ub_count() != before:
            continue
        res = sc.fuzz_overlap(lambda C, B, r, bpm: sr.ScheduleOnlyEngine(C, B, r, bpm, seed % 2 == 0), seed)
        for k in ref:
            assert _same(res[k], ref[k]), "fuzz_overlap%d: %s" % (seed, k)
        ran += 1
    assert ran > 50


def test_host_scheduler_matches_port_edit_fuzz(wb):
    """The product's clip editing (move / resize / delete / duplicate, also on the playing clip) + scheduler against the
    C restatement on random sessions."""
    import oracle_api as o
    L = o.lib("port")
    L.wbo_ub_count.restype = ctypes.c_uint64
    ran = 0
    for seed in range(100):
        before = L.wbo_ub_count()
        ref = sc.fuzz_edits(lambda C, B, r, bpm: o.Session("port", C, B, r, bpm), seed)
        if L.wbo_ub_count() != before:
            continue
        res = sc.fuzz_edits(lambda C, B, r, bpm: sr.ScheduleOnlyEngine(C, B, r, bpm, seed % 2 == 0), seed)
        for k in ref:
            assert _same(res[k], ref[k]), "fuzz_edits%d: %s" % (seed, k)
        ran += 1
    assert ran > 60


@pytest.mark.parametrize("batched", [True, False])
def test_fade_extension_host_matches_port(wb, batched):
    """EXTENSION (parity unpinned w.r.t. whitebox): the product's scheduler + documented segment semantics
    reproduce the C port's fade specification bit for bit."""
    import oracle_api as o
    ref = sc.fades(lambda C, B, r, bpm: o.Session("port", C, B, r, bpm))
    res = sc.fades(lambda C, B, r, bpm: sr.ScheduleOnlyEngine(C, B, r, bpm, batched))
    for k in ref:
        assert _same(res[k], ref[k]), k


def test_runs_are_merged(wb):
    """A clip that plays through many callbacks is ONE segment with n_blocks = run length."""
    eng = wb.Engine(2, 512, 48000, 120.0, device=-1)
    for t in range(3):
        eng.add_track(-3.0, 0.0, False)
        sid = eng.add_sample(np.zeros((2, 512 * 40), np.float32), 48000)
        eng.add_clip(t, sid, 0.0, 100.0, 0.0, 1.0, 1.0)
    eng.play()
    segs, gains = eng.schedule(32)
    assert len(segs) == 3 and all(segs["n_blocks"] == 32) and all(segs["length"] == 512)
    assert gains.shape == (3, 2)
    segs, _ = eng.schedule(4)  # continues where it left off
    assert len(segs) == 3 and all(segs["src_pos"] == 32 * 512.0)


def test_no_fma_contraction_in_mix_kernels(wb):
    """Parity depends on every multiply and add being separately rounded. nvcc -fmad=false covers scalar code,
    but ptxas (12.9) still contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 when the product has a single use,
    so lint the SASS: inside mix_kernel the only fused ops allowed are the intended `a * -1 + b` (= b - a) and
    `prod * one + a` with a run-time 1.0f (= prod + a), one of each per lerp site."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", wb.LIB_PATH], capture_output=True, text=True).stdout
    fn, bad, seen = None, [], 0
    neg1, other, fmul2 = {}, {}, {}
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
            continue
        if fn and "mix_kernel" in fn:
            seen += 1
            # DFMA is not linted: the f64 path uses only __dadd_rn/__dmul_rn/__ddiv_rn and the correctly rounded
            # division itself expands to DFMA Newton steps (fade envelope, K_FADE)
            if " FMUL2 " in line or " FMUL2." in line:
                fmul2[fn] = fmul2.get(fn, 0) + 1
            if " FFMA2 " in line or " FFMA2." in line:
                if ", -1, " in line:
                    neg1[fn] = neg1.get(fn, 0) + 1  # b - a
                else:
                    other[fn] = other.get(fn, 0) + 1  # prod * one + a with the run-time 1.0f (consume_lin_t)
            elif " FFMA " in line or " FFMA." in line:
                pass  # integer-division helpers (work-item decode) use scalar FFMA on non-audio values
    assert seen > 1000, "mix_kernel SASS not found"
    # every 2-tap lerp site holds exactly one `a * -1 + b`, one `prod * one + a` and three packed multiplies (fx * df,
    # * clip gain, * track gain); a contraction anywhere (lerp, fast path, unity path) adds an FFMA2 or removes an FMUL2
    # (the EXT = true builds also hold the polyphase path: 16 packed tap FMAs per frame site, which ARE the specification's
    # fused multiply-adds — there the surplus must be a multiple of 16)
    for f in set(neg1) | set(other):
        n, m = neg1.get(f, 0), other.get(f, 0)
        ext = "Lb1EEEv" in f
        ok = (m >= n and (m - n) % 16 == 0) if ext else (m == n)
        if not ok or fmul2.get(f, 0) < 3 * n:
            bad.append((f[:50], n, m, fmul2.get(f, 0)))
    assert neg1, "lerp sites not found"
    assert not bad, "fused multiply-adds on the audio path (function, b-a, other FFMA2, FMUL2): %s" % bad[:5]


def build_dropin_demo(tmp):
    import shutil
    import subprocess
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    exe = os.path.join(tmp, "dropin_demo")
    lib_dir = os.path.join(ROOT, "whitebox_b200")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "dropin_demo.cpp"), "-L" + lib_dir, "-lwbx",
                    "-Wl,-rpath," + lib_dir, "-o", exe], check=True)
    return exe


def test_cpp_adaptor_builds_and_refuses_without_device(wb, tmp_path):
    """The C++ adaptor (include/wbx_engine.hpp) links against libwbx.so from plain g++; without a GPU the demo
    exits with the no-device status instead of rendering on the CPU."""
    import subprocess
    import torch
    exe = build_dropin_demo(str(tmp_path))
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU path" in r.stderr


def test_position_recurrence_closed_form_is_exact():
    """wbx::advance_rounded == the sampler's step-by-step `off = off + adv` (dsp/sampler.cpp:103,209), one rounding per
    callback, bit for bit: random starts / advances / limits, exact ties (adv/ulp halfway), binade crossings, tiny and
    huge values. Python floats are IEEE doubles, so the naive loop below is the reference's arithmetic."""
    import ctypes as C
    import whitebox_b200 as wb
    L = wb.lib()
    rng = np.random.default_rng(0)

    def naive(x, a, n, lim):
        s = 0
        while s < n and x < lim:
            x = x + a
            s += 1
        return x, s

    cases = [(3.0, 512 * 44100 / 48000, 4096, 1e300), (0.0, 470.4, 4096, 1e6), (1.0, 2.0 ** -30, 3000, 1e300),
             (2.0 ** 40, 0.75, 2500, 1e300), (5.0, 0.0, 10, 1e300)]
    for i in range(1500):
        kind = i % 6
        if kind == 0:
            x, a = float(rng.integers(0, 1 << 20)), 512 * 44100 / 48000
        elif kind == 1:
            x, a = rng.random() * 1e6, rng.random() * 1000 + 1e-3
        elif kind == 2:
            x, a = float(rng.integers(0, 1000)) + 0.25, float(rng.integers(1, 2000)) + 0.5
        elif kind == 3:
            x, a = rng.random() * 10, 512 * rng.random() * 4
        elif kind == 4:
            x, a = float(2 ** int(rng.integers(0, 30))), 2.0 ** -int(rng.integers(0, 40)) * int(rng.integers(1, 1 << 20))
        else:
            x, a = rng.random() * 1e-3, rng.random() * 1e-2
        n = int(rng.integers(1, 5000))
        lim = x + a * rng.random() * 6000 if i % 3 else 1e300
        cases.append((float(x), float(a), n, float(lim)))
    for x, a, n, lim in cases:
        want_x, want_s = naive(x, a, n, lim)
        off = C.c_double(x)
        got_s = L.wbxh_advance_rounded(C.byref(off), a, n, lim)
        assert (got_s, off.value) == (want_s, want_x), (x, a, n, lim)


def test_bench_reference_arm_prints_one_json_line():
    """The driver parses stdout of `bench.py --impl reference`: exactly one line, JSON, with the contract's keys; everything
    else (progress, library banners) goes to stderr."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-blocks", "64", "--tracks", "32"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


# ---- threading contract of the host engine (include/wbx_engine.hpp) under the sanitizers ---------------------------

def _build_stress(tmp_path, san):
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not shutil.which("g++"):
        pytest.skip("no g++")
    exe = str(tmp_path / ("thread_stress_" + san.split(",")[0]))
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=" + san, "-ffp-contract=off", "-D__align__(n)=alignas(n)",
           "-I" + os.path.join(root, "include"), os.path.join(root, "tests", "cpp", "thread_stress.cpp"),
           os.path.join(root, "whitebox_b200", "csrc", "wbx_host.cpp"), "-o", exe, "-lpthread"]
    if san != "thread":
        cmd.insert(5, "-fno-sanitize-recover=undefined")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("sanitizer runtime not available: " + r.stderr[:200])
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.parametrize("san", ["thread", "address,undefined"])
def test_threading_contract_under_sanitizers(tmp_path, san):
    """SPSC parameter ring (core/queue.h:142-196), atomic VU levels (vu_meter.h:17-40), editor lock across callbacks and
    edits (engine.cpp:1587,1651), load meter (timing.h:54-67): a UI thread hammering set_volume / set_pan / set_mute /
    add_audio_clip / set_bpm / set_track_effects while the audio thread runs >= 10^4 callbacks is ThreadSanitizer-clean
    (and ASan/UBSan-clean) on wbx_host.cpp, and every callback uses a volume of some serialised schedule."""
    import subprocess
    exe = _build_stress(tmp_path, san)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66", ASAN_OPTIONS="detect_leaks=1", UBSAN_OPTIONS="halt_on_error=1")
    r = subprocess.run([exe, "10000", "20000"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "errors 0" in r.stdout
