"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against
  (a) the golden vectors the reference's own Engine::process produced (tests/golden/*.npz),
  (b) the CPU checkers (the C restatement; the compiled reference when oracle/_ref travelled) on larger seeded
      inputs, and
  (c) size-independent properties at BASELINE.json's full track count.
Bar: bit-exact for everything in WBX_SUM_EXACT mode (sequential track order, no FMA contraction — the f32/f64
operations are the reference's, in the reference's order); WBX_SUM_TREE re-associates the bus sum and is held
to |gpu - ref| <= 1e-5 * block peak (north_star's 1e-5 relative tolerance, stated on the block peak because a
per-sample relative bound is meaningless at zero crossings); VU peaks stay bit-exact."""
import ctypes
import os

import numpy as np
import pytest

import oracle_api as o
import scenarios as sc

pytestmark = pytest.mark.gpu

TREE_TOL = 1e-5


@pytest.fixture(scope="module")
def wb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as ge
    ge.build_library()
    import whitebox_b200
    whitebox_b200.DeviceEngine(0).close()  # fails loudly if the extension cannot run here
    return whitebox_b200


def same_bits(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def assert_exact(res, ref, what):
    for k in ref:
        if not same_bits(res[k], ref[k]):
            a, b = np.asarray(res[k]), np.asarray(ref[k])
            bad = np.argwhere(a != b) if a.shape == b.shape else None
            n = 0 if bad is None else len(bad)
            first = None if not n else (bad[0].tolist(), float(a[tuple(bad[0])]), float(b[tuple(bad[0])]))
            raise AssertionError("%s: %s differs (%d elements, first %s)" % (what, k, n, first))


def assert_tree(res, ref, what):
    for k in ref:
        if k.startswith("out"):
            peak = np.abs(ref[k]).max(axis=(1, 2), keepdims=True)
            err = np.abs(res[k].astype(np.float64) - ref[k].astype(np.float64))
            assert np.all(err <= TREE_TOL * np.maximum(peak, 1e-30)), "%s: tree-mode bus error %.3g of block peak" % (
                what, float((err / np.maximum(peak, 1e-30)).max()))
        else:
            assert same_bits(res[k], ref[k]), "%s: %s differs" % (what, k)


def gpu_engine(wb, batched=True, mode=None):
    mode = wb.SUM_EXACT if mode is None else mode
    return lambda C, B, r, bpm: wb.Engine(C, B, r, bpm, device=0, batched=batched, sum_mode=mode)


def cpu_engine():
    kind = "reference" if o.have_ref() else "port"
    return lambda C, B, r, bpm: o.Session(kind, C, B, r, bpm)


# ---- (a) golden vectors from the reference ---------------------------------------------------------------

@pytest.mark.parametrize("batched", [True, False])
@pytest.mark.parametrize("name", sorted(sc.ALL))
def test_golden_exact(wb, golden_dir, name, batched):
    gold = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    assert_exact(sc.ALL[name](gpu_engine(wb, batched)), gold, name)


@pytest.mark.parametrize("seed", range(4))
def test_golden_fuzz_exact(wb, golden_dir, seed):
    gold = dict(np.load(os.path.join(golden_dir, "fuzz%d.npz" % seed)))
    assert_exact(sc.fuzz(gpu_engine(wb, seed % 2 == 0), seed), gold, "fuzz%d" % seed)


@pytest.mark.parametrize("name", ["cfg1", "cfg2_small", "cfg3_small", "event_split", "int_formats", "hot_clamp"])
def test_golden_tree(wb, golden_dir, name, monkeypatch):
    monkeypatch.setenv("WBX_GROUPS", "3")
    gold = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    assert_tree(sc.ALL[name](gpu_engine(wb, True, wb.SUM_TREE)), gold, name)


@pytest.mark.parametrize("fpl", ["4", "8", "16"])
@pytest.mark.parametrize("name", ["cfg2_small", "cfg3_small", "event_split", "int_formats", "ragged"])
def test_golden_every_tile_shape(wb, golden_dir, name, fpl, monkeypatch):
    """Frame tiles of 128 / 256 / 512 frames (peaks combined with atomicMax when a block spans tiles)."""
    monkeypatch.setenv("WBX_FPL", fpl)
    gold = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    assert_exact(sc.ALL[name](gpu_engine(wb, True)), gold, "%s fpl=%s" % (name, fpl))


@pytest.mark.parametrize("batched", [True, False])
def test_fade_extension_vs_port(wb, batched):
    """EXTENSION, parity unpinned w.r.t. whitebox (the reference has no fades): CUDA == the C port's
    specification of the fade envelope, bit for bit; and a (0, 0) fade is the reference path."""
    ref = sc.fades(lambda C, B, r, bpm: o.Session("port", C, B, r, bpm))
    assert_exact(sc.fades(gpu_engine(wb, batched)), ref, "fades")


@pytest.mark.parametrize("batched", [True, False])
def test_polyphase_extension_vs_port(wb, batched):
    """EXTENSION, parity unpinned (BASELINE cfg 3 names a polyphase resampler; the reference only has the linear
    one): CUDA == the C port's 128-phase x 16-tap specification bit for bit, incl. fades and the effect path."""
    ref = sc.polyphase(lambda C, B, r, bpm: o.Session("port", C, B, r, bpm), wb.effect_params)
    assert_exact(sc.polyphase(gpu_engine(wb, batched), wb.effect_params), ref, "polyphase")


@pytest.mark.parametrize("batched", [True, False])
def test_effects_extension_vs_port(wb, batched):
    """EXTENSION, parity unpinned w.r.t. whitebox (BASELINE cfg 4: 4-band biquad EQ + compressor per track): CUDA ==
    the C port's specification bit for bit, including chain state carried across renders."""
    ref = sc.effects(lambda C, B, r, bpm: o.Session("port", C, B, r, bpm), wb.effect_params)
    assert_exact(sc.effects(gpu_engine(wb, batched), wb.effect_params), ref, "effects")


@pytest.mark.parametrize("shape", [
    dict(n_tracks=24, block=512, n_blocks=5),                  # whole chunks
    dict(n_tracks=12, block=256, n_blocks=7, chunks=[3, 1, 3]),  # half chunks, state carried across renders
    dict(n_tracks=12, block=100, n_blocks=9),                  # partial segment (100 = 6 * 16 + 4)
    dict(n_tracks=12, block=101, n_blocks=6),                  # odd block: partial look-ahead block, unaligned stores
    dict(n_tracks=8, block=1, n_blocks=40),                    # one frame per callback: the one-step form only
    dict(n_tracks=8, block=3, n_blocks=21),
    dict(n_tracks=8, block=513, n_blocks=4),                   # a second chunk of one frame
    dict(n_tracks=8, block=1030, n_blocks=3),                  # three chunks per callback, the last one ragged
    dict(n_tracks=12, block=512, n_blocks=4, out_channels=1),  # mono bus
    dict(n_tracks=12, block=512, n_blocks=4, loud=False),      # far above the threshold
])
def test_effects_time_parallel_shapes_vs_port(wb, shape):
    """EXTENSION, parity unpinned: fx_chain_kernel == the C specification of the time-parallel chain (oracle/wb_oracle.c
    apply_effects) bit for bit at every chunk shape."""
    ref = sc.effects_shapes(lambda C, B, r, bpm: o.Session("port", C, B, r, bpm), wb.effect_params, **shape)
    assert_exact(sc.effects_shapes(gpu_engine(wb, True), wb.effect_params, **shape), ref, "effects %r" % (shape,))


@pytest.mark.parametrize("n_tracks", [512, 4096])
def test_effects_bench_shape_vs_port(wb, n_tracks):
    """BASELINE cfg 4 at bench shape (512 tracks per GPU; 4096 = the whole session on one GPU): every track with the 4-band
    EQ + compressor, CUDA == the C spec bit for bit (bus and VU peaks), and a per-callback render equals the batched one."""
    shape = dict(n_tracks=n_tracks, block=512, n_blocks=4 if n_tracks <= 512 else 2)
    ref = sc.effects_shapes(lambda C, B, r, bpm: o.Session("port", C, B, r, bpm), wb.effect_params, **shape)
    assert_exact(sc.effects_shapes(gpu_engine(wb, True), wb.effect_params, **shape), ref, "effects bench shape %d" % n_tracks)
    if n_tracks <= 512:
        assert_exact(sc.effects_shapes(gpu_engine(wb, False), wb.effect_params, **shape), ref, "effects per callback")


FX_SHAPES = ["1,4,1,1", "2,4,1,1", "4,2,1,1", "4,2,2,1", "2,2,1,2", "2,2,2,2", "1,2,1,4", "1,2,2,4", "1,4,2,2"]


@pytest.mark.parametrize("fx_shape", FX_SHAPES)
def test_effects_every_kernel_shape_vs_port(wb, fx_shape, monkeypatch):
    """Every compiled instantiation of fx_chain_kernel (tracks per CTA, output warps, EQ warps, CTAs per SM — normally
    chosen from the session size), forced with WBX_FX_SHAPE: whole chunks, a ragged multi-chunk block with state carried
    across renders, and an odd block, each == the C spec bit for bit."""
    monkeypatch.setenv("WBX_FX_SHAPE", fx_shape)
    for shape in (dict(n_tracks=24, block=512, n_blocks=5), dict(n_tracks=10, block=1030, n_blocks=3, chunks=[1, 2]),
                  dict(n_tracks=9, block=101, n_blocks=6)):
        ref = sc.effects_shapes(lambda C, B, r, bpm: o.Session("port", C, B, r, bpm), wb.effect_params, **shape)
        assert_exact(sc.effects_shapes(gpu_engine(wb, True), wb.effect_params, **shape), ref, "effects %s %r" % (fx_shape, shape))


@pytest.mark.parametrize("mode", ["direct", "tc", "fft", "fft512", "fftauto"])
@pytest.mark.parametrize("taps", [1, 2, 777, 2048])
def test_reverb_extension_vs_port(wb, taps, mode, monkeypatch):
    """EXTENSION, parity unpinned (BASELINE cfg 5 at test size): the convolution reverb — direct form on the CUDA
    cores, the tcgen05 tensor-core Toeplitz GEMM (2-term fp16 split) and the partitioned FFT convolution (2048- and
    512-tap partitions) — against the port's f64-accumulated specification, within 1e-5 of the block peak (bus and VU
    peaks), history carried across renders."""
    _fir_mode(monkeypatch, mode)
    ref = sc.reverb(lambda C, B, r, bpm: o.Session("port", C, B, r, bpm), wb.effect_params, taps)
    res = sc.reverb(gpu_engine(wb, True), wb.effect_params, taps)
    peak = np.abs(ref["out"]).max(axis=(1, 2), keepdims=True)
    err = np.abs(res["out"].astype(np.float64) - ref["out"])
    assert np.all(err <= 1e-5 * peak), "reverb bus error %.3g of block peak" % float((err / peak).max())
    pk = np.abs(ref["peaks"]).max()
    assert np.all(np.abs(res["peaks"].astype(np.float64) - ref["peaks"]) <= 1e-5 * pk)
    assert same_bits(res["sampler_offsets"], ref["sampler_offsets"])


@pytest.mark.parametrize("mode", ["direct", "tc", "fft", "fft512"])
@pytest.mark.parametrize("shape", [dict(B=101, out_channels=2), dict(B=256, out_channels=1), dict(B=77, out_channels=1),
                                   dict(B=1000, out_channels=2)])
def test_reverb_odd_blocks_and_mono_bus(wb, shape, mode, monkeypatch):
    """The reverb paths on odd block sizes (renders that are no whole number of partitions, unaligned track buffers) and
    on a mono bus (the FFT path then carries a real signal through its complex transforms), 3000 taps, vs the f64 spec."""
    _fir_mode(monkeypatch, mode)
    ref = sc.reverb(lambda C, B, r, bpm: o.Session("port", C, B, r, bpm), wb.effect_params, 3000, **shape)
    res = sc.reverb(gpu_engine(wb, True), wb.effect_params, 3000, **shape)
    peak = np.abs(ref["out"]).max(axis=(1, 2), keepdims=True)
    err = np.abs(res["out"].astype(np.float64) - ref["out"])
    assert np.all(err <= 1e-5 * peak), "reverb bus error %.3g of block peak" % float((err / peak).max())
    pk = np.abs(ref["peaks"]).max()
    assert np.all(np.abs(res["peaks"].astype(np.float64) - ref["peaks"]) <= 1e-5 * pk)


def _fir_mode(monkeypatch, mode):
    """WBX_FIR picks the reverb path when an impulse response is set; "fft512" = the FFT path with 512-tap partitions."""
    monkeypatch.setenv("WBX_FIR", "fft" if mode.startswith("fft") else mode)
    if mode == "fftauto":  # partition size chosen from the render length
        monkeypatch.delenv("WBX_FFT_P", raising=False)
    else:
        monkeypatch.setenv("WBX_FFT_P", "512" if mode == "fft512" else "2048")


@pytest.mark.parametrize("mode", ["tc", "fft", "fft512"])
def test_reverb_cfg5_tap_count(wb, mode, monkeypatch):
    """BASELINE cfg 5's 65536-tap impulse response on the tensor-core and FFT paths, against the f64 specification."""
    _fir_mode(monkeypatch, mode)
    ref = sc.reverb(lambda C, B, r, bpm: o.Session("port", C, B, r, bpm), wb.effect_params, 65536)
    res = sc.reverb(gpu_engine(wb, True), wb.effect_params, 65536)
    peak = np.abs(ref["out"]).max(axis=(1, 2), keepdims=True)
    err = np.abs(res["out"].astype(np.float64) - ref["out"])
    assert np.all(err <= 1e-5 * peak), "reverb bus error %.3g of block peak" % float((err / peak).max())


@pytest.mark.parametrize("mode", ["tc", "fft", "fft512"])
@pytest.mark.parametrize("chunks", [None, [64, 64, 32]])
def test_reverb_cfg5_full_accumulation_depth(wb, chunks, mode, monkeypatch):
    """BASELINE cfg 5 on the tensor-core and FFT paths at full depth: 64 stereo tracks (128 signals) x 160 callbacks x 512 frames
    (81920 frames of signal > 65536 taps, in one render and carried across three), against the f64 specification; bus
    and VU peaks within 1e-5 of the block / track peak. Every one of the 1027 tap chunks multiplies real history here
    (the 777...65536-tap scenario above only ever fills 33 of them)."""
    _fir_mode(monkeypatch, mode)
    res = sc.reverb_full_depth(gpu_engine(wb, True), wb.effect_params, chunks=chunks)
    want, want_peaks = sc.reverb_f64_expected(res, wb.panning_coefs, wb.db_to_linear)
    peak = np.abs(want).max(axis=(1, 2), keepdims=True)
    err = np.abs(res["out"].astype(np.float64) - want) / peak
    assert float(err.max()) <= 1e-5, "reverb bus error %.3g of block peak (worst callback %d)" % (float(err.max()), int(err.max(axis=(1, 2)).argmax()))
    # late callbacks (history longer than the response) are as good as early ones: no drift with accumulation depth
    late = float(err[136:].max())
    assert late <= 1e-5, "late-callback error %.3g" % late
    pk = np.abs(want_peaks).max()
    assert float(np.abs(res["peaks"].astype(np.float64) - want_peaks).max()) <= 1e-5 * pk


@pytest.mark.parametrize("mode", ["fft", "fft512", "fftauto"])
def test_reverb_fft_spectra_ring_warm_equals_cold(wb, mode, monkeypatch):
    """FFT path: renders that are a whole number of partitions long reuse the previous render's window spectra (the ring
    in wbx_fir_fft.cu); WBX_FFT_COLD rebuilds every window from the time-domain history instead — also across a render
    that is NOT a whole number of partitions (5 callbacks) and a chain reset in the middle of the session (both fall back
    to the rebuild). The two agree to rounding, not to the bit: a reused window may still hold frames older than the
    taps - 1 frames the time-domain history keeps; they only meet zero-padded taps, but their rounding noise differs."""
    def run(cold):
        _fir_mode(monkeypatch, mode)
        if cold:
            monkeypatch.setenv("WBX_FFT_COLD", "1")
        else:
            monkeypatch.delenv("WBX_FFT_COLD", raising=False)
        rng = np.random.RandomState(77)
        eng = wb.Engine(2, 512, 48000, 120.0, device=0, sum_mode=wb.SUM_EXACT)
        taps = 9000
        ir = (rng.standard_normal(taps) * np.exp(-np.arange(taps) / (taps / 5.0)) * 0.02).astype(np.float32)
        eng.set_impulse_response(ir)
        for t in range(6):
            eng.add_track(-5.0, 0.15 * t - 0.3, False)
            sid = eng.add_sample(sc._src(rng, 2, 80 * 512, 6), 48000)
            eng.add_clip(t, sid, 0.0, 1e6, 0.0, 1.0, 0.7)
            if t != 4:
                eng.set_effects(t, wb.effect_params(reverb=True))
        eng.play()
        outs = [eng.render(n) for n in (16, 16, 8, 5, 16, 16)]
        eng.set_effects(2, wb.effect_params(reverb=True))  # re-attach: track 2 restarts from silence
        outs += [eng.render(n) for n in (16, 16)]
        return outs
    warm, cold = run(False), run(True)
    for i, ((a, pa), (b, pb)) in enumerate(zip(warm, cold)):
        peak = float(np.abs(b).max())
        assert peak > 1e-4
        err = float(np.abs(a.astype(np.float64) - b).max()) / peak
        assert err <= 2e-6, "render %d: warm vs cold differ by %.3g of the peak" % (i, err)
        assert float(np.abs(pa.astype(np.float64) - pb).max()) <= 2e-6 * float(np.abs(pb).max())


def test_reverb_delta_is_identity(wb):
    """h = [1]: the convolution multiplies by exactly 1, so the render equals the chain-free one bit for bit."""
    def run(with_ir):
        rng = np.random.RandomState(12)
        eng = wb.Engine(2, 512, 48000, 120.0, device=0, sum_mode=wb.SUM_EXACT)
        if with_ir:
            eng.set_impulse_response(np.array([1.0, 0.0, 0.0], np.float32))
        for t in range(4):
            eng.add_track(-4.0, 0.1 * t, False)
            sid = eng.add_sample(sc._src(rng, 2, 5000, 4), 48000)
            eng.add_clip(t, sid, 0.0, 8.0, 0.0, 1.0, 0.8)
            if with_ir and t < 2:
                eng.set_effects(t, wb.effect_params(reverb=True))
        eng.play()
        return eng.render(5)
    (a, pa), (b, pb) = run(False), run(True)
    assert same_bits(a, b) and same_bits(pa, pb)


def test_effects_identity_properties(wb):
    """A compressor that never reaches its threshold multiplies by exactly 1: bit-identical to no chain."""
    def run(with_fx):
        rng = np.random.RandomState(11)
        eng = wb.Engine(2, 512, 48000, 120.0, device=0, sum_mode=wb.SUM_EXACT)
        for t in range(6):
            eng.add_track(-4.0, 0.1 * t, False)
            sid = eng.add_sample(sc._src(rng, 2, 5000, 6), 48000)
            eng.add_clip(t, sid, 0.0, 8.0, 0.0, 1.0, 0.8)
            if with_fx and t % 2 == 0:
                eng.set_effects(t, wb.effect_params(threshold_db=20.0, ratio_code=2))
        eng.play()
        return eng.render(6)
    (a, pa), (b, pb) = run(False), run(True)
    assert same_bits(a, b) and same_bits(pa, pb)


def test_waveform_mipmaps_vs_reference(wb):
    """WaveformVisual mip-maps (gfx/waveform_visual.cpp) of resident samples == the CPU checker
    (the reference's own summarize_for_mipmaps_impl when oracle/_ref travelled, else the C restatement)."""
    kind = "reference" if o.have_ref() else "port"
    rng = np.random.RandomState(5)
    dev = wb.DeviceEngine(0)
    dev.configure(2, 512, 48000)
    for fmt, frames, ch in [(9, 100000, 2), (9, 70, 1), (9, 65, 2), (9, 64, 1), (3, 33333, 2), (7, 5000, 1),
                            (9, 4097, 2), (9, 1 << 16, 1), (9, 300001, 2)]:
        data = sc._src(rng, ch, frames, 1, fmt)
        if fmt == 9:
            data[:, :7] = [1, -1, 0, 0.5, -0.5, 1, -1]
        s = o.Session(kind)
        ref_id = s.add_sample(data, 48000, fmt)
        sid = dev.sample_upload(data, 48000, fmt)
        for q in (0, 1):
            want = s.mipmaps(ref_id, q)
            got = dev.sample_mipmaps(sid, q, ch)
            assert len(got) == len(want), (fmt, frames, ch, q)
            for lv, (a, b) in enumerate(zip(got, want)):
                assert a.shape == b.shape and np.array_equal(a, b), "mip level %d differs (%s)" % (lv, (fmt, frames, ch, q))
        dev.sample_release(sid)


def test_scalars_and_interleave(wb, golden_dir):
    g = np.load(os.path.join(golden_dir, "scalars.npz"))
    planar = g["planar"] + np.float32(0)  # the bus starts at +0, so a -0.0 source sample mixes to +0.0
    n = planar.shape[1]
    eng = wb.Engine(2, n, 48000, 120.0, device=0)
    eng.add_track(0.0, 0.0, False)  # 0 dB, centre pan: gains exactly (1, 1)
    sid = eng.add_sample(planar, 48000)
    eng.add_clip(0, sid, 0.0, 1e6, 0.0, 1.0, 1.0)
    eng.play()
    out, _ = eng.render(1)
    assert same_bits(out, planar)
    for f in (wb.FMT_I16, wb.FMT_I24, wb.FMT_I24_X8, wb.FMT_I32, wb.FMT_F32):
        got = eng.dev.fetch_interleaved(f)
        want = g["conv_%d" % f]
        if f == wb.FMT_F32:
            want = np.ascontiguousarray(planar.T).reshape(-1).view(np.uint8)
        if f == wb.FMT_I24:  # the reference writes frames*3 bytes (no channel stride); its buffer is channels x larger
            want = want[: got.size]
        assert same_bits(got, want), "interleave format %d" % f


# ---- (b) larger seeded inputs against the CPU checker ------------------------------------------------

def test_fuzz_vs_cpu(wb):
    L = o.lib("port")
    L.wbo_ub_count.restype = ctypes.c_uint64
    ran = 0
    for seed in range(4, 44):
        before = L.wbo_ub_count()
        ref = sc.fuzz(lambda C, B, r, bpm: o.Session("port", C, B, r, bpm), seed)
        if L.wbo_ub_count() != before:
            continue  # reference UB (out-of-order events), no defined answer
        assert_exact(sc.fuzz(gpu_engine(wb, seed % 3 != 0), seed), ref, "fuzz%d" % seed)
        ran += 1
    assert ran > 25


def test_cfg2_256_tracks(wb):
    ref = sc.standard(cpu_engine(), 256, 2, 48000, 12)
    assert_exact(sc.standard(gpu_engine(wb), 256, 2, 48000, 12), ref, "cfg2 N=256")
    assert_tree(sc.standard(gpu_engine(wb, True, wb.SUM_TREE), 256, 2, 48000, 12), ref, "cfg2 N=256 tree")


def test_cfg3_128_tracks(wb):
    ref = sc.standard(cpu_engine(), 128, 2, 44100, 10)
    assert_exact(sc.standard(gpu_engine(wb), 128, 2, 44100, 10), ref, "cfg3 N=128")


def test_cfg1_mono_many_blocks(wb):
    """K = 700 callbacks: every warp handles several work items, stage parities wrap many times."""
    ref = sc.standard(cpu_engine(), 16, 1, 48000, 700)
    assert_exact(sc.standard(gpu_engine(wb), 16, 1, 48000, 700), ref, "cfg1 K=700")


@pytest.mark.parametrize("block", [1, 37, 1000, 2048, 4096])
def test_block_sizes(wb, block):
    """Engine::set_audio_channel_config buffer sizes off the beaten path: odd sizes (scalar bus stores), blocks larger
    than a 512-frame tile (several tiles per block, VU peaks merged with atomicMax), a single-frame block."""
    def run(mk):
        rng = np.random.RandomState(block)
        eng = mk(2, block, 48000, 120.0)
        n = 9
        for t in range(n):
            eng.add_track(-3.0 - t, -0.9 + 0.2 * t, False)
            fmt = sc.FMT_I16 if t == 4 else sc.FMT_F32
            sid = eng.add_sample(sc._src(rng, 1 if t == 6 else 2, 30000, n, fmt), 44100 if t % 3 == 2 else 48000, fmt)
            eng.add_clip(t, sid, 0.013 * t, 64.0, float(t % 5), 1.0, 0.8)
        eng.play()
        k = max(3, min(40, 9000 // block))
        return sc._collect(eng, [eng.process(k), eng.process(2)], n)
    assert_exact(run(gpu_engine(wb)), run(cpu_engine()), "block=%d" % block)
    assert_tree(run(gpu_engine(wb, True, wb.SUM_TREE)), run(cpu_engine()), "block=%d tree" % block)


def test_unaligned_offsets_and_speeds(wb):
    """Start offsets 1..7 frames (16-byte-unaligned windows), speeds straddling the staged-window limit."""
    def run(mk):
        rng = np.random.RandomState(2024)
        eng = mk(2, 512, 48000, 120.0)
        speeds = [1.0, 1.0, 1.0, 1.0, 0.91875, 0.5, 1.02, 1.0625, 1.5, 2.0, 3.7, 0.1]
        for t, sp in enumerate(speeds):
            eng.add_track(-3.0, 0.1 * (t - 5), False)
            sid = eng.add_sample(sc._src(rng, 2, 9000, len(speeds)), 48000)
            eng.add_clip(t, sid, 0.0, 64.0, float(t % 8), sp, 0.9)
        eng.play()
        return sc._collect(eng, [eng.process(5)], len(speeds))
    assert_exact(run(gpu_engine(wb)), run(cpu_engine()), "unaligned/speeds")


def test_lean_and_general_batches_alternate(wb):
    """The mix kernel resolves 16 cells at a time and runs a batch whose cells all resolved to the same whole-tile
    stereo-f32 kind through a lean loop, anything else through the general one; both share the staging counters and the
    stage parities. 112 tracks = seven batches of every flavour in one tile — unity (K_FAST), a batch with silent and
    late-starting tracks, linear resample (K_LIN), odd start offsets (K_UNI), int16 sources, a batch whose LAST cells are
    silent in front of a lean batch (the general loop then leaves the pipeline full), unity again — over callbacks in
    which the late clips start mid-block (partial tiles turn a lean batch into a general one for that callback only)."""
    def run(mk):
        rng = np.random.RandomState(77)
        eng = mk(2, 512, 48000, 120.0)
        n = 112
        for t in range(n):
            b = t // 16
            eng.add_track(-4.0 - (t % 5), -0.8 + 0.1 * (t % 17), False)
            fmt = sc.FMT_I16 if b == 4 else sc.FMT_F32
            rate = 44100 if b == 2 else 48000
            sid = eng.add_sample(sc._src(rng, 2, 40000, n, fmt), rate, fmt)
            start = 0.0
            if b == 1 and t % 3 == 0:
                continue  # a silent track inside batch 1
            if b == 1 and t % 3 == 1:
                start = 0.011 * (t % 7 + 1)  # starts mid-block in a later callback
            if b == 5 and t % 16 >= 13:
                continue  # the batch's last cells are silent
            if b == 6 and t % 16 == 2:
                start = 0.037  # one late clip: batch 6 is general for one callback, lean afterwards
            off = float(1 + t % 5) if b == 3 else 0.0
            eng.add_clip(t, sid, start, 64.0, off, 1.0, 0.08 + 0.001 * (t % 9))  # the bus stays inside the clamp
        eng.play()
        return sc._collect(eng, [eng.process(6), eng.process(1), eng.process(3)], n)
    ref = run(cpu_engine())
    assert_exact(run(gpu_engine(wb)), ref, "lean/general batches")
    assert_exact(run(gpu_engine(wb, False)), ref, "lean/general batches, per callback")
    assert_tree(run(gpu_engine(wb, True, wb.SUM_TREE)), ref, "lean/general batches, tree order")


def test_cfg2_full_track_count_vs_cpu(wb):
    """BASELINE cfg 2 at its full 1024 tracks, 4 callbacks: direct bit-exact comparison."""
    ref = sc.standard(cpu_engine(), 1024, 2, 48000, 4)
    assert_exact(sc.standard(gpu_engine(wb), 1024, 2, 48000, 4), ref, "cfg2 N=1024")


def test_cfg3_full_track_count_vs_cpu(wb):
    ref = sc.standard(cpu_engine(), 1024, 2, 44100, 3)
    assert_exact(sc.standard(gpu_engine(wb), 1024, 2, 44100, 3), ref, "cfg3 N=1024")


# ---- (c) properties at full size -------------------------------------------------------------------------

@pytest.fixture(scope="module")
def big(wb):
    """1024 stereo tracks x 96 callbacks resident on the device (BASELINE cfg 2 shape)."""
    N, K, B = 1024, 96, 512
    rng = np.random.default_rng(7)
    dev = wb.DeviceEngine(0)
    dev.configure(2, B, 48000)
    dev.set_track_count(N)
    frames = K * B + 64
    segs = np.zeros(N, wb.SEGMENT_DTYPE)
    for t in range(N):
        x = ((rng.random((2, frames), dtype=np.float32) * 2 - 1) * np.float32(0.5 / 32)).astype(np.float32)
        sid = dev.sample_upload(x, 48000)
        segs[t] = (t, 0, K, 0, B, sid, 0.0, 1.0, 0.5 + 0.001 * (t % 512), 0, 0.0, 0.0, 0.0, 0.0)
    gains = np.stack([np.float32(0.15) + np.float32(0.001) * (np.arange(N) % 97),
                      np.float32(0.3) - np.float32(0.0005) * (np.arange(N) % 89)], axis=1).astype(np.float32)
    return dict(dev=dev, segs=segs, gains=gains, N=N, K=K, B=B)


def test_full_size_properties(wb, big):
    dev, segs, gains, N, K, B = (big[k] for k in ("dev", "segs", "gains", "N", "K", "B"))
    dev.set_sum_mode(wb.SUM_EXACT)
    out1, pk1 = dev.render(segs, gains, K)
    out1b, pk1b = dev.render(segs, gains, K)
    assert same_bits(out1, out1b) and same_bits(pk1, pk1b), "run-to-run determinism"
    assert np.abs(out1).max() > 0.02 and np.all(np.isfinite(out1))
    # linearity: doubling every track gain doubles bus and peaks exactly (power-of-two scaling, no clamp hit)
    out2, pk2 = dev.render(segs, gains * 2, K)
    assert np.abs(out2).max() < 1.0
    assert same_bits(out2, out1 * np.float32(2)) and same_bits(pk2, pk1 * np.float32(2)), "linearity in gain"
    # silence in the gains -> silence out
    out0, pk0 = dev.render(segs, gains * 0, K)
    assert not out0.any() and not pk0.any()
    # the bus can never exceed the sum of the per-track block peaks
    bound = pk1.sum(axis=1)  # [K][2]
    blockmax = np.abs(out1.reshape(2, K, B)).max(axis=2).T
    assert np.all(blockmax <= bound * (1 + 1e-5))
    # tree order agrees with exact order within tolerance; peaks identical
    dev.set_sum_mode(wb.SUM_TREE)
    outt, pkt = dev.render(segs, gains, K)
    dev.set_sum_mode(wb.SUM_EXACT)
    assert same_bits(pkt, pk1)
    peak = np.abs(out1.reshape(2, K, B)).max(axis=(0, 2))
    err = np.abs(outt.astype(np.float64) - out1).reshape(2, K, B).max(axis=(0, 2))
    assert np.all(err <= TREE_TOL * peak)
    # f64 ground truth of one callback (numpy, from the same device-resident definition): both orders within 1e-5
    # sharding: two half-track shards mixed unclamped and added == what one NCCL reduce would produce
    half = N // 2
    a = segs[:half].copy()
    b = segs[half:].copy()
    dev.submit(a, gains, K)
    dev.mix(wb.MIX_NO_CLAMP)
    pa, ka = dev.fetch(True)
    dev.submit(b, gains, K)
    dev.mix(wb.MIX_NO_CLAMP)
    pb, kb = dev.fetch(True)
    summed = np.clip(pa + pb, -1.0, 1.0)
    err = np.abs(summed.astype(np.float64) - out1).reshape(2, K, B).max(axis=(0, 2))
    assert np.all(err <= TREE_TOL * peak), "track-sharded partial buses"
    assert same_bits(np.maximum(ka, kb), pk1), "sharded peaks"


def test_single_track_solo_matches_peak(wb, big):
    """With every other track at gain 0 the bus IS that track's term: max|bus| per block == its VU peak."""
    dev, segs, gains, N, K, B = (big[k] for k in ("dev", "segs", "gains", "N", "K", "B"))
    g = np.zeros_like(gains)
    g[777] = gains[777]
    out, pk = dev.render(segs, g, K)
    blockmax = np.abs(out.reshape(2, K, B)).max(axis=2).T
    assert same_bits(np.ascontiguousarray(blockmax), np.ascontiguousarray(pk[:, 777, :]))
    assert not pk[:, :777].any() and not pk[:, 778:].any()


def test_errors_are_reported(wb):
    dev = wb.DeviceEngine(0)
    dev.configure(2, 512, 48000)
    dev.set_track_count(2)
    segs = np.zeros(1, wb.SEGMENT_DTYPE)
    segs[0] = (5, 0, 1, 0, 512, 0, 0.0, 1.0, 1.0, 0, 0.0, 0.0, 0.0, 0.0)  # track out of range, unknown sample
    with pytest.raises(wb.WbxError):
        dev.submit(segs, np.ones((2, 2), np.float32), 1)
    with pytest.raises(wb.WbxError):
        dev.configure(3, 512, 48000)  # pan_coeffs[2]: at most 2 output channels
    with pytest.raises(wb.WbxError):
        dev.mix()  # nothing submitted


def test_cpp_dropin_demo(wb, tmp_path):
    """examples/dropin_demo.cpp: a pure C++ caller (no Python) drives wbx::Engine::process per callback with
    AudioBuffer-shaped buffers and gets the same samples as one offline bounce."""
    import subprocess
    from test_host_cpu import build_dropin_demo
    exe = build_dropin_demo(str(tmp_path))
    r = subprocess.run([exe, "96", "40"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 samples differ" in r.stdout


def test_device_api_side_doors(wb):
    """wbx_sample_update (streamed sources), wbx_fetch_levels (VU level = max over callbacks), page-locked output
    channels (direct D2H) and the three-stage submit / mix / fetch split all agree with the plain wbx_render path."""
    rng = np.random.RandomState(77)
    N, K, B = 24, 9, 512
    dev = wb.DeviceEngine(0)
    dev.configure(2, B, 48000)
    dev.set_track_count(N)
    dev.set_sum_mode(wb.SUM_EXACT)
    frames = K * B + 64
    data = [sc._src(rng, 2, frames, N) for _ in range(N)]
    segs = np.zeros(N, wb.SEGMENT_DTYPE)
    for t in range(N):
        sid = dev.sample_upload(np.zeros_like(data[t]), 48000)  # placeholder content ...
        segs[t] = (t, 0, K, 0, B, sid, 0.0, 1.0, 0.7, 0, 0.0, 0.0, 0.0, 0.0)
    for t in range(N):
        dev.sample_update_planar(t, [data[t][0], data[t][1]])  # ... replaced in place
    gains = np.full((N, 2), 0.5, np.float32)
    out, peaks = dev.render(segs, gains, K)
    ref = wb.DeviceEngine(0)
    ref.configure(2, B, 48000)
    ref.set_track_count(N)
    ref.set_sum_mode(wb.SUM_EXACT)
    for t in range(N):
        ref.sample_upload(data[t], 48000)
    out2, peaks2 = ref.render(segs, gains, K)
    assert same_bits(out, out2) and same_bits(peaks, peaks2), "sample_update != fresh upload"
    assert same_bits(dev.fetch_levels(), peaks.max(axis=0)), "fetch_levels != max over callbacks"
    pinned = wb.PinnedArray((2, K * B))
    assert dev.L.wbx_fetch(dev.h, wb._chan_ptrs(pinned.array), None) == 0
    assert same_bits(pinned.array, out), "page-locked direct fetch differs"
    assert dev.launch_count() > 0 and "exact" in dev.last_kernel()


# ---- page-locked output channels written by the mix kernel itself; the sharded bus exchange over peer memory ---------

@pytest.mark.parametrize("K,B,mode", [(33, 512, "exact"), (1, 512, "tree"), (5, 37, "exact"), (3, 1000, "tree"),
                                      (2, 2048, "exact")])
def test_render_into_page_locked_channels(wb, K, B, mode):
    """wbx_render_levels with page-locked out_channels: the kernel's own stores into host memory (no D2H copy) carry
    exactly the device bus, and the levels returned under the same synchronise are the max of the block peaks."""
    rng = np.random.RandomState(5)
    N = 40
    dev = wb.DeviceEngine(0)
    dev.configure(2, B, 48000)
    dev.set_track_count(N)
    dev.set_sum_mode(wb.SUM_EXACT if mode == "exact" else wb.SUM_TREE)
    frames = K * B + 64
    segs = np.zeros(N, wb.SEGMENT_DTYPE)
    for t in range(N):
        sid = dev.sample_upload(sc._src(rng, 2, frames, 4), 48000)  # hot enough to hit the clamp now and then
        segs[t] = (t, 0, K, 0, B, sid, 0.0, 1.0 if t % 3 else 0.77, 0.7, 0, 0.0, 0.0, 0.0, 0.0)
    gains = np.full((N, 2), 0.5, np.float32)
    out, peaks = dev.render(segs, gains, K)
    pinned = wb.PinnedArray((2, K * B))
    pinned.array[:] = 7.0
    lv = np.zeros((N, 2), np.float32)
    pk = np.zeros((K, N, 2), np.float32)
    rc = dev.L.wbx_render_levels(dev.h, segs.ctypes.data, N, gains.ctypes.data, K, wb._chan_ptrs(pinned.array),
                                 pk.ctypes.data, lv.ctypes.data)
    assert rc == 0
    assert same_bits(pinned.array, out), "kernel-written host channels differ from the device bus"
    assert same_bits(pk, peaks) and same_bits(lv, peaks.max(axis=0))
    out_dev, _ = dev.fetch(False)  # the device copy of the bus is still there (fetch_interleaved etc. rely on it)
    assert same_bits(out_dev, out)
    # pageable channels through the same entry point
    plain = np.zeros((2, K * B), np.float32)
    assert dev.L.wbx_render_levels(dev.h, segs.ctypes.data, N, gains.ctypes.data, K, wb._chan_ptrs(plain), None, None) == 0
    assert same_bits(plain, out)
    # ... and the same pageable allocation page-locked in place (wbx_host_register): the kernel writes it directly
    plain[:] = 3.0
    assert dev.L.wbx_host_register(plain.ctypes.data, plain.nbytes) == 0
    try:
        assert dev.L.wbx_render_levels(dev.h, segs.ctypes.data, N, gains.ctypes.data, K, wb._chan_ptrs(plain), None, None) == 0
        assert same_bits(plain, out)
    finally:
        dev.synchronize()
        assert dev.L.wbx_host_unregister(plain.ctypes.data) == 0


def _sharded_setup(wb, world, N, K, B, C=2, seed=11, empty_rank=None):
    """`world` engines on device 0, tracks dealt contiguously (whitebox_b200.shard.track_range); -> engines + data."""
    from whitebox_b200 import shard
    rng = np.random.RandomState(seed)
    frames = K * B + 64
    data = [sc._src(rng, 2, frames, 1) for _ in range(N)]  # hot: the summed bus reaches the clamp
    gains = (0.2 + 0.6 * rng.rand(N, 2)).astype(np.float32)
    clip_gain = (0.5 + 0.5 * rng.rand(N)).astype(np.float32)
    devs, parts = [], []
    for r in range(world):
        lo, hi = shard.track_range(N, r, world)
        if empty_rank == r:
            hi = lo
        dev = wb.DeviceEngine(0)
        dev.configure(C, B, 48000)
        dev.set_track_count(hi - lo)
        dev.set_sum_mode(wb.SUM_EXACT)
        segs = np.zeros(hi - lo, wb.SEGMENT_DTYPE)
        for i, t in enumerate(range(lo, hi)):
            sid = dev.sample_upload(data[t], 48000)
            segs[i] = (i, 0, K, 0, B, sid, 3.0 if t % 2 else 0.0, 1.0 if t % 4 else 0.9, clip_gain[t], 0, 0.0, 0.0, 0.0, 0.0)
        devs.append(dev)
        parts.append((segs, gains[lo:hi].copy()))
    return devs, parts


@pytest.mark.parametrize("world,K,B,C", [(2, 8, 512, 2), (3, 7, 512, 2), (4, 2, 37, 2), (2, 5, 256, 1), (8, 3, 128, 2)])
def test_sharded_peer_memory_exchange(wb, world, K, B, C):
    """The sharded render (wbx_mix_sharded: tiles stored into the owners' exchange buffers, flag barrier, owner reduce
    in rank order + clamp into rank 0's master bus) against the same partial buses mixed unclamped per shard and added
    in rank order on the host — bit for bit — for even / ragged callback splits and mono / stereo buses."""
    from whitebox_b200 import shard
    N = 24
    devs, parts = _sharded_setup(wb, world, N, K, B, C)
    partial = []
    for dev, (segs, gains) in zip(devs, parts):
        dev.submit(segs, gains, K)
        dev.mix(wb.MIX_NO_CLAMP)
        partial.append(dev.fetch(True))
    want = partial[0][0].copy()
    for r in range(1, world):
        want = (want + partial[r][0]).astype(np.float32)
    assert np.abs(want).max() > 1.0, "fixture should reach the clamp"
    want = np.where(want > 1, np.float32(1), np.where(want < -1, np.float32(-1), want)).astype(np.float32)
    for r, dev in enumerate(devs):
        dev.shard_init(r, world, K + 3)
    for dev in devs:
        dev.shard_connect_local(devs)
    host = wb.PinnedArray((C, K * B))
    for rep in range(3):  # epochs advance; buffers are reused
        if rep == 2:  # owners also mirror their slices into one host buffer every rank maps: rank 0 copies nothing
            host.array[:] = 9.0
            for dev in devs:
                dev.shard_set_host_output(host.array)
        for dev, (segs, gains) in zip(devs, parts):
            dev.submit(segs, gains, K)
        shard.mix_sharded_lockstep(devs)
        out, pk0 = devs[0].fetch(True)
        assert same_bits(out, want), "sharded master bus (repeat %d)" % rep
        if rep == 2:
            assert devs[0].L.wbx_fetch(devs[0].h, wb._chan_ptrs(host.array), None) == 0
            for dev in devs[1:]:
                dev.synchronize()
            assert same_bits(host.array, want), "host output written by the owner ranks"
        assert same_bits(pk0, partial[0][1])
        for r in range(1, world):
            _, pk = devs[r].fetch(True, want_bus=False)
            assert same_bits(pk, partial[r][1]), "rank %d peaks" % r
            with pytest.raises(wb.WbxError):
                devs[r].fetch(False)  # only rank 0 holds the master bus
    assert "peer-reduce" in devs[0].last_kernel()
    assert devs[1].shard_info() == (1, world)
    for dev in devs:
        dev.shard_close()
    assert devs[1].shard_info() == (0, 1)


def test_sharded_rank_without_tracks_and_full_mix(wb):
    """A rank whose shard is empty contributes silence; the sharded result stays within tolerance of one engine holding
    every track (re-association across shards only)."""
    world, N, K, B = 3, 18, 6, 512
    devs, parts = _sharded_setup(wb, world, N, K, B, seed=3, empty_rank=1)
    for r, dev in enumerate(devs):
        dev.shard_init(r, world, K)
    for dev in devs:
        dev.shard_connect_local(devs)
    from whitebox_b200 import shard
    for dev, (segs, gains) in zip(devs, parts):
        dev.submit(segs, gains, K)
    shard.mix_sharded_lockstep(devs)
    out, _ = devs[0].fetch(False)
    devs[1].synchronize()
    devs[2].synchronize()
    # one engine with the tracks of ranks 0 and 2
    one, parts1 = _sharded_setup(wb, 1, N, K, B, seed=3)
    keep = [t for r in (0, 2) for t in range(*shard.track_range(N, r, world))]
    segs, gains = parts1[0]
    full, _ = one[0].render(segs[keep], gains, K)  # gains stay indexed by track id
    peak = max(float(np.abs(full).max()), 1e-30)
    assert np.abs(out.astype(np.float64) - full).max() <= TREE_TOL * peak


def test_sharded_errors(wb):
    dev = wb.DeviceEngine(0)
    dev.configure(2, 512, 48000)
    dev.set_track_count(0)
    dev.submit(np.zeros(0, wb.SEGMENT_DTYPE), np.zeros((0, 2), np.float32), 2)
    with pytest.raises(wb.WbxError):
        dev.mix_sharded()  # not initialised
    with pytest.raises(wb.WbxError):
        dev.shard_init(3, 2, 8)  # rank >= world
    dev.shard_init(0, 1, 1)
    dev.shard_connect_local([dev])
    with pytest.raises(wb.WbxError):
        dev.mix_sharded()  # 2 callbacks > max_blocks 1
    dev.shard_init(0, 1, 4)  # re-init replaces the block; world 1 is a plain clamp
    dev.shard_connect_local([dev])
    dev.mix_sharded()
    out, _ = dev.fetch(False)
    assert not out.any()
    # host output: pageable memory is refused, registered memory is accepted and written by the owner's reduce
    plain = np.full((2, 2 * 512), 5.0, np.float32)
    with pytest.raises(wb.WbxError):
        dev.shard_set_host_output(plain)
    assert dev.L.wbx_host_register(plain.ctypes.data, plain.nbytes) == 0
    dev.shard_set_host_output(plain)
    dev.mix_sharded()
    assert dev.L.wbx_fetch(dev.h, wb._chan_ptrs(plain), None) == 0
    assert not plain.any()
    dev.shard_set_host_output(None)
    assert dev.L.wbx_host_unregister(plain.ctypes.data) == 0
    assert dev.L.wbx_host_register(None, 16) != 0
    with pytest.raises(wb.WbxError):
        dev.mix_sharded(2)  # phase out of order


def test_sharded_two_devices_one_process(wb):
    """Two engines on two GPUs of one box, driven by one thread: real peer stores over NVLink (skipped on 1-GPU boxes;
    the one-process-per-GPU form over CUDA IPC is what bench.py --gpus N runs)."""
    import torch
    from whitebox_b200 import shard
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.RandomState(21)
    N, K, B, world = 16, 10, 512, 2
    frames = K * B + 64
    devs, parts, partial = [], [], []
    for r in range(world):
        lo, hi = shard.track_range(N, r, world)
        dev = wb.DeviceEngine(r)
        dev.configure(2, B, 48000)
        dev.set_track_count(hi - lo)
        dev.set_sum_mode(wb.SUM_EXACT)
        segs = np.zeros(hi - lo, wb.SEGMENT_DTYPE)
        for i in range(hi - lo):
            sid = dev.sample_upload(sc._src(rng, 2, frames, 1), 48000)
            segs[i] = (i, 0, K, 0, B, sid, 0.0, 1.0, 0.8, 0, 0.0, 0.0, 0.0, 0.0)
        gains = np.full((hi - lo, 2), 0.4, np.float32)
        dev.submit(segs, gains, K)
        dev.mix(wb.MIX_NO_CLAMP)
        partial.append(dev.fetch(False)[0])
        devs.append(dev)
        parts.append((segs, gains))
    want = np.clip((partial[0] + partial[1]).astype(np.float32), np.float32(-1), np.float32(1))
    for r, dev in enumerate(devs):
        dev.shard_init(r, world, K)
    for dev in devs:
        dev.shard_connect_local(devs)
    for _ in range(2):
        for dev, (segs, gains) in zip(devs, parts):
            dev.submit(segs, gains, K)
        shard.mix_sharded_lockstep(devs)
        out, _ = devs[0].fetch(False)
        devs[1].synchronize()
        assert same_bits(out, want)


# ---- every golden scenario through a sharded session (tracks dealt over 3 engines, bus exchange over peer memory) -------

@pytest.mark.parametrize("name", sorted(sc.ALL))
def test_golden_sharded_session(wb, golden_dir, name):
    """whitebox_b200.shard.ShardedEngine (= include/wbx_sharded.hpp): the same editing / transport calls, tracks i % 3 on
    three engines, master bus from the exchange. Against the reference's golden vectors: VU peaks bit-exact, bus within
    the re-association tolerance (the partial buses are added in rank order, not track order)."""
    from whitebox_b200 import shard
    gold = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    made = []

    def make(C, B, r, bpm):
        eng = shard.ShardedEngine([0, 0, 0], C, B, r, bpm, max_blocks=4096)
        made.append(eng)
        return eng

    res = sc.ALL[name](make)
    assert_tree(res, gold, name + " (3 shards)")
    for eng in made:
        eng.close()


def test_cpp_sharded_demo(wb, tmp_path):
    """examples/sharded_demo.cpp: wbx::ShardedEngine (header-only C++) with 2 shards == one wbx::Engine within tolerance,
    through Engine::process-shaped calls."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "sharded_demo")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "sharded_demo.cpp"),
                    "-L" + os.path.join(root, "whitebox_b200"), "-lwbx", "-Wl,-rpath," + os.path.join(root, "whitebox_b200"),
                    "-o", exe], check=True)
    r = subprocess.run([exe, "48", "12"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "sharded == single within tolerance" in r.stdout


# ---- threading contract of the boundary (include/wbx_engine.hpp) ------------------------------------------------

def test_ui_thread_edits_while_audio_thread_renders(wb):
    """A UI thread hammers set_volume / set_pan / set_mute (lock-free SPSC ring) and add_clip (editor lock) through the C
    exports while this thread renders callbacks on the GPU (ctypes drops the GIL inside both). Track 0 plays a DC source
    at pan 0, so every callback's bus reveals the volume it used: it must be one the UI wrote, never an older one than
    the callback before (some serialised schedule), and the run must not crash or deadlock."""
    import threading
    B = 128
    eng = wb.Engine(2, B, 48000, 120.0, device=0, sum_mode=wb.SUM_EXACT)
    n_tracks = 6
    dc = np.full((2, 1 << 20), 0.25, np.float32)
    noise = (np.random.RandomState(5).uniform(-1, 1, (2, 1 << 16)) * 1e-3).astype(np.float32)
    for t in range(n_tracks):
        eng.add_track(0.0, 0.0, t != 0)  # only track 0 is audible: the others are muted (and stay so: see ui())
        sid = eng.add_sample(dc if t == 0 else noise, 48000)
        eng.add_clip(t, sid, 0.0, 1e6, 0.0, 1.0, 1.0)
    n_writes = 4000
    dbs = (-40.0 + 40.0 * np.arange(n_writes) / n_writes).astype(np.float32)
    vols = np.array([wb.db_to_linear(float(d)) for d in dbs], np.float32)
    written = [0]
    stop = threading.Event()

    def ui():
        i = 0
        while not stop.is_set() and i + 1 < n_writes:
            i += 1
            eng.set_volume(0, float(dbs[i]))
            written[0] = i
            eng.set_pan(1 + i % (n_tracks - 1), -1.0 + (i % 50) / 25.0)
            eng.set_mute(1 + i % (n_tracks - 1), True)
            if i % 40 == 0:
                eng.add_clip(1 + (i // 40) % (n_tracks - 1), 1, 2.0 + 0.01 * i, 2.2 + 0.01 * i, 0.0, 1.0, 0.5)
                eng.level(0, 0, True)
                eng.cpu_usage()

    eng.play()
    eng.render(1)  # consumes the constructor's messages
    th = threading.Thread(target=ui)
    th.start()
    last = 0
    try:
        for k in range(1500):
            out, _ = eng.render(1, want_peaks=False)
            hi = written[0]
            v = np.float32(out[0, 0] * np.float32(4.0))  # 0.25 * volume * 1.0, exact
            assert np.all(out[0] == out[0, 0]) and np.all(out[1] == out[0, 0]), "callback %d: not one volume" % k
            cand = np.nonzero(vols[last:min(hi + 2, n_writes)] == v)[0]
            assert v == np.float32(1.0) and last == 0 or len(cand), "callback %d: volume %r was never written (%d..%d)" % (k, v, last, hi)
            if len(cand):
                last = last + int(cand[0])
    finally:
        stop.set()
        th.join()
    assert last > 0, "the audio thread never saw a UI write"
    assert 0.0 <= eng.cpu_usage() <= 1.0
    eng.close()


# ---- INTEGRATION.md section A, compiled: the reference's own engine with its sample loops replaced by wbx ----------------

@pytest.mark.parametrize("name", sorted(sc.ALL) + ["fuzz0", "fuzz1", "fuzz2", "fuzz3"])
def test_reference_with_wbx_binding_equals_reference(wb, golden_dir, name):
    """oracle/_ref/libwbref_gpu.so = the reference's engine.cpp / track.cpp with exactly the sample loops of the hot path
    (Sampler::stream calls, apply_gain + VU loop, clear / mix / clamp) replaced by wbx_segment records and one
    wbx_render_levels per callback (oracle/patch_ref_gpu.py, oracle/ref_gpu_hooks.cpp), every other line — transport,
    editor lock, process_event, parameter queue, plugin slot — the reference's own. Driven through the reference's public
    API it must reproduce the CPU reference's golden vectors bit for bit: the drop-in claim, executed."""
    if not o.have_ref_gpu():
        pytest.skip("oracle/_ref/libwbref_gpu.so was not built (needs /root/reference at build time)")
    gold = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    mk = lambda C, B, r, bpm: o.Session("reference_gpu", C, B, r, bpm)  # noqa: E731
    res = sc.fuzz(mk, int(name[4:])) if name.startswith("fuzz") else sc.ALL[name](mk)
    assert o.lib("reference_gpu").wbo_kind() == b"reference+wbx"
    assert_exact(res, gold, "reference+wbx " + name)


# ---- offline bounce / export driver ----------------------------------------------------------------------

def _bounce_session(make, n_tracks=24, B=256, rate=48000):
    rng = np.random.RandomState(77)
    eng = make(2, B, rate, 120.0)
    spb = rate * 0.5
    for t in range(n_tracks):
        eng.add_track(-5.0 - (t % 6), -0.8 + 0.07 * t, False)
        fmt = o.FMT_I16 if t % 7 == 3 else o.FMT_F32
        sid = eng.add_sample(sc._src(rng, 2, 40000, 8, fmt), 44100 if t % 5 == 1 else 48000, fmt)
        eng.add_clip(t, sid, (13 * t) / spb, (13 * t + 30000) / spb, float(t), 1.0, 0.7)
        eng.add_clip(t, sid, (13 * t + 31000 + 5 * t) / spb, 1e6, 100.0, 1.0 if t % 3 else 0.75, 0.5)
    return eng


@pytest.mark.parametrize("fmt", ["I16", "I24_X8", "I32", "F32"])
def test_bounce_equals_reference_process_loop(wb, fmt, tmp_path):
    """wbx::Engine::bounce(start_beat, end_beat, format, sink): chunked render with the transport and sampler state carried
    from chunk to chunk, conversion on the device, copy-out overlapped with the next chunk's mix. Byte-identical to the
    reference's Engine::process loop from the same playhead followed by convert_f32_to_interleaved_* (core/
    audio_format_conv.cpp:5-106), cut at end_beat; the WAV file holds the same samples."""
    code = {"I16": o.FMT_I16, "I24_X8": o.FMT_I24_X8, "I32": o.FMT_I32, "F32": o.FMT_F32}[fmt]
    B, rate = 256, 48000
    start, end = 0.37, 3.21
    total_frames = int(np.ceil((end - start) * 0.5 * rate))
    n_blocks = (total_frames + B - 1) // B
    kind = "reference" if o.have_ref() else "port"
    ref = _bounce_session(lambda C, B_, r, bpm: o.Session(kind, C, B_, r, bpm))
    ref.set_playhead(start)
    ref.play()
    out, _ = ref.process(n_blocks)
    planar = np.ascontiguousarray(out.transpose(1, 0, 2).reshape(2, n_blocks * B))[:, :total_frames]
    want = o.interleave(kind, planar, code)
    eng = _bounce_session(gpu_engine(wb))
    got = eng.bounce(start, end, code, chunk_blocks=7)  # 15 chunks, the last one partial
    assert got.size == want.size and np.array_equal(got, want), "bounce %s differs from the reference's process loop" % fmt
    again = eng.bounce(start, end, code, chunk_blocks=4096)  # one chunk; the transport was left stopped at `start`
    assert np.array_equal(again, want)
    path = tmp_path / ("bounce_%s.wav" % fmt)
    frames = eng.bounce(start, end, code, chunk_blocks=16, path=path)
    assert frames == total_frames
    raw = np.fromfile(path, np.uint8)
    assert raw[:4].tobytes() == b"RIFF" and raw[8:16].tobytes() == b"WAVEfmt " and raw[36:40].tobytes() == b"data"
    data_len = int(np.frombuffer(raw[40:44].tobytes(), "<u4")[0])
    body = raw[44:44 + data_len]
    if fmt == "I24_X8":  # the file holds 24-bit PCM: the low three bytes of every I24_X8 word
        assert data_len == total_frames * 2 * 3
        assert np.array_equal(body.reshape(-1, 3), want.reshape(-1, 4)[:, :3])
    else:
        assert data_len == want.size and np.array_equal(body, want)
    eng.close()
