"""Seeded mixing scenarios shared by every implementation under test. TEST INFRASTRUCTURE ONLY.

A scenario is a function `run(make_engine) -> dict[str, np.ndarray]`. `make_engine(out_channels, block, rate,
bpm)` returns an object with the reference's editing API (tests/oracle_api.Session for the CPU checkers,
whitebox_b200.Engine for the CUDA path): add_track / add_sample / add_clip / set_volume / set_pan / set_mute /
set_playhead / play / stop / process(n_blocks) -> (out[K][C][B], peaks[K][N][2]) / sampler_offset /
sample_position / playhead.

Fixture values follow SURVEY.md §8(d): MT19937 seeds, sources uniform(-1,1)*0.5/sqrt(N), clip gain
0.5+0.001*(t%512), volume dB -6-(t%7), pan -1+0.2*(t%11), one clip per track from beat 0, bpm 120.
"""
import numpy as np

FMT_I16, FMT_I24, FMT_I32, FMT_F32 = 3, 5, 7, 9


def _src(rng, channels, frames, n_tracks, fmt=FMT_F32):
    x = (rng.uniform(-1.0, 1.0, size=(channels, frames)) * (0.5 / np.sqrt(n_tracks))).astype(np.float32)
    if fmt == FMT_F32:
        return x
    if fmt == FMT_I16:
        return np.clip(np.round(x.astype(np.float64) * 40000.0 * np.sqrt(n_tracks)), -32768, 32767).astype(np.int16)
    if fmt == FMT_I24:  # 24-bit widened to int32 (dsp/sample.cpp:20); some values beyond full scale -> clamp path
        return np.clip(np.round(x.astype(np.float64) * 1.2e7 * np.sqrt(n_tracks)), -(2**31), 2**31 - 1).astype(np.int32)
    return np.clip(np.round(x.astype(np.float64) * 3.0e9 * np.sqrt(n_tracks)), -(2**31), 2**31 - 1).astype(np.int32)


def _track_params(t):
    return dict(volume_db=-6.0 - (t % 7), pan=-1.0 + 0.2 * (t % 11), gain=np.float32(0.5 + 0.001 * (t % 512)))


def _collect(eng, outs, n_tracks):
    out = np.concatenate([o for o, _ in outs], axis=0)
    peaks = np.concatenate([p for _, p in outs], axis=0)
    offs = np.array([eng.sampler_offset(t) for t in range(n_tracks)], np.float64)
    pos = np.array([eng.sample_position(), eng.playhead()], np.float64)
    return dict(out=out, peaks=peaks, sampler_offsets=offs, transport=pos)


def standard(make_engine, n_tracks, src_channels, src_rate, n_blocks, block=512, rate=48000, seed=1234,
             fmt=FMT_F32, out_channels=2, chunks=None):
    """The SURVEY §8(d) fixture: one clip per track from beat 0 covering the whole render."""
    rng = np.random.RandomState(seed)
    eng = make_engine(out_channels, block, rate, 120.0)
    frames = int((n_blocks + 4) * block * src_rate / rate) + 64
    beats = (n_blocks + 2) * block / rate * 2.0  # bpm 120 -> 2 beats/s
    for t in range(n_tracks):
        p = _track_params(t)
        eng.add_track(p["volume_db"], p["pan"], False)
        sid = eng.add_sample(_src(rng, src_channels, frames, n_tracks, fmt), src_rate, fmt)
        eng.add_clip(t, sid, 0.0, beats, 0.0, 1.0, float(p["gain"]))
    eng.play()
    outs = []
    for n in (chunks or [n_blocks]):
        outs.append(eng.process(n))
    return _collect(eng, outs, n_tracks)


def kat(make_engine):
    """SURVEY §8(c) known answer: 1 track, ramp source @44.1k, block 16."""
    eng = make_engine(2, 16, 48000, 120.0)
    t = eng.add_track(-6.0, 0.25, False)
    i = np.arange(4096, dtype=np.float32)
    src = np.stack([(i + 1) * np.float32(1e-3), -(i + 1) * np.float32(1e-3)]).astype(np.float32)
    sid = eng.add_sample(src, 44100, FMT_F32)
    eng.add_clip(t, sid, 0.0, 1.0, 0.0, 1.0, 0.5)
    eng.play()
    return _collect(eng, [eng.process(1), eng.process(1)], 1)


def cfg1(make_engine, n_blocks=8):
    """BASELINE cfg 1: 16 mono tracks, 48 kHz f32, gain+pan only, 512-sample block."""
    return standard(make_engine, 16, 1, 48000, n_blocks)


def cfg2_small(make_engine, n_tracks=64, n_blocks=6):
    """BASELINE cfg 2 shape (stereo 48k unity, gain/pan + bus sum), reduced N."""
    return standard(make_engine, n_tracks, 2, 48000, n_blocks, chunks=[1, 2, n_blocks - 3])


def cfg3_small(make_engine, n_tracks=32, n_blocks=6):
    """BASELINE cfg 3 shape (stereo 44.1k -> 48k linear resample + mix), reduced N."""
    return standard(make_engine, n_tracks, 2, 44100, n_blocks, chunks=[2, n_blocks - 2])


def int_formats(make_engine, n_blocks=4):
    """I16 / I24-in-I32 / I32 sources, unity and resampled (dsp/sampler.cpp:109-144,161-193)."""
    rng = np.random.RandomState(77)
    eng = make_engine(2, 256, 48000, 120.0)
    n = 0
    for fmt in (FMT_I16, FMT_I24, FMT_I32, FMT_F32):
        for rate in (48000, 44100, 96000):
            p = _track_params(n)
            eng.add_track(p["volume_db"], p["pan"], False)
            sid = eng.add_sample(_src(rng, 2, 4096, 12, fmt), rate, fmt)
            eng.add_clip(n, sid, 0.0, 64.0, 3.0, 1.0, float(p["gain"]))
            n += 1
    eng.play()
    return _collect(eng, [eng.process(n_blocks)], n)


def event_split(make_engine):
    """Clips that start / stop mid-block, abut, exhaust their sample, play at speeds != 1, start with an
    offset; playback starting mid-clip (track.cpp:347-446, 664-724; sampler.cpp:99-104)."""
    rng = np.random.RandomState(4321)
    B, rate = 128, 48000
    eng = make_engine(2, B, rate, 120.0)
    spb = rate * 0.5  # samples per beat at bpm 120
    n = 0

    def tr(vol=-3.0, pan=0.0):
        nonlocal n
        eng.add_track(vol, pan, False)
        n += 1
        return n - 1

    def smp(frames, ch=2, r=48000):
        return eng.add_sample(_src(rng, ch, frames, 8), r, FMT_F32)

    # t0: clip starts mid-block (300 frames in) and ends mid-block
    t = tr(-2.0, -0.3)
    eng.add_clip(t, smp(5000), 300.0 / spb, 900.0 / spb, 0.0, 1.0, 0.8)
    # t1: two abutting clips on one track, boundary mid-block, second with a start offset
    t = tr(-4.0, 0.4)
    eng.add_clip(t, smp(5000), 0.0, 333.0 / spb, 0.0, 1.0, 0.7)
    eng.add_clip(t, smp(5000), 333.0 / spb, 1000.0 / spb, 17.0, 1.0, 0.9)
    # t2: sample shorter than its clip (exhausts mid-block)
    t = tr(0.0, 0.0)
    eng.add_clip(t, smp(421), 0.0, 4.0, 0.0, 1.0, 1.0)
    # t3: speed 0.5 / t4: speed 2.0 (sample exhausts) / t5: speed 1.25 on a 44.1k source
    t = tr(-1.0, 0.1)
    eng.add_clip(t, smp(3000), 10.0 / spb, 4.0, 5.0, 0.5, 0.6)
    t = tr(-1.5, -0.8)
    eng.add_clip(t, smp(900), 0.0, 4.0, 0.0, 2.0, 0.5)
    t = tr(-7.0, 0.9)
    eng.add_clip(t, smp(4000, 2, 44100), 64.0 / spb, 700.0 / spb, 11.0, 1.25, 1.1)
    # t6: mono source at unity speed -> both channels (sampler.cpp:147 `i % channels`)
    t = tr(-5.0, 0.5)
    eng.add_clip(t, smp(2000, 1), 50.0 / spb, 1200.0 / spb, 0.0, 1.0, 0.75)
    # t7: gap between two clips, second starts exactly on a block boundary
    t = tr(-3.0, -1.0)
    eng.add_clip(t, smp(1000), 0.0, 200.0 / spb, 0.0, 1.0, 0.5)
    eng.add_clip(t, smp(1000), 512.0 / spb, 1100.0 / spb, 0.0, 1.0, 0.5)
    # t8: muted / t9: no clips / t10: -inf dB
    t = tr(-3.0, 0.0)
    eng.set_mute(t, True)
    eng.add_clip(t, smp(2000), 0.0, 4.0, 0.0, 1.0, 1.0)
    tr(-3.0, 0.0)
    t = tr(-80.0, 0.0)
    eng.add_clip(t, smp(2000), 0.0, 4.0, 0.0, 1.0, 1.0)
    eng.play()
    outs = [eng.process(4), eng.process(1), eng.process(7)]
    # stop, move the playhead into the middle of clips, play again (track.cpp:376-395 mid-clip start)
    eng.stop()
    eng.set_playhead(450.0 / spb)
    eng.play()
    outs.append(eng.process(6))
    return _collect(eng, outs, n)


def overlaps(make_engine):
    """Clips added on top of existing ones: Engine::add_to_cliplist + reserve_track_region (engine.cpp:409-461, 478-569)
    trim the old clip's tail, trim its head (shifting its content start, clip_edit.h:128-150), split it in two, delete
    it, or do all of that across several clips; plus add-to-front / add-to-back / no-overlap insertion in the middle, a
    resampled clip whose head is trimmed, and an overlapping clip added while the session is playing."""
    rng = np.random.RandomState(9876)
    B, rate = 128, 48000
    eng = make_engine(2, B, rate, 120.0)
    spb = rate * 0.5  # samples per beat at bpm 120
    n = 0

    def tr(vol=-3.0, pan=0.0):
        nonlocal n
        eng.add_track(vol, pan, False)
        n += 1
        return n - 1

    def smp(frames, ch=2, r=48000):
        return eng.add_sample(_src(rng, ch, frames, 8), r, FMT_F32)

    # t0: new clip cuts the TAIL of an old one (old [0, 900) -> [0, 500), new [500, 1100))
    t = tr(-2.0, -0.3)
    eng.add_clip(t, smp(6000), 0.0, 900.0 / spb, 0.0, 1.0, 0.8)
    eng.add_clip(t, smp(6000), 500.0 / spb, 1100.0 / spb, 3.0, 1.0, 0.6)
    # t1: new clip cuts the HEAD of an old one (old [300, 1200) -> [700, 1200) with its content shifted by 400 frames)
    t = tr(-4.0, 0.4)
    eng.add_clip(t, smp(6000), 300.0 / spb, 1200.0 / spb, 21.0, 1.0, 0.7)
    eng.add_clip(t, smp(6000), 100.0 / spb, 700.0 / spb, 0.0, 1.0, 0.9)
    # t2: new clip in the MIDDLE of an old one: split into [0, 400) and [650, 1300) (right part shifted)
    t = tr(0.0, 0.0)
    eng.add_clip(t, smp(8000), 0.0, 1300.0 / spb, 5.0, 1.0, 1.0)
    eng.add_clip(t, smp(3000), 400.0 / spb, 650.0 / spb, 0.0, 1.0, 0.5)
    # t3: new clip COVERS an old one entirely (deleted), old neighbours untouched
    t = tr(-1.0, 0.1)
    eng.add_clip(t, smp(3000), 0.0, 150.0 / spb, 0.0, 1.0, 0.6)
    eng.add_clip(t, smp(3000), 300.0 / spb, 500.0 / spb, 0.0, 1.0, 0.6)
    eng.add_clip(t, smp(3000), 900.0 / spb, 1250.0 / spb, 0.0, 1.0, 0.6)
    eng.add_clip(t, smp(3000), 250.0 / spb, 600.0 / spb, 7.0, 1.0, 0.4)
    # t4: new clip spans FOUR old ones: first tail-trimmed, two deleted, last head-trimmed
    t = tr(-1.5, -0.8)
    for i in range(4):
        eng.add_clip(t, smp(3000), (i * 300.0) / spb, (i * 300.0 + 260.0) / spb, 2.0 * i, 1.0, 0.5 + 0.1 * i)
    eng.add_clip(t, smp(6000), 130.0 / spb, 1010.0 / spb, 0.0, 1.0, 0.3)
    # t5: add to the front, to the back, and into a gap (no overlap: plain insertion + ordering)
    t = tr(-7.0, 0.9)
    eng.add_clip(t, smp(2000), 600.0 / spb, 800.0 / spb, 0.0, 1.0, 1.1)
    eng.add_clip(t, smp(2000), 0.0, 200.0 / spb, 0.0, 1.0, 0.9)
    eng.add_clip(t, smp(2000), 1000.0 / spb, 1300.0 / spb, 0.0, 1.0, 0.8)
    eng.add_clip(t, smp(2000), 300.0 / spb, 500.0 / spb, 0.0, 1.0, 0.7)
    # t6: resampled 44.1 kHz clip at speed 1.25 whose head is cut: the shift is scaled by the clip speed and uses the
    # asset's own sample rate
    t = tr(-5.0, 0.5)
    eng.add_clip(t, smp(9000, 2, 44100), 100.0 / spb, 1200.0 / spb, 11.0, 1.25, 0.75)
    eng.add_clip(t, smp(2000), 0.0, 450.0 / spb, 0.0, 1.0, 0.5)
    # t7: edited while playing (below)
    t7 = tr(-3.0, -1.0)
    eng.add_clip(t7, smp(8000), 0.0, 1400.0 / spb, 0.0, 1.0, 0.5)
    late = smp(4000)
    eng.play()
    outs = [eng.process(4)]
    eng.add_clip(t7, late, 700.0 / spb, 1000.0 / spb, 0.0, 1.0, 0.9)  # splits the clip that is playing right now
    outs.append(eng.process(8))
    return _collect(eng, outs, n)


def edits(make_engine):
    """Clip editing calls that feed the scheduler: Engine::move_clip (sets internal_state_changed: a clip moved while it
    plays is stopped and restarted at its new offset, track.cpp:394-419), resize_clip (left / right, shift, stretch —
    calc_resize_clip, clip_edit.h:18-126), delete_clip, duplicate_clip, and a move onto a neighbour (reserve_track_region
    with the moved clip ignored). `clip` arguments are indices into the track's clip list ordered by min_time."""
    rng = np.random.RandomState(2468)
    B, rate = 128, 48000
    eng = make_engine(2, B, rate, 120.0)
    spb = rate * 0.5
    n = 0

    def tr(vol=-3.0, pan=0.0):
        nonlocal n
        eng.add_track(vol, pan, False)
        n += 1
        return n - 1

    def smp(frames, ch=2, r=48000):
        return eng.add_sample(_src(rng, ch, frames, 8), r, FMT_F32)

    t0 = tr(-2.0, -0.3)  # moved while it plays
    eng.add_clip(t0, smp(6000), 0.0, 1500.0 / spb, 0.0, 1.0, 0.8)
    t1 = tr(-4.0, 0.4)  # second clip deleted before it starts
    eng.add_clip(t1, smp(3000), 0.0, 400.0 / spb, 0.0, 1.0, 0.7)
    eng.add_clip(t1, smp(3000), 600.0 / spb, 1000.0 / spb, 0.0, 1.0, 0.9)
    t2 = tr(0.0, 0.0)  # duplicated further down the timeline (and onto nothing)
    eng.add_clip(t2, smp(4000), 100.0 / spb, 700.0 / spb, 9.0, 1.0, 1.0)
    eng.duplicate_clip(t2, 0, 900.0 / spb, 1400.0 / spb)
    t3 = tr(-1.0, 0.1)  # left edge resized before playback: content start follows the edge
    eng.add_clip(t3, smp(6000), 0.0, 1200.0 / spb, 50.0, 1.0, 0.6)
    eng.resize_clip(t3, 0, 300.0 / spb, 1200.0 / spb, 1.0 / spb, True)
    t4 = tr(-1.5, -0.8)  # right edge pulled in, then a second clip stretched (speed changes) on its right edge
    eng.add_clip(t4, smp(6000), 0.0, 800.0 / spb, 0.0, 1.0, 0.5)
    eng.resize_clip(t4, 0, -250.0 / spb, 0.0, 1.0 / spb, False)
    eng.add_clip(t4, smp(1000), 700.0 / spb, 1300.0 / spb, 0.0, 1.0, 0.7)
    eng.resize_clip(t4, 1, 150.0 / spb, 700.0 / spb, 1.0 / spb, False, False, True)
    t5 = tr(-7.0, 0.9)  # a clip moved onto its neighbour: the neighbour's head is trimmed
    eng.add_clip(t5, smp(3000), 0.0, 300.0 / spb, 0.0, 1.0, 1.1)
    eng.add_clip(t5, smp(3000), 500.0 / spb, 1100.0 / spb, 4.0, 1.0, 0.9)
    eng.move_clip(t5, 0, 350.0 / spb)
    t6 = tr(-5.0, 0.5)  # left edge resized with shift (content stays put), 44.1 kHz source at speed 1.25
    eng.add_clip(t6, smp(9000, 2, 44100), 200.0 / spb, 1300.0 / spb, 30.0, 1.25, 0.75)
    eng.resize_clip(t6, 0, 120.0 / spb, 1300.0 / spb, 1.0 / spb, True, True)
    eng.play()
    outs = [eng.process(3)]
    eng.move_clip(t0, 0, 200.0 / spb)  # the clip that is playing right now
    eng.delete_clip(t1, 1)
    outs.append(eng.process(4))
    eng.resize_clip(t3, 0, -100.0 / spb, 1200.0 / spb, 1.0 / spb, True)  # playing clip, no shift: plain refresh
    outs.append(eng.process(5))
    return _collect(eng, outs, n)


def tempo(make_engine):
    """Engine::set_bpm while the session plays: beat_duration feeds the transport math of every callback
    (engine.cpp:1578-1585) and the beat -> sample conversions of the clip events (track.cpp:359-361, 423-425); voices that
    are already streaming are untouched. Clips start / stop after the change, one clip is added under the new tempo."""
    rng = np.random.RandomState(1357)
    B, rate = 128, 48000
    eng = make_engine(2, B, rate, 120.0)
    spb = rate * 0.5
    for t in range(4):
        eng.add_track(-3.0 - t, -0.6 + 0.4 * t, False)
    s = [eng.add_sample(_src(rng, 2, 8000, 4), 48000 if i % 2 == 0 else 44100, FMT_F32) for i in range(4)]
    eng.add_clip(0, s[0], 0.0, 2000.0 / spb, 0.0, 1.0, 0.8)            # plays across both tempo changes
    eng.add_clip(1, s[1], 500.0 / spb, 900.0 / spb, 5.0, 1.0, 0.7)     # starts after the first change
    eng.add_clip(2, s[2], 100.0 / spb, 700.0 / spb, 0.0, 1.0, 0.9)     # stops after the first change
    eng.add_clip(3, s[3], 1000.0 / spb, 1500.0 / spb, 0.0, 1.3, 0.6)
    eng.play()
    outs = [eng.process(3)]
    eng.set_bpm(93.7)
    outs.append(eng.process(5))
    eng.add_clip(1, s[0], 0.05, 0.058, 12.0, 1.0, 0.5)                 # added under the new tempo, ahead of the playhead
    eng.set_bpm(171.0)
    outs.append(eng.process(6))
    return _collect(eng, outs, 4)


def mixer(make_engine):
    """Mixer-side calls between callbacks: Engine::set_clip_gain on a clip that is playing (the voice reads the clip gain
    every callback, track.cpp:676,716), solo_track on / other / off (through set_mute + the parameter queue), move_track
    (tracks.cpp order = bus summation order, engine.cpp:1600-1617) and delete_track before playback."""
    rng = np.random.RandomState(8642)
    B, rate = 128, 48000
    eng = make_engine(2, B, rate, 120.0)
    spb = rate * 0.5
    for t in range(6):
        eng.add_track(-3.0 - 1.5 * t, -0.8 + 0.3 * t, False)
        sid = eng.add_sample(_src(rng, 2, 6000, 3), 48000, FMT_F32)
        eng.add_clip(t, sid, (40.0 * t) / spb, 1800.0 / spb, 3.0 * t, 1.0, 0.5 + 0.1 * t)
    eng.delete_track(4)  # tracks 0,1,2,3,5 remain (5 becomes slot 4)
    eng.play()
    outs = [eng.process(2)]
    eng.set_clip_gain(1, 0, 1.7)
    outs.append(eng.process(2))
    eng.solo_track(2)
    outs.append(eng.process(2))
    eng.solo_track(0)  # solo moves to another track
    outs.append(eng.process(2))
    eng.solo_track(0)  # solo off: everything unmuted
    eng.move_track(0, 3)  # summation order changes: the bus differs in the last bits, the per-slot peaks move
    outs.append(eng.process(3))
    eng.move_track(4, 1)
    eng.set_clip_gain(2, 0, 0.05)
    outs.append(eng.process(2))
    return _collect(eng, outs, 5)


def erase(make_engine):
    """Engine::delete_region (engine.cpp:463-473): a time range erased from a track — inside one clip (split), across a
    clip's tail and the next clip's head, over whole clips, over nothing, and under the clip that is playing (its tail
    goes; the voice stops at the new end)."""
    rng = np.random.RandomState(1122)
    B, rate = 128, 48000
    eng = make_engine(2, B, rate, 120.0)
    spb = rate * 0.5
    for t in range(5):
        eng.add_track(-3.0 - t, -0.7 + 0.35 * t, False)

    def smp(frames, r=48000):
        return eng.add_sample(_src(rng, 2, frames, 5), r, FMT_F32)

    eng.add_clip(0, smp(8000), 0.0, 1500.0 / spb, 0.0, 1.0, 0.8)
    eng.delete_region(0, 400.0 / spb, 650.0 / spb)                      # split
    eng.add_clip(1, smp(4000), 0.0, 500.0 / spb, 0.0, 1.0, 0.7)
    eng.add_clip(1, smp(4000, 44100), 600.0 / spb, 1400.0 / spb, 9.0, 1.25, 0.9)
    eng.delete_region(1, 350.0 / spb, 820.0 / spb)                      # tail of one, head of the next (resampled: shifted)
    for i in range(4):
        eng.add_clip(2, smp(3000), (i * 350.0) / spb, (i * 350.0 + 300.0) / spb, 0.0, 1.0, 0.6)
    eng.delete_region(2, 340.0 / spb, 1060.0 / spb)                     # two whole clips
    eng.add_clip(3, smp(3000), 200.0 / spb, 600.0 / spb, 0.0, 1.0, 0.5)
    eng.delete_region(3, 700.0 / spb, 900.0 / spb)                      # nothing there
    eng.delete_region(3, 0.0, 100.0 / spb)                              # nothing there either
    eng.add_clip(4, smp(8000), 0.0, 1500.0 / spb, 0.0, 1.0, 0.8)
    eng.play()
    outs = [eng.process(4)]
    eng.delete_region(4, 800.0 / spb, 2000.0 / spb)                     # the tail of the clip that is playing
    outs.append(eng.process(8))
    return _collect(eng, outs, 5)


def params(make_engine):
    """Volume / pan / mute changes between callbacks, not-playing callbacks, stop/play (track.cpp:618-643)."""
    rng = np.random.RandomState(99)
    eng = make_engine(2, 64, 44100, 97.0)
    for t in range(5):
        eng.add_track(-3.0 * t, -0.5 + 0.25 * t, False)
        sid = eng.add_sample(_src(rng, 2, 6000, 5), 44100 if t % 2 == 0 else 48000, FMT_F32)
        eng.add_clip(t, sid, 0.0, 8.0, 0.0, 1.0, 0.9)
    outs = [eng.process(2)]  # not playing: silence, params applied
    eng.play()
    outs.append(eng.process(3))
    eng.set_volume(1, 2.5)
    eng.set_pan(2, 1.0)
    eng.set_mute(3, True)
    outs.append(eng.process(2))
    eng.set_mute(3, False)
    eng.set_pan(0, -1.0)
    eng.set_volume(4, -71.9)
    outs.append(eng.process(2))
    eng.stop()
    outs.append(eng.process(1))
    eng.play()
    outs.append(eng.process(3))
    return _collect(eng, outs, 5)


def hot_clamp(make_engine):
    """Deliberately hot mix so the +/-1 clamp (engine.cpp:1627-1636) is exercised on many samples."""
    rng = np.random.RandomState(5)
    eng = make_engine(2, 512, 48000, 120.0)
    for t in range(12):
        eng.add_track(6.0, 0.0, False)
        sid = eng.add_sample((rng.uniform(-1, 1, size=(2, 4096))).astype(np.float32), 48000, FMT_F32)
        eng.add_clip(t, sid, 0.0, 16.0, 0.0, 1.0, 1.0)
    eng.play()
    return _collect(eng, [eng.process(3)], 12)


def ragged(make_engine):
    """Odd block size, mono output bus, zero-track engine is covered by `empty`."""
    rng = np.random.RandomState(31)
    eng = make_engine(1, 100, 32000, 133.0)
    for t in range(7):
        eng.add_track(-2.0 * t, 0.3 * (t - 3), False)
        sid = eng.add_sample(_src(rng, 1 + (t % 2), 3000, 7), 32000 if t % 3 else 22050, FMT_F32)
        # mono bus: only src channel 0 is ever indexed, so mono + resample is well defined here
        eng.add_clip(t, sid, 0.01 * t, 3.0, 2.0 * t, 1.0, 0.8)
    eng.play()
    return _collect(eng, [eng.process(9)], 7)


def empty(make_engine):
    """No tracks at all, then tracks without clips."""
    eng = make_engine(2, 512, 48000, 120.0)
    eng.play()
    a = eng.process(2)
    eng2 = make_engine(2, 512, 48000, 120.0)
    for t in range(3):
        eng2.add_track(0.0, 0.0, False)
    eng2.play()
    b = eng2.process(2)
    return dict(out=np.concatenate([a[0], b[0]]), peaks=b[1])


def fades(make_engine):
    """EXTENSION scenario (parity unpinned w.r.t. whitebox — the reference has no fades): fade-in / fade-out
    ramps on unity, resampled, mono and int16 clips, ramps longer than the clip, a mid-clip start inside a ramp."""
    rng = np.random.RandomState(606)
    B, rate = 256, 48000
    eng = make_engine(2, B, rate, 120.0)
    spb = rate * 0.5
    cfgs = [  # (channels, src_rate, fmt, start_frame, length_frames, speed, fade_in_frames, fade_out_frames)
        (2, 48000, FMT_F32, 0, 2000, 1.0, 700, 500),
        (2, 44100, FMT_F32, 100, 1800, 1.0, 300, 0),
        (1, 48000, FMT_F32, 37, 1500, 1.0, 0, 1000),
        (2, 48000, FMT_I16, 0, 900, 1.0, 2000, 2000),   # ramps longer than the clip: they overlap
        (2, 48000, FMT_F32, 256, 1024, 0.5, 256, 256),
        (2, 96000, FMT_F32, 10, 2200, 1.0, 1, 3),
    ]
    for t, (ch, sr, fmt, start, length, speed, fi, fo) in enumerate(cfgs):
        eng.add_track(-2.0 - t, -0.6 + 0.25 * t, False)
        sid = eng.add_sample(_src(rng, ch, 6000, 6, fmt), sr, fmt)
        eng.add_clip(t, sid, start / spb, (start + length) / spb, 3.0, speed, 0.9, fi / spb, fo / spb)
    eng.play()
    outs = [eng.process(5), eng.process(6)]
    eng.stop()
    eng.set_playhead(333.0 / spb)  # inside several fade-ins
    eng.play()
    outs.append(eng.process(4))
    return _collect(eng, outs, len(cfgs))


def effects(make_engine, fxp):
    """EXTENSION scenario (parity unpinned w.r.t. whitebox): BASELINE cfg 4 shape at test size — tracks with a
    4-band EQ + compressor chain next to plain tracks, chain state carried across renders, one chain removed
    mid-session. fxp(eq=..., threshold_db=..., ratio_code=...) builds the parameter struct."""
    rng = np.random.RandomState(808)
    B, rate = 256, 48000
    eng = make_engine(2, B, rate, 120.0)
    n = 10
    for t in range(n):
        eng.add_track(-3.0 - t, -0.8 + 0.18 * t, False)
        fmt = FMT_I16 if t == 7 else FMT_F32
        sid = eng.add_sample(_src(rng, 1 if t == 5 else 2, 9000, 4, fmt), 44100 if t % 4 == 3 else 48000, fmt)
        eng.add_clip(t, sid, (t % 3) * 0.01, 8.0, float(t), 1.0, 0.9, 0.02 if t == 2 else 0.0, 0.0)
    eq_a = ((120.0, 4.0, 0.7), (800.0, -6.0, 1.2), (2500.0, 3.0, 2.0), (8000.0, 5.0, 0.7))
    eq_b = ((80.0, -3.0, 0.9), (400.0, 2.0, 0.8), (5000.0, -4.0, 1.5), (12000.0, 2.5, 0.6))
    eng.set_effects(0, fxp(eq=eq_a, threshold_db=-30.0, ratio_code=2, attack_ms=2.0, release_ms=60.0, makeup_db=3.0))
    eng.set_effects(2, fxp(eq=eq_b))                                     # EQ only
    eng.set_effects(3, fxp(threshold_db=-36.0, ratio_code=4))             # limiter only, resampled source
    eng.set_effects(5, fxp(eq=eq_a, threshold_db=-40.0, ratio_code=1))    # mono source
    eng.set_effects(7, fxp(eq=eq_b, threshold_db=-32.0, ratio_code=3))    # int16 source
    eng.play()
    outs = [eng.process(3), eng.process(4)]
    eng.set_effects(2, None)  # chain removed: the track is the reference path again
    eng.set_effects(9, fxp(eq=eq_a, threshold_db=-28.0, ratio_code=2))
    outs.append(eng.process(3))
    return _collect(eng, outs, n)


def effects_shapes(make_engine, fxp, n_tracks=24, block=512, n_blocks=5, out_channels=2, loud=True, chunks=None):
    """EXTENSION scenario (parity unpinned w.r.t. whitebox): the chain's time-parallel evaluation at arbitrary shapes —
    block sizes that leave partial segments / partial look-ahead blocks / several 512-frame chunks per callback, mono
    bus, attack slower than release (the follower then takes the SMALLER candidate), EQ-only and compressor-only chains,
    every ratio code, silent gaps between clips, signals well above the threshold."""
    rng = np.random.RandomState(4242 + n_tracks + block)
    eng = make_engine(out_channels, block, 48000, 120.0)
    frames = (n_blocks + 3) * block + 64
    eq_a = ((120.0, 4.0, 0.7), (800.0, -6.0, 1.2), (2500.0, 3.0, 2.0), (8000.0, 5.0, 0.7))
    eq_b = ((60.0, 6.0, 1.1), (300.0, -9.0, 3.0), (6000.0, 4.0, 0.5), (15000.0, -5.0, 0.9))
    spb = 48000 * 0.5
    for t in range(n_tracks):
        eng.add_track(-3.0 - (t % 9), -0.9 + 0.17 * (t % 11), False)
        scale = 1 if not loud else max(1, n_tracks // 8)
        sid = eng.add_sample(_src(rng, 2, frames, scale), 48000)
        start = (t % 5) * 7 / spb if t % 4 == 1 else 0.0  # some clips start mid-block
        end = (n_blocks - 1) * block / spb if t % 6 == 2 else 1e6  # some end before the render does
        eng.add_clip(t, sid, start, end, float(t % 3), 1.0, 0.9)
        kind = t % 6
        if kind == 0:
            eng.set_effects(t, fxp(eq=eq_a, threshold_db=-30.0, ratio_code=2, attack_ms=2.0, release_ms=60.0, makeup_db=3.0))
        elif kind == 1:
            eng.set_effects(t, fxp(eq=eq_b))
        elif kind == 2:
            eng.set_effects(t, fxp(threshold_db=-40.0, ratio_code=4, attack_ms=0.1, release_ms=5.0))
        elif kind == 3:
            eng.set_effects(t, fxp(eq=eq_b, threshold_db=-36.0, ratio_code=1, attack_ms=30.0, release_ms=3.0, makeup_db=-2.0))
        elif kind == 4:
            eng.set_effects(t, fxp(eq=eq_a, threshold_db=-33.0, ratio_code=3, attack_ms=1.0, release_ms=200.0))
        # kind 5: no chain
    eng.play()
    outs = [eng.process(n) for n in (chunks or [n_blocks])]
    return _collect(eng, outs, n_tracks)


def reverb(make_engine, fxp, taps=777, B=256, out_channels=2):
    """EXTENSION scenario (parity unpinned w.r.t. whitebox): BASELINE cfg 5 shape at test size — tracks whose chain
    ends in a convolution with a shared impulse response, history carried across renders."""
    rng = np.random.RandomState(909)
    rate = 48000
    eng = make_engine(out_channels, B, rate, 120.0)
    ir = (rng.uniform(-1, 1, taps) * np.exp(-np.arange(taps) / (taps / 5.0)) * 0.2).astype(np.float32)
    ir[0] = 1.0
    eng.set_impulse_response(ir)
    n = 5
    for t in range(n):
        eng.add_track(-6.0 - t, -0.5 + 0.25 * t, False)
        sid = eng.add_sample(_src(rng, 2, 6000, 4), 48000)
        eng.add_clip(t, sid, 0.0, 8.0, 0.0, 1.0, 0.9)
    eq = ((120.0, 4.0, 0.7), (800.0, -6.0, 1.2), (2500.0, 3.0, 2.0), (8000.0, 5.0, 0.7))
    eng.set_effects(0, fxp(reverb=True))
    eng.set_effects(2, fxp(eq=eq, threshold_db=-30.0, ratio_code=2, reverb=True))
    eng.set_effects(3, fxp(eq=eq))
    eng.play()
    outs = [eng.process(3), eng.process(1), eng.process(4)]
    return _collect(eng, outs, n)


def reverb_full_depth(make_engine, fxp, taps=65536, n_tracks=64, n_blocks=160, block=512, chunks=None, seed=5150):
    """EXTENSION scenario (parity unpinned w.r.t. whitebox): BASELINE cfg 5 at FULL accumulation depth — more frames of
    non-zero signal than the impulse response has taps (160 x 512 = 81920 > 65536) on >= 128 signals, so every tap chunk
    of the tensor-core path multiplies real history and the accumulators see the whole chain. Returns the session's
    arrays plus what an f64 evaluation needs (sources, gains, impulse response)."""
    rng = np.random.RandomState(seed)
    eng = make_engine(2, block, 48000, 120.0)
    ir = (rng.standard_normal(taps) * np.exp(-np.arange(taps) / (taps / 6.0)) * 0.01).astype(np.float32)
    ir[0] = 1.0
    eng.set_impulse_response(ir)
    frames = (n_blocks + 2) * block
    srcs, params = [], []
    for t in range(n_tracks):
        vol, pan, gain = -6.0 - (t % 7), -1.0 + 0.2 * (t % 11), float(np.float32(0.5 + 0.001 * (t % 512)))
        eng.add_track(vol, pan, False)
        x = _src(rng, 2, frames, n_tracks)
        sid = eng.add_sample(x, 48000)
        eng.add_clip(t, sid, 0.0, 1e6, 0.0, 1.0, gain)
        eng.set_effects(t, fxp(reverb=True))
        srcs.append(x)
        params.append((vol, pan, gain))
    eng.play()
    outs = [eng.process(n) for n in (chunks or [n_blocks])]
    res = _collect(eng, outs, n_tracks)
    res["_ir"], res["_srcs"], res["_params"] = ir, srcs, params
    return res


def reverb_f64_expected(res, panning_coefs, db_to_linear, block=512):
    """The specification of the convolution reverb (oracle/wb_oracle.c apply_reverb: y[n] = (float) sum_k h[k] x[n-k],
    accumulated in f64) evaluated with an f64 FFT — the direct sum is ~1e15 multiply-adds at this size; the FFT evaluation
    differs from it by ~1e-13 relative (tests/test_oracle.py pins that on a size the direct sum can do). -> the clamped
    bus [K][2][B] and the per-callback VU peaks [K][N][2] the session should produce."""
    ir, srcs, params = res["_ir"].astype(np.float64), res["_srcs"], res["_params"]
    K = res["out"].shape[0]
    n_out = K * block
    nfft = 1 << int(np.ceil(np.log2(n_out + ir.size)))
    H = np.fft.rfft(ir, nfft)
    bus = np.zeros((2, n_out), np.float64)
    peaks = np.zeros((K, len(srcs), 2), np.float32)
    for t, (x, (vol, pan, gain)) in enumerate(zip(srcs, params)):
        pl, pr = panning_coefs(pan)
        v = np.float32(db_to_linear(vol))
        for c, pc in ((0, pl), (1, pr)):
            xin = (x[c][:n_out] * np.float32(gain)).astype(np.float32).astype(np.float64)  # Sampler::stream: src * clip gain
            y = np.fft.irfft(np.fft.rfft(xin, nfft) * H, nfft)[:n_out].astype(np.float32)  # the chain's output
            term = (y * np.float32(v * np.float32(pc))).astype(np.float32)                 # apply_gain: * (volume * pan)
            peaks[:, t, c] = np.abs(term).reshape(K, block).max(axis=1)
            bus[c] += term
    out = np.clip(bus, -1.0, 1.0).astype(np.float32)
    return np.ascontiguousarray(out.reshape(2, K, block).transpose(1, 0, 2)), peaks


def polyphase(make_engine, fxp=None):
    """EXTENSION scenario (parity unpinned w.r.t. whitebox): BASELINE cfg 3 wording — 44.1 -> 48 kHz through the
    polyphase windowed-sinc resampler — plus other ratios, clip starts at frame 0 (taps reach before the sample),
    sample exhaustion (taps reach past the end), a fade on a resampled clip, mono / int16 sources (stay linear)."""
    rng = np.random.RandomState(1212)
    B, rate = 256, 48000
    eng = make_engine(2, B, rate, 120.0)
    eng.set_resampler(1)
    spb = rate * 0.5
    cfgs = [  # (channels, src_rate, fmt, start_frame, speed, start_offset, frames, fade_in)
        (2, 44100, FMT_F32, 0, 1.0, 0.0, 6000, 0.0),
        (2, 44100, FMT_F32, 100, 1.0, 33.0, 6000, 200.0),
        (2, 96000, FMT_F32, 0, 1.0, 0.0, 3000, 0.0),      # exhausts
        (2, 48000, FMT_F32, 37, 0.5, 5.0, 4000, 0.0),
        (2, 48000, FMT_F32, 0, 1.3, 0.0, 5000, 0.0),
        (1, 44100, FMT_F32, 0, 1.0, 0.0, 6000, 0.0),      # mono: linear
        (2, 44100, FMT_I16, 0, 1.0, 0.0, 6000, 0.0),      # int16: linear
        (2, 48000, FMT_F32, 0, 1.0, 0.0, 6000, 0.0),      # unity: copy
    ]
    for t, (ch, sr, fmt, start, speed, off, frames, fi) in enumerate(cfgs):
        eng.add_track(-2.0 - t, -0.6 + 0.2 * t, False)
        sid = eng.add_sample(_src(rng, ch, frames, 6, fmt), sr, fmt)
        eng.add_clip(t, sid, start / spb, 8.0, off, speed, 0.9, fi / spb, 0.0)
    if fxp is not None:  # one resampled track through the effect path as well
        eng.set_effects(1, fxp(threshold_db=-30.0, ratio_code=1))
    eng.play()
    outs = [eng.process(4), eng.process(5)]
    return _collect(eng, outs, len(cfgs))


def fuzz(make_engine, seed):
    """Random session: random rates / formats / speeds / clip layouts / block size, params changed mid-run."""
    rng = np.random.RandomState(1000 + seed)
    B = int(rng.choice([32, 64, 100, 128, 256, 512]))
    rate = int(rng.choice([44100, 48000, 96000]))
    bpm = float(rng.choice([90.0, 120.0, 133.3, 150.0]))
    eng = make_engine(2, B, rate, bpm)
    spb = rate * 60.0 / bpm
    n_tracks = int(rng.randint(1, 12))
    n_blocks = int(rng.randint(6, 14))
    total = n_blocks * B
    for t in range(n_tracks):
        eng.add_track(float(rng.uniform(-30, 6)), float(rng.uniform(-1, 1)), bool(rng.rand() < 0.1))
        pos = float(rng.randint(0, max(1, total // 3)))
        for _ in range(int(rng.randint(0, 4))):
            fmt = int(rng.choice([FMT_F32, FMT_F32, FMT_I16, FMT_I24, FMT_I32]))
            srate = int(rng.choice([rate, rate, 44100, 22050, 96000]))
            speed = float(rng.choice([1.0, 1.0, 0.5, 2.0, 0.91875, 1.3]))
            frames = int(rng.randint(16, 3000))
            length = float(rng.randint(1, total // 2 + 2))
            sid = eng.add_sample(_src(rng, 2, frames, n_tracks, fmt), srate, fmt)
            eng.add_clip(t, sid, pos / spb, (pos + length) / spb, float(rng.randint(0, 40)), speed,
                         float(rng.uniform(0.1, 1.5)))
            pos += length + float(rng.choice([0.0, 0.0, rng.randint(1, 200)]))
    eng.play()
    outs = []
    done = 0
    while done < n_blocks:
        n = int(min(n_blocks - done, rng.randint(1, 5)))
        outs.append(eng.process(n))
        done += n
        t = int(rng.randint(0, n_tracks))
        which = rng.randint(0, 3)
        if which == 0:
            eng.set_volume(t, float(rng.uniform(-20, 3)))
        elif which == 1:
            eng.set_pan(t, float(rng.uniform(-1, 1)))
        else:
            eng.set_mute(t, bool(rng.rand() < 0.5))
    return _collect(eng, outs, n_tracks)


def fuzz_overlap(make_engine, seed):
    """Random session whose clips are dropped anywhere on the timeline — on top of one another, inside one another, across
    several — so every add goes through add_to_cliplist / reserve_track_region (engine.cpp:409-461, 478-569); a third of
    the sessions add more clips while playing."""
    rng = np.random.RandomState(5000 + seed)
    B = int(rng.choice([32, 64, 128, 256]))
    rate = int(rng.choice([44100, 48000]))
    bpm = float(rng.choice([90.0, 120.0, 150.0]))
    eng = make_engine(2, B, rate, bpm)
    spb = rate * 60.0 / bpm
    n_tracks = int(rng.randint(1, 6))
    n_blocks = int(rng.randint(6, 14))
    total = n_blocks * B
    samples = []

    def drop_clip(t):
        srate = int(rng.choice([rate, rate, 44100, 96000]))
        speed = float(rng.choice([1.0, 1.0, 1.0, 0.5, 1.3]))
        if not samples or rng.rand() < 0.5:
            samples.append((eng.add_sample(_src(rng, 2, int(rng.randint(200, 6000)), n_tracks), srate, FMT_F32), srate))
        sid, _ = samples[int(rng.randint(0, len(samples)))]
        a = float(rng.randint(0, total))
        b = a + float(rng.randint(1, total // 2 + 2))
        # distinct fractional parts keep min_time ties (an unstable sort in the reference) out of the fixture
        a += float(rng.randint(1, 1000)) / 1024.0
        eng.add_clip(t, sid, a / spb, b / spb, float(rng.randint(0, 40)), speed, float(rng.uniform(0.1, 1.2)))

    for t in range(n_tracks):
        eng.add_track(float(rng.uniform(-20, 3)), float(rng.uniform(-1, 1)), False)
        for _ in range(int(rng.randint(1, 7))):
            drop_clip(t)
    eng.play()
    outs = []
    done = 0
    live_edits = seed % 3 == 0
    while done < n_blocks:
        n = int(min(n_blocks - done, rng.randint(1, 5)))
        outs.append(eng.process(n))
        done += n
        if live_edits and done < n_blocks:
            drop_clip(int(rng.randint(0, n_tracks)))
    return _collect(eng, outs, n_tracks)


def fuzz_edits(make_engine, seed):
    """Random sessions edited between callbacks: move / resize (left, right, shift) / delete / duplicate / add, some of
    them on the clip that is playing. (stretch is exercised by `edits`; random stretches drive the speed to absurd
    values.)"""
    rng = np.random.RandomState(7000 + seed)
    B = int(rng.choice([32, 64, 128]))
    rate = int(rng.choice([44100, 48000]))
    eng = make_engine(2, B, rate, 120.0)
    spb = rate * 0.5
    n_tracks = int(rng.randint(1, 5))
    n_blocks = int(rng.randint(8, 16))
    total = n_blocks * B
    sids = []
    for _ in range(3):
        srate = int(rng.choice([rate, 44100, 96000]))
        sids.append(eng.add_sample(_src(rng, 2, int(rng.randint(500, 8000)), n_tracks), srate, FMT_F32))

    def frac():
        return float(rng.randint(1, 1000)) / 1024.0  # keeps min_time ties out of the fixture

    def drop_clip(t):
        a = float(rng.randint(0, total)) + frac()
        b = a + float(rng.randint(8, total // 2 + 9))
        eng.add_clip(t, sids[int(rng.randint(0, 3))], a / spb, b / spb, float(rng.randint(0, 40)),
                     float(rng.choice([1.0, 1.0, 0.5, 1.3])), float(rng.uniform(0.1, 1.2)))

    for t in range(n_tracks):
        eng.add_track(float(rng.uniform(-20, 3)), float(rng.uniform(-1, 1)), False)
        for _ in range(int(rng.randint(1, 5))):
            drop_clip(t)
    eng.play()
    outs = []
    done = 0
    while done < n_blocks:
        n = int(min(n_blocks - done, rng.randint(1, 4)))
        outs.append(eng.process(n))
        done += n
        if done >= n_blocks:
            break
        t = int(rng.randint(0, n_tracks))
        nc = eng.clip_count(t)
        op = int(rng.randint(0, 8))
        if op == 6:  # tempo change: later edits shift clip content with the new beat duration
            eng.set_bpm(float(rng.choice([90.0, 120.0, 133.3, 150.0])))
            continue
        if op == 7:  # erase a time range
            a = float(rng.randint(0, total)) + frac()
            eng.delete_region(t, a / spb, (a + float(rng.randint(8, total // 3 + 9))) / spb)
            continue
        if nc == 0 or op == 0:
            drop_clip(t)
            continue
        c = int(rng.randint(0, nc))
        lo, hi = eng.clip_range(t, c)  # the UI passes the clip's own opposite edge as the resize limit (timeline.cpp:1428,1436)
        if op == 1:
            eng.move_clip(t, c, (float(rng.randint(-total // 3, total // 3)) + frac()) / spb)
        elif op == 2:
            eng.resize_clip(t, c, (float(rng.randint(-200, 200)) + frac()) / spb, lo, 4.0 / spb, False, bool(rng.rand() < 0.3))
        elif op == 3:
            eng.resize_clip(t, c, (float(rng.randint(-200, 200)) + frac()) / spb, hi, 4.0 / spb, True, bool(rng.rand() < 0.3))
        elif op == 4:
            eng.delete_clip(t, c)
        else:
            a = float(rng.randint(0, total)) + frac()
            eng.duplicate_clip(t, c, a / spb, (a + float(rng.randint(8, total // 3 + 9))) / spb)
    return _collect(eng, outs, n_tracks)


MIP_CASES = dict(f32s=(FMT_F32, 20011, 2), i16m=(FMT_I16, 4100, 1), i32s=(FMT_I32, 777, 2))


def mip_source(fmt, frames, ch):
    data = _src(np.random.RandomState(frames), ch, frames, 1, fmt)
    if fmt == FMT_F32:
        data[:, :7] = [1, -1, 0, 0.5, -0.5, 1, -1]
    return data


EXT = dict(fades=fades, effects=effects, reverb=reverb, polyphase=polyphase)  # builder-specified extensions: checked against the C port only

def plugin_silences(make_engine):
    """a plugin in a track's slot (engine/track.h:124). Track::process then renders the clips into the
    plugin's effect_buffer, which is never mixed (track.cpp:600,645-724): the track contributes only what the plugin
    writes — nothing for the no-op plugin the compiled reference is driven with — while its scheduler and sampler keep
    running, so the clip is where it should be when the plugin is removed again."""
    rng = np.random.RandomState(1111)
    B, rate = 128, 48000
    eng = make_engine(2, B, rate, 120.0)
    spb = rate * 0.5
    for t in range(5):
        eng.add_track(-3.0 - t, -0.7 + 0.35 * t, False)
        sid = eng.add_sample(_src(rng, 2, 5000, 3), 44100 if t == 3 else 48000, FMT_F32)
        eng.add_clip(t, sid, (30.0 * t) / spb, 2600.0 / spb, 2.0 * t, 1.0, 0.6 + 0.1 * t)
    eng.set_plugin(1, True)  # before playback starts
    eng.play()
    outs = [eng.process(4)]
    eng.set_plugin(3, True)  # while its (resampled) clip plays
    outs.append(eng.process(5))
    eng.set_plugin(1, False)  # the clip resumes where the transport is, not where it stopped being heard
    outs.append(eng.process(5))
    eng.set_plugin(3, False)
    eng.set_plugin(0, True)
    outs.append(eng.process(8))  # past the clips' ends
    return _collect(eng, outs, 5)


def reconfigure(make_engine):
    """the audio device changes under a live session (config.cpp:198-232 -> Engine::set_audio_channel_config
    again, app.cpp:263-264): new block size, then a new sample rate and channel count, with tracks, clips, resident
    samples, the transport and playing voices persisting."""
    rng = np.random.RandomState(2222)
    eng = make_engine(2, 256, 48000, 120.0)
    for t in range(6):
        eng.add_track(-4.0 - t, -0.9 + 0.36 * t, False)
        sid = eng.add_sample(_src(rng, 2 if t != 2 else 1, 30000, 4, FMT_I16 if t == 4 else FMT_F32),
                             44100 if t % 3 == 1 else 48000, FMT_I16 if t == 4 else FMT_F32)
        eng.add_clip(t, sid, 0.01 * t, 1.2 + 0.1 * t, float(t), 1.0, 0.8)
    eng.play()
    parts = [eng.process(5)]
    eng.configure(2, 96, 48000)     # smaller block, mid-playback
    parts.append(eng.process(9))
    eng.configure(2, 512, 44100)    # new rate: running voices keep their speed until their next event (sampler.h:18-27)
    parts.append(eng.process(4))
    eng.stop()
    eng.configure(1, 200, 96000)    # mono bus, stopped
    eng.play()
    parts.append(eng.process(6))
    # block sizes differ between the parts: flatten each part's arrays instead of stacking them
    res = {}
    for i, (o, p) in enumerate(parts):
        res["out%d" % i] = o
        res["peaks%d" % i] = p
    res["sampler_offsets"] = np.array([eng.sampler_offset(t) for t in range(6)], np.float64)
    res["transport"] = np.array([eng.sample_position(), eng.playhead()], np.float64)
    return res


ALL = dict(kat=kat, cfg1=cfg1, cfg2_small=cfg2_small, cfg3_small=cfg3_small, int_formats=int_formats,
           event_split=event_split, overlaps=overlaps, edits=edits, tempo=tempo, mixer=mixer, erase=erase, params=params, hot_clamp=hot_clamp, ragged=ragged, empty=empty,
           plugin_silences=plugin_silences, reconfigure=reconfigure)
