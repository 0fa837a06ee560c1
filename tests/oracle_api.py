"""ctypes wrapper over the CPU checkers (oracle/wbo.h). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Two libraries export the same wbo_* symbols:
  oracle/_ref/libwbref.so  - the reference's own Engine::process (built from /root/reference/src)
  oracle/liboracle.so      - the plain-C restatement (oracle/wb_oracle.c)
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libwbref.so")
PORT_SO = os.path.join(ROOT, "oracle", "liboracle.so")
# the reference's engine with its sample loops replaced by the wbx C ABI (oracle/patch_ref_gpu.py)
REF_GPU_SO = os.path.join(ROOT, "oracle", "_ref", "libwbref_gpu.so")

FMT_I16, FMT_I24, FMT_I24_X8, FMT_I32, FMT_F32 = 3, 5, 6, 7, 9
_NP = {FMT_I16: np.int16, FMT_I24: np.int32, FMT_I32: np.int32, FMT_F32: np.float32}

_libs = {}


def _load(path):
    if path in _libs:
        return _libs[path]
    lib = C.CDLL(path)
    vp, u32, u64, dbl, flt, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_double, C.c_float, C.c_int
    lib.wbo_create.restype = vp
    lib.wbo_create.argtypes = [u32, u32, u32, dbl]
    lib.wbo_destroy.argtypes = [vp]
    lib.wbo_kind.restype = C.c_char_p
    lib.wbo_add_track.argtypes = [vp, flt, flt, i32]
    lib.wbo_set_volume.argtypes = [vp, i32, flt]
    lib.wbo_set_pan.argtypes = [vp, i32, flt]
    lib.wbo_set_mute.argtypes = [vp, i32, i32]
    lib.wbo_add_sample.argtypes = [vp, i32, u32, u64, u32, C.POINTER(vp)]
    lib.wbo_add_clip.argtypes = [vp, i32, i32, dbl, dbl, dbl, dbl, flt]
    lib.wbo_add_clip_fade.argtypes = [vp, i32, i32, dbl, dbl, dbl, dbl, flt, dbl, dbl]
    lib.wbo_set_bpm.argtypes = [vp, dbl]
    lib.wbo_set_bpm.restype = None
    lib.wbo_set_clip_gain.argtypes = [vp, i32, i32, flt]
    for f in ("wbo_solo_track", "wbo_delete_track"):
        getattr(lib, f).argtypes = [vp, i32]
        getattr(lib, f).restype = None
    lib.wbo_move_track.argtypes = [vp, i32, i32]
    lib.wbo_move_track.restype = None
    lib.wbo_clip_count.argtypes = [vp, i32]
    lib.wbo_clip_range.argtypes = [vp, i32, i32, C.POINTER(dbl), C.POINTER(dbl)]
    lib.wbo_move_clip.argtypes = [vp, i32, i32, dbl]
    lib.wbo_resize_clip.argtypes = [vp, i32, i32, dbl, dbl, dbl, i32, i32, i32]
    lib.wbo_delete_clip.argtypes = [vp, i32, i32]
    lib.wbo_duplicate_clip.argtypes = [vp, i32, i32, dbl, dbl]
    lib.wbo_delete_region.argtypes = [vp, i32, dbl, dbl]
    lib.wbo_set_effects.argtypes = [vp, i32, vp]
    lib.wbo_set_resampler.argtypes = [vp, i32]
    lib.wbo_set_plugin.argtypes = [vp, i32, i32]
    lib.wbo_reconfigure.argtypes = [vp, u32, u32, u32]
    lib.wbo_set_impulse_response.argtypes = [vp, vp, u32]
    lib.wbo_set_playhead.argtypes = [vp, dbl]
    lib.wbo_play.argtypes = [vp]
    lib.wbo_stop.argtypes = [vp]
    lib.wbo_process.argtypes = [vp, u32, vp, vp]
    lib.wbo_time_process.argtypes = [vp, u32]
    lib.wbo_time_process.restype = dbl
    for f in ("wbo_sampler_offset",):
        getattr(lib, f).argtypes = [vp, i32]
        getattr(lib, f).restype = dbl
    for f in ("wbo_sample_position", "wbo_playhead"):
        getattr(lib, f).argtypes = [vp]
        getattr(lib, f).restype = dbl
    lib.wbo_panning_coefs.argtypes = [flt, C.POINTER(flt), C.POINTER(flt)]
    lib.wbo_db_to_linear.argtypes = [flt]
    lib.wbo_db_to_linear.restype = flt
    lib.wbo_interleave.argtypes = [vp, C.POINTER(vp), u32, u32, u32, i32]
    lib.wbo_mipmap.argtypes = [vp, i32, i32, i32, vp, u64, C.POINTER(u32)]
    _libs[path] = lib
    return lib


def have_ref():
    return os.path.exists(REF_SO)


def have_port():
    return os.path.exists(PORT_SO)


def have_ref_gpu():
    return os.path.exists(REF_GPU_SO)


def lib(kind):
    return _load({"reference": REF_SO, "reference_gpu": REF_GPU_SO}.get(kind, PORT_SO))


class Session:
    """One engine session on a CPU checker; mirrors the reference's editing API."""

    def __init__(self, kind, out_channels=2, block=512, rate=48000, bpm=120.0):
        self.lib = lib(kind)
        self.kind = kind
        self.C, self.B, self.rate = out_channels, block, rate
        self.h = self.lib.wbo_create(out_channels, block, rate, bpm)
        self.n_tracks = 0
        self._keep = []
        self._channels = {}

    def close(self):
        if self.h:
            self.lib.wbo_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def add_track(self, volume_db=0.0, pan=0.0, mute=False):
        self.n_tracks += 1
        return self.lib.wbo_add_track(self.h, volume_db, pan, int(mute))

    def set_volume(self, t, db):
        self.lib.wbo_set_volume(self.h, t, db)

    def set_pan(self, t, pan):
        self.lib.wbo_set_pan(self.h, t, pan)

    def set_mute(self, t, m):
        self.lib.wbo_set_mute(self.h, t, int(m))

    def add_sample(self, data, rate, fmt=FMT_F32):
        """data: [channels][frames] array of the format's dtype."""
        data = np.ascontiguousarray(data, dtype=_NP[fmt])
        ch, frames = data.shape
        ptrs = (C.c_void_p * ch)(*[data[c].ctypes.data for c in range(ch)])
        sid = self.lib.wbo_add_sample(self.h, fmt, ch, frames, rate, ptrs)
        self._channels[sid] = ch
        return sid

    def add_clip(self, track, sample, min_beat, max_beat, start_offset=0.0, speed=1.0, gain=1.0, fade_start=0.0,
                 fade_end=0.0):
        if fade_start or fade_end:
            return self.lib.wbo_add_clip_fade(self.h, track, sample, min_beat, max_beat, start_offset, speed, gain,
                                              fade_start, fade_end)
        return self.lib.wbo_add_clip(self.h, track, sample, min_beat, max_beat, start_offset, speed, gain)

    def set_bpm(self, bpm):
        self.lib.wbo_set_bpm(self.h, bpm)

    def set_clip_gain(self, track, clip, gain):
        return self.lib.wbo_set_clip_gain(self.h, track, clip, gain)

    def solo_track(self, track):
        self.lib.wbo_solo_track(self.h, track)

    def move_track(self, a, b):
        self.lib.wbo_move_track(self.h, a, b)

    def delete_track(self, track):
        self.lib.wbo_delete_track(self.h, track)
        self.n_tracks -= 1

    def clip_count(self, track):
        return self.lib.wbo_clip_count(self.h, track)

    def clip_range(self, track, clip):
        a, b = C.c_double(), C.c_double()
        assert self.lib.wbo_clip_range(self.h, track, clip, C.byref(a), C.byref(b)) == 0
        return a.value, b.value

    def move_clip(self, track, clip, relative_pos):
        return self.lib.wbo_move_clip(self.h, track, clip, relative_pos)

    def resize_clip(self, track, clip, relative_pos, resize_limit, min_length, left_side, shift=False, stretch=False):
        return self.lib.wbo_resize_clip(self.h, track, clip, relative_pos, resize_limit, min_length, int(left_side),
                                        int(shift), int(stretch))

    def delete_clip(self, track, clip):
        return self.lib.wbo_delete_clip(self.h, track, clip)

    def duplicate_clip(self, track, clip, min_beat, max_beat):
        return self.lib.wbo_duplicate_clip(self.h, track, clip, min_beat, max_beat)

    def delete_region(self, track, min_beat, max_beat):
        return self.lib.wbo_delete_region(self.h, track, min_beat, max_beat)

    def set_effects(self, track, params):
        """params: a ctypes struct laid out like wbo_effects (whitebox_b200.EffectParams) or None."""
        return self.lib.wbo_set_effects(self.h, track, C.byref(params) if params is not None else None)

    def set_impulse_response(self, h):
        if h is None:
            return self.lib.wbo_set_impulse_response(self.h, None, 0)
        h = np.ascontiguousarray(h, np.float32)
        return self.lib.wbo_set_impulse_response(self.h, h.ctypes.data, h.size)

    def set_resampler(self, mode):
        self.lib.wbo_set_resampler(self.h, mode)

    def set_plugin(self, track, present=True):
        assert self.lib.wbo_set_plugin(self.h, track, int(present)) == 0

    def configure(self, out_channels, block, rate):
        assert self.lib.wbo_reconfigure(self.h, out_channels, block, rate) == 0
        self.C, self.B, self.rate = out_channels, block, rate

    def set_playhead(self, beat):
        self.lib.wbo_set_playhead(self.h, beat)

    def play(self):
        self.lib.wbo_play(self.h)

    def stop(self):
        self.lib.wbo_stop(self.h)

    def process(self, n_blocks):
        out = np.zeros((n_blocks, self.C, self.B), np.float32)
        peaks = np.zeros((n_blocks, self.n_tracks, 2), np.float32)
        self.lib.wbo_process(self.h, n_blocks, out.ctypes.data, peaks.ctypes.data)
        return out, peaks

    def mipmaps(self, sample, quality):
        """-> list of [channels][count] arrays (int16 for quality 1, int8 for 0), one per mip level."""
        dt = np.int16 if quality else np.int8
        cnt = C.c_uint32()
        n = self.lib.wbo_mipmap(self.h, sample, quality, -1, None, 0, C.byref(cnt))
        out = []
        for lv in range(max(n, 0)):
            self.lib.wbo_mipmap(self.h, sample, quality, lv, None, 0, C.byref(cnt))
            buf = np.zeros(cnt.value * self._channels[sample], dt)
            self.lib.wbo_mipmap(self.h, sample, quality, lv, buf.ctypes.data, buf.size, C.byref(cnt))
            out.append(buf.reshape(self._channels[sample], cnt.value))
        return out

    def time_process(self, n_blocks):
        return self.lib.wbo_time_process(self.h, n_blocks)

    def sampler_offset(self, t):
        return self.lib.wbo_sampler_offset(self.h, t)

    def sample_position(self):
        return self.lib.wbo_sample_position(self.h)

    def playhead(self):
        return self.lib.wbo_playhead(self.h)


def panning_coefs(kind, pan):
    l, r = C.c_float(), C.c_float()
    lib(kind).wbo_panning_coefs(pan, C.byref(l), C.byref(r))
    return np.float32(l.value), np.float32(r.value)


def db_to_linear(kind, db):
    return np.float32(lib(kind).wbo_db_to_linear(db))


def interleave(kind, planar, fmt, offset=0, frames=None):
    """planar: [channels][n] f32 -> interleaved bytes in device format `fmt`."""
    planar = np.ascontiguousarray(planar, np.float32)
    ch, n = planar.shape
    frames = n - offset if frames is None else frames
    size = {FMT_I16: 2, FMT_I24: 3, FMT_I24_X8: 4, FMT_I32: 4, FMT_F32: 4}[fmt]
    dst = np.zeros(frames * ch * size, np.uint8)
    ptrs = (C.c_void_p * ch)(*[planar[c].ctypes.data for c in range(ch)])
    lib(kind).wbo_interleave(dst.ctypes.data, ptrs, offset, frames, ch, fmt)
    return dst
