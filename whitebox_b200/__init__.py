"""whitebox_b200 — B200-native mixing hot path of native-m/whitebox.

Python is only a thin ctypes binding over the C ABI in include/wbx.h (device engine) and include/wbx_host.h
(host engine with the reference's editing/transport API). All sample work runs in hand-written sm_100a CUDA
kernels inside whitebox_b200/libwbx.so; there is no CPU render path and importing this package fails loudly
if the library has not been built (python whitebox_b200/build.py, or __graft_entry__.build()).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwbx.so")

FMT_I16, FMT_I24, FMT_I24_X8, FMT_I32, FMT_F32 = 3, 5, 6, 7, 9
SUM_AUTO, SUM_EXACT, SUM_TREE = 0, 1, 2
MIX_NO_CLAMP = 1
_NP = {FMT_I16: np.int16, FMT_I24: np.int32, FMT_I32: np.int32, FMT_F32: np.float32}


class Segment(C.Structure):
    """wbx_segment (include/wbx.h)."""
    _fields_ = [("track", C.c_uint32), ("block", C.c_uint32), ("n_blocks", C.c_uint32), ("dst_offset", C.c_uint32),
                ("length", C.c_uint32), ("sample_id", C.c_uint32), ("src_pos", C.c_double), ("speed", C.c_double),
                ("gain", C.c_float), ("flags", C.c_uint32), ("clip_frame", C.c_double), ("fade_in_frames", C.c_double),
                ("fade_out_frames", C.c_double), ("clip_len_frames", C.c_double)]


class EffectParams(C.Structure):
    """wbx_effect_params (include/wbx.h) — same layout as the oracle's wbo_effects."""
    _fields_ = [("eq_freq", C.c_float * 4), ("eq_gain_db", C.c_float * 4), ("eq_q", C.c_float * 4),
                ("comp_threshold_db", C.c_float), ("comp_attack_ms", C.c_float), ("comp_release_ms", C.c_float),
                ("comp_makeup_db", C.c_float), ("comp_ratio_code", C.c_int32), ("reverb_on", C.c_int32)]


def effect_params(eq=((100.0, 0.0, 0.7), (500.0, 0.0, 1.0), (3000.0, 0.0, 1.0), (9000.0, 0.0, 0.7)),
                  threshold_db=0.0, ratio_code=0, attack_ms=5.0, release_ms=80.0, makeup_db=0.0, reverb=False):
    p = EffectParams()
    for b, (f, g, q) in enumerate(eq):
        p.eq_freq[b], p.eq_gain_db[b], p.eq_q[b] = f, g, q
    p.comp_threshold_db, p.comp_ratio_code = threshold_db, ratio_code
    p.comp_attack_ms, p.comp_release_ms, p.comp_makeup_db = attack_ms, release_ms, makeup_db
    p.reverb_on = int(reverb)
    return p


SEGMENT_DTYPE = np.dtype([("track", "<u4"), ("block", "<u4"), ("n_blocks", "<u4"), ("dst_offset", "<u4"),
                          ("length", "<u4"), ("sample_id", "<u4"), ("src_pos", "<f8"), ("speed", "<f8"),
                          ("gain", "<f4"), ("flags", "<u4"), ("clip_frame", "<f8"), ("fade_in_frames", "<f8"),
                          ("fade_out_frames", "<f8"), ("clip_len_frames", "<f8")])
assert SEGMENT_DTYPE.itemsize == C.sizeof(Segment) == 80
SEG_FADE, SEG_POLYPHASE = 1, 2
IPC_HANDLE_BYTES = 64

# every symbol include/wbx.h and include/wbx_host.h declare
WBX_SYMBOLS = [
    "wbx_abi_version", "wbx_create", "wbx_destroy", "wbx_last_error", "wbx_configure", "wbx_set_track_count",
    "wbx_set_sum_mode", "wbx_set_stream", "wbx_sample_upload", "wbx_sample_release", "wbx_sample_update", "wbx_sample_mipmap", "wbx_render", "wbx_submit",
    "wbx_mix", "wbx_fetch", "wbx_fetch_levels", "wbx_host_alloc", "wbx_host_free", "wbx_fetch_interleaved", "wbx_device_bus", "wbx_device_peaks", "wbx_clamp_device",
    "wbx_synchronize", "wbx_launch_count", "wbx_fir_split_factor", "wbx_fir_path", "wbx_last_kernel", "wbx_effects_design", "wbx_set_track_effects",
    "wbx_set_impulse_response", "wbx_render_levels", "wbx_bounce_begin", "wbx_bounce_push", "wbx_bounce_pop", "wbx_shard_init", "wbx_shard_connect_ipc",
    "wbx_shard_connect_local", "wbx_mix_sharded", "wbx_mix_sharded_phase", "wbx_shard_reset", "wbx_shard_set_host_output", "wbx_host_register",
    "wbx_host_unregister", "wbx_shard_close", "wbx_shard_info",
]
WBXH_SYMBOLS = [
    "wbxh_create", "wbxh_destroy", "wbxh_last_error", "wbxh_device", "wbxh_add_track", "wbxh_set_volume",
    "wbxh_set_pan", "wbxh_set_mute", "wbxh_add_sample", "wbxh_add_clip", "wbxh_add_clip_fade", "wbxh_set_playhead",
    "wbxh_play", "wbxh_set_effects", "wbxh_set_impulse_response", "wbxh_set_resampler",
    "wbxh_stop", "wbxh_set_fast_forward", "wbxh_render", "wbxh_schedule", "wbxh_sampler_offset",
    "wbxh_sample_position", "wbxh_playhead", "wbxh_level", "wbxh_panning_coefs", "wbxh_db_to_linear",
    "wbxh_advance_rounded", "wbxh_render_begin", "wbxh_render_end", "wbxh_clip_count", "wbxh_clip_range", "wbxh_set_bpm", "wbxh_delete_track", "wbxh_move_track", "wbxh_solo_track",
    "wbxh_set_clip_gain", "wbxh_move_clip",
    "wbxh_resize_clip", "wbxh_delete_clip", "wbxh_duplicate_clip", "wbxh_delete_region", "wbxh_set_plugin", "wbxh_configure",
    "wbxh_cpu_usage", "wbxh_bounce", "wbxh_bounce_wav",
]

_lib = None


class WbxError(RuntimeError):
    pass


def lib():
    """Loads libwbx.so (never falls back to anything else)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WbxError("whitebox_b200/libwbx.so is not built: run `python whitebox_b200/build.py` "
                       "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, dbl, flt, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_double, C.c_float, C.c_int
    pp = C.POINTER(vp)
    L.wbx_abi_version.restype = i32
    L.wbx_create.argtypes = [pp, i32]
    L.wbx_destroy.argtypes = [vp]
    L.wbx_last_error.argtypes = [vp]
    L.wbx_last_error.restype = C.c_char_p
    L.wbx_configure.argtypes = [vp, u32, u32, u32]
    L.wbx_set_track_count.argtypes = [vp, u32]
    L.wbx_set_sum_mode.argtypes = [vp, i32]
    L.wbx_set_stream.argtypes = [vp, vp]
    L.wbx_sample_upload.argtypes = [vp, i32, u32, u64, u32, pp, C.POINTER(u32)]
    L.wbx_sample_release.argtypes = [vp, u32]
    L.wbx_sample_update.argtypes = [vp, u32, pp]
    L.wbx_sample_mipmap.argtypes = [vp, u32, i32, i32, vp, u64, C.POINTER(u32)]
    L.wbx_render.argtypes = [vp, vp, u32, vp, u32, pp, vp]
    L.wbx_render_levels.argtypes = [vp, vp, u32, vp, u32, pp, vp, vp]
    L.wbx_shard_init.argtypes = [vp, u32, u32, u32, vp]
    L.wbx_shard_connect_ipc.argtypes = [vp, vp]
    L.wbx_shard_connect_local.argtypes = [vp, pp]
    L.wbx_mix_sharded.argtypes = [vp]
    L.wbx_mix_sharded_phase.argtypes = [vp, i32]
    L.wbx_shard_close.argtypes = [vp]
    L.wbx_shard_reset.argtypes = [vp]
    L.wbx_shard_set_host_output.argtypes = [vp, pp, u64]
    L.wbx_host_register.argtypes = [vp, C.c_size_t]
    L.wbx_host_unregister.argtypes = [vp]
    L.wbx_shard_info.argtypes = [vp, C.POINTER(u32), C.POINTER(u32)]
    L.wbx_submit.argtypes = [vp, vp, u32, vp, u32]
    L.wbx_mix.argtypes = [vp, u32]
    L.wbx_fetch.argtypes = [vp, pp, vp]
    L.wbx_fetch_interleaved.argtypes = [vp, vp, i32]
    L.wbx_fetch_levels.argtypes = [vp, vp]
    L.wbx_host_alloc.argtypes = [C.c_size_t]
    L.wbx_host_alloc.restype = vp
    L.wbx_host_free.argtypes = [vp]
    L.wbx_host_free.restype = None
    L.wbx_device_bus.argtypes = [vp, pp, C.POINTER(u64)]
    L.wbx_device_peaks.argtypes = [vp, pp, C.POINTER(u64)]
    L.wbx_clamp_device.argtypes = [vp, vp, u64]
    L.wbx_synchronize.argtypes = [vp]
    L.wbx_effects_design.argtypes = [vp, u32, vp]
    L.wbx_set_track_effects.argtypes = [vp, u32, vp]
    L.wbx_launch_count.argtypes = [vp]
    L.wbx_launch_count.restype = u64
    L.wbx_last_kernel.argtypes = [vp]
    L.wbx_last_kernel.restype = C.c_char_p
    L.wbxh_create.argtypes = [pp, i32, u32, u32, u32, dbl]
    L.wbxh_destroy.argtypes = [vp]
    L.wbxh_destroy.restype = None
    L.wbxh_last_error.argtypes = [vp]
    L.wbxh_last_error.restype = C.c_char_p
    L.wbxh_device.argtypes = [vp]
    L.wbxh_device.restype = vp
    L.wbxh_add_track.argtypes = [vp, flt, flt, i32]
    for f in ("wbxh_set_volume", "wbxh_set_pan"):
        getattr(L, f).argtypes = [vp, i32, flt]
    L.wbxh_set_mute.argtypes = [vp, i32, i32]
    L.wbxh_set_plugin.argtypes = [vp, i32, i32]
    L.wbxh_configure.argtypes = [vp, u32, u32, u32]
    L.wbxh_cpu_usage.argtypes = [vp]
    L.wbxh_cpu_usage.restype = dbl
    L.wbxh_bounce.argtypes = [vp, dbl, dbl, i32, u32, vp, u64, C.POINTER(u64)]
    L.wbxh_bounce_wav.argtypes = [vp, dbl, dbl, i32, u32, C.c_char_p, C.POINTER(u64)]
    L.wbx_bounce_begin.argtypes = [vp, i32]
    L.wbx_bounce_push.argtypes = [vp]
    L.wbx_bounce_pop.argtypes = [vp, pp, C.POINTER(C.c_size_t)]
    L.wbxh_add_sample.argtypes = [vp, i32, u32, u64, u32, pp]
    L.wbxh_add_clip.argtypes = [vp, i32, i32, dbl, dbl, dbl, dbl, flt]
    L.wbxh_add_clip_fade.argtypes = [vp, i32, i32, dbl, dbl, dbl, dbl, flt, dbl, dbl]
    L.wbxh_delete_track.argtypes = [vp, i32]
    L.wbxh_move_track.argtypes = [vp, i32, i32]
    L.wbxh_solo_track.argtypes = [vp, i32]
    L.wbxh_set_clip_gain.argtypes = [vp, i32, i32, flt]
    L.wbxh_clip_count.argtypes = [vp, i32]
    L.wbxh_clip_range.argtypes = [vp, i32, i32, C.POINTER(dbl), C.POINTER(dbl)]
    L.wbxh_move_clip.argtypes = [vp, i32, i32, dbl]
    L.wbxh_resize_clip.argtypes = [vp, i32, i32, dbl, dbl, dbl, i32, i32, i32]
    L.wbxh_delete_clip.argtypes = [vp, i32, i32]
    L.wbxh_duplicate_clip.argtypes = [vp, i32, i32, dbl, dbl]
    L.wbxh_delete_region.argtypes = [vp, i32, dbl, dbl]
    L.wbxh_set_effects.argtypes = [vp, i32, vp]
    L.wbxh_set_resampler.argtypes = [vp, i32]
    L.wbxh_set_resampler.restype = None
    L.wbxh_set_impulse_response.argtypes = [vp, vp, u32]
    L.wbx_set_impulse_response.argtypes = [vp, vp, u32]
    L.wbxh_set_bpm.argtypes = [vp, dbl]
    L.wbxh_set_bpm.restype = None
    L.wbxh_set_playhead.argtypes = [vp, dbl]
    L.wbxh_set_playhead.restype = None
    for f in ("wbxh_play", "wbxh_stop"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = None
    L.wbxh_set_fast_forward.argtypes = [vp, i32]
    L.wbxh_set_fast_forward.restype = None
    L.wbxh_render.argtypes = [vp, u32, pp, vp]
    L.wbxh_render_begin.argtypes = [vp, u32]
    L.wbxh_render_end.argtypes = [vp, pp, vp]
    L.wbxh_schedule.argtypes = [vp, u32, pp, C.POINTER(u32), pp]
    L.wbxh_sampler_offset.argtypes = [vp, i32]
    L.wbxh_sampler_offset.restype = dbl
    for f in ("wbxh_sample_position", "wbxh_playhead"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = dbl
    L.wbxh_level.argtypes = [vp, i32, i32, i32]
    L.wbxh_level.restype = flt
    L.wbxh_panning_coefs.argtypes = [flt, C.POINTER(flt), C.POINTER(flt)]
    L.wbxh_panning_coefs.restype = None
    L.wbxh_advance_rounded.argtypes = [C.POINTER(dbl), dbl, u32, dbl]
    L.wbxh_advance_rounded.restype = u32
    L.wbxh_db_to_linear.argtypes = [flt]
    L.wbxh_db_to_linear.restype = flt
    _lib = L
    return L


def _chan_ptrs(arr):
    """arr: [channels][n] contiguous -> (void*[channels])"""
    return (C.c_void_p * arr.shape[0])(*[arr[c].ctypes.data for c in range(arr.shape[0])])


class PinnedArray:
    """float32 numpy array in page-locked host memory (wbx_host_alloc): D2H copies land in it directly."""

    def __init__(self, shape):
        self.n = int(np.prod(shape))
        self.ptr = lib().wbx_host_alloc(self.n * 4)
        if not self.ptr:
            raise WbxError("wbx_host_alloc failed")
        self.array = np.frombuffer((C.c_float * self.n).from_address(self.ptr), dtype=np.float32).reshape(shape)

    def __del__(self):
        try:
            if self.ptr:
                lib().wbx_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def panning_coefs(pan):
    l, r = C.c_float(), C.c_float()
    lib().wbxh_panning_coefs(pan, C.byref(l), C.byref(r))
    return np.float32(l.value), np.float32(r.value)


def db_to_linear(db):
    return np.float32(lib().wbxh_db_to_linear(db))


class DeviceEngine:
    """The device engine of include/wbx.h: resident samples + segment table -> mixed bus."""

    def __init__(self, device=0, handle=None):
        self.L = lib()
        self._own = handle is None
        if handle is None:
            h = C.c_void_p()
            rc = self.L.wbx_create(C.byref(h), device)
            if rc != 0:
                raise WbxError("wbx_create(device=%d) failed with status %d: no sm_100 CUDA device "
                               "(there is no CPU fallback)" % (device, rc))
            handle = h
        self.h = handle
        self.C = self.B = self.n_tracks = 0
        self.n_blocks = 0

    def close(self):
        if self.h and self._own:
            self.L.wbx_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise WbxError("wbx status %d: %s" % (rc, (self.L.wbx_last_error(self.h) or b"").decode()))

    def configure(self, out_channels, block, rate):
        self._ck(self.L.wbx_configure(self.h, out_channels, block, rate))
        self.C, self.B = out_channels, block

    def set_track_count(self, n):
        self._ck(self.L.wbx_set_track_count(self.h, n))
        self.n_tracks = n

    def set_sum_mode(self, mode):
        self._ck(self.L.wbx_set_sum_mode(self.h, mode))

    def set_stream(self, cuda_stream):
        self._ck(self.L.wbx_set_stream(self.h, C.c_void_p(cuda_stream)))

    def sample_upload(self, data, rate, fmt=FMT_F32):
        data = np.ascontiguousarray(data, dtype=_NP[fmt])
        sid = C.c_uint32()
        self._ck(self.L.wbx_sample_upload(self.h, fmt, data.shape[0], data.shape[1], rate, _chan_ptrs(data),
                                          C.byref(sid)))
        return sid.value

    def sample_upload_planar(self, channels, rate, fmt=FMT_F32):
        ptrs = (C.c_void_p * len(channels))(*[c.ctypes.data for c in channels])
        sid = C.c_uint32()
        self._ck(self.L.wbx_sample_upload(self.h, fmt, len(channels), channels[0].size, rate, ptrs, C.byref(sid)))
        return sid.value

    def sample_mipmaps(self, sid, quality, channels):
        """-> list of [channels][count] arrays (int16 for quality 1, int8 for 0), one per mip level."""
        dt = np.int16 if quality else np.int8
        cnt = C.c_uint32()
        n = self.L.wbx_sample_mipmap(self.h, sid, quality, -1, None, 0, C.byref(cnt))
        if n < 0:
            self._ck(n)
        out = []
        for lv in range(n):
            self.L.wbx_sample_mipmap(self.h, sid, quality, lv, None, 0, C.byref(cnt))
            buf = np.zeros(cnt.value * channels, dt)
            r = self.L.wbx_sample_mipmap(self.h, sid, quality, lv, buf.ctypes.data, buf.size, C.byref(cnt))
            if r < 0:
                self._ck(r)
            out.append(buf.reshape(channels, cnt.value))
        return out

    def sample_update_planar(self, sid, channels):
        ptrs = (C.c_void_p * len(channels))(*[c.ctypes.data for c in channels])
        self._ck(self.L.wbx_sample_update(self.h, sid, ptrs))

    def sample_release(self, sid):
        self._ck(self.L.wbx_sample_release(self.h, sid))

    def submit(self, segs, gains, n_blocks):
        segs = np.ascontiguousarray(segs, dtype=SEGMENT_DTYPE)
        gains = np.ascontiguousarray(gains, dtype=np.float32)
        self._ck(self.L.wbx_submit(self.h, segs.ctypes.data, len(segs), gains.ctypes.data, n_blocks))
        self.n_blocks = n_blocks

    def submit_raw(self, segs_ptr, n_segs, gains_ptr, n_blocks):
        self._ck(self.L.wbx_submit(self.h, segs_ptr, n_segs, gains_ptr, n_blocks))
        self.n_blocks = n_blocks

    def mix(self, flags=0):
        self._ck(self.L.wbx_mix(self.h, flags))

    def fetch(self, want_peaks=True, want_bus=True):
        out = np.empty((self.C, self.n_blocks * self.B), np.float32) if want_bus else None
        peaks = np.empty((self.n_blocks, self.n_tracks, 2), np.float32) if want_peaks else None
        self._ck(self.L.wbx_fetch(self.h, _chan_ptrs(out) if want_bus else None,
                                  peaks.ctypes.data if want_peaks else None))
        return out, peaks

    # ---- sharded render (include/wbx.h "sharded render"): the bus exchange over peer memory ----------------
    def shard_init(self, rank, world, max_blocks):
        """-> this rank's CUDA IPC handle (64 bytes) for the other ranks' shard_connect_ipc."""
        h = C.create_string_buffer(IPC_HANDLE_BYTES)
        self._ck(self.L.wbx_shard_init(self.h, rank, world, max_blocks, h))
        return h.raw

    def shard_connect_ipc(self, handles):
        """handles: list of `world` 64-byte handles, indexed by rank (one process per GPU)."""
        blob = b"".join(handles)
        assert len(blob) == IPC_HANDLE_BYTES * len(handles)
        self._ck(self.L.wbx_shard_connect_ipc(self.h, blob))

    def shard_connect_local(self, engines):
        """engines: the `world` DeviceEngines of this process, indexed by rank (all after shard_init)."""
        arr = (C.c_void_p * len(engines))(*[getattr(en.h, "value", en.h) for en in engines])
        self._ck(self.L.wbx_shard_connect_local(self.h, arr))

    def mix_sharded(self, phase=None):
        """The collective sharded mix; phase 0/1/2 = its three stages (one thread driving several engines)."""
        self._ck(self.L.wbx_mix_sharded(self.h) if phase is None else self.L.wbx_mix_sharded_phase(self.h, phase))

    def shard_set_host_output(self, out):
        """out: [C][frames] f32 array in page-locked memory every rank maps (PinnedArray within one process, a registered
        shared-memory segment across processes), or None to clear."""
        if out is None:
            self._ck(self.L.wbx_shard_set_host_output(self.h, None, 0))
        else:
            self._ck(self.L.wbx_shard_set_host_output(self.h, _chan_ptrs(out), out.shape[1]))

    def shard_close(self):
        self._ck(self.L.wbx_shard_close(self.h))

    def shard_reset(self):
        """Recovery after a failed sharded mix: call on EVERY rank, then barrier on the host, then render again."""
        self._ck(self.L.wbx_shard_reset(self.h))

    def shard_info(self):
        r, w = C.c_uint32(), C.c_uint32()
        self._ck(self.L.wbx_shard_info(self.h, C.byref(r), C.byref(w)))
        return r.value, w.value

    def fetch_levels(self):
        lv = np.zeros((self.n_tracks, 2), np.float32)
        self._ck(self.L.wbx_fetch_levels(self.h, lv.ctypes.data))
        return lv

    def fetch_interleaved(self, fmt):
        frames = self.n_blocks * self.B
        size = {FMT_I16: 2, FMT_I24: 3, FMT_I24_X8: 4, FMT_I32: 4, FMT_F32: 4}[fmt]
        n = frames * 3 if fmt == FMT_I24 else frames * self.C * size
        dst = np.zeros(n, np.uint8)
        self._ck(self.L.wbx_fetch_interleaved(self.h, dst.ctypes.data, fmt))
        return dst

    def render(self, segs, gains, n_blocks, want_peaks=True):
        self.submit(segs, gains, n_blocks)
        self.mix()
        return self.fetch(want_peaks)

    def device_bus(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self.L.wbx_device_bus(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def device_peaks(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self.L.wbx_device_peaks(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def clamp_device(self, ptr, n):
        self._ck(self.L.wbx_clamp_device(self.h, C.c_void_p(ptr), n))

    def synchronize(self):
        self._ck(self.L.wbx_synchronize(self.h))

    def launch_count(self):
        return self.L.wbx_launch_count(self.h)

    def last_kernel(self):
        return (self.L.wbx_last_kernel(self.h) or b"").decode()


class Engine:
    """Host engine (include/wbx_host.h): the reference's editing / transport API in front of the CUDA mix.

    batched=True renders process(n_blocks) as ONE device launch over n_blocks callbacks (offline bounce /
    throughput mode); batched=False issues one launch per callback like the realtime audio thread.
    """

    def __init__(self, out_channels=2, block=512, rate=48000, bpm=120.0, device=0, batched=True,
                 sum_mode=SUM_AUTO):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.wbxh_create(C.byref(h), device, out_channels, block, rate, bpm)
        if rc != 0:
            raise WbxError("wbxh_create failed with status %d: no sm_100 CUDA device or bad config "
                           "(there is no CPU fallback)" % rc)
        self.h = h
        self.C, self.B, self.rate = out_channels, block, rate
        self._bpm = bpm
        self.n_tracks = 0
        self.batched = batched
        self.dev = None
        if device >= 0:  # device < 0: scheduling-only engine (host logic tests); render() then fails
            self.dev = DeviceEngine(handle=C.c_void_p(self.L.wbxh_device(self.h)))
            self.dev.C, self.dev.B = out_channels, block
            if sum_mode != SUM_AUTO:
                self.dev.set_sum_mode(sum_mode)

    def close(self):
        if self.h:
            self.L.wbxh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise WbxError("wbx status %d: %s" % (rc, (self.L.wbxh_last_error(self.h) or b"").decode()))
        return rc

    def add_track(self, volume_db=0.0, pan=0.0, mute=False):
        self.n_tracks += 1
        return self.L.wbxh_add_track(self.h, volume_db, pan, int(mute))

    def set_volume(self, t, db):
        self._ck(self.L.wbxh_set_volume(self.h, t, db))

    def set_pan(self, t, pan):
        self._ck(self.L.wbxh_set_pan(self.h, t, pan))

    def set_mute(self, t, m):
        self._ck(self.L.wbxh_set_mute(self.h, t, int(m)))

    def set_plugin(self, t, present=True):
        """A plugin in the track's slot: the reference then drops the track's clip audio (engine/track.cpp:600,645-724)."""
        self._ck(self.L.wbxh_set_plugin(self.h, t, int(present)))

    def configure(self, out_channels, block, rate):
        """Engine::set_audio_channel_config again (device change): tracks, clips, samples and transport persist."""
        self._ck(self.L.wbxh_configure(self.h, out_channels, block, rate))
        self.C, self.B, self.rate = out_channels, block, rate
        if self.dev is not None:
            self.dev.C, self.dev.B = out_channels, block

    def cpu_usage(self):
        return self.L.wbxh_cpu_usage(self.h)

    def bounce(self, start_beat, end_beat, fmt=FMT_F32, chunk_blocks=256, path=None, out=None):
        """Offline export of [start_beat, end_beat): -> interleaved bytes in device format `fmt` (np.uint8 array), or the
        frame count when `path` names a WAV file to write. `out`: optional preallocated np.uint8 destination."""
        frames = C.c_uint64()
        if path is not None:
            self._ck(self.L.wbxh_bounce_wav(self.h, start_beat, end_beat, fmt, chunk_blocks, str(path).encode(), C.byref(frames)))
            return frames.value
        size = {FMT_I16: 2, FMT_I24_X8: 4, FMT_I32: 4, FMT_F32: 4}[fmt]
        if out is None:
            bpm_frames = int(np.ceil((end_beat - start_beat) * 60.0 / self._bpm * self.rate)) + self.B
            out = np.zeros(bpm_frames * self.C * size, np.uint8)
        self._ck(self.L.wbxh_bounce(self.h, start_beat, end_beat, fmt, chunk_blocks, out.ctypes.data, out.nbytes, C.byref(frames)))
        return out[:frames.value * self.C * size]

    def add_sample(self, data, rate, fmt=FMT_F32):
        data = np.ascontiguousarray(data, dtype=_NP[fmt])
        return self._ck(self.L.wbxh_add_sample(self.h, fmt, data.shape[0], data.shape[1], rate, _chan_ptrs(data)))

    def add_sample_planar(self, channels, rate, fmt=FMT_F32):
        """channels: list of 1-D contiguous arrays (one per channel, equal length) - no host-side copy."""
        ptrs = (C.c_void_p * len(channels))(*[c.ctypes.data for c in channels])
        assert all(c.flags["C_CONTIGUOUS"] and c.dtype == _NP[fmt] and c.size == channels[0].size for c in channels)
        return self._ck(self.L.wbxh_add_sample(self.h, fmt, len(channels), channels[0].size, rate, ptrs))

    def add_clip(self, track, sample, min_beat, max_beat, start_offset=0.0, speed=1.0, gain=1.0, fade_start=0.0,
                 fade_end=0.0):
        if fade_start or fade_end:
            return self._ck(self.L.wbxh_add_clip_fade(self.h, track, sample, min_beat, max_beat, start_offset, speed,
                                                      gain, fade_start, fade_end))
        return self._ck(self.L.wbxh_add_clip(self.h, track, sample, min_beat, max_beat, start_offset, speed, gain))

    # clip editing (clip = index in the track's clip list ordered by min_beat): Engine::move_clip / resize_clip /
    # delete_clip / duplicate_clip of the reference
    def delete_track(self, track):
        self._ck(self.L.wbxh_delete_track(self.h, track))
        self.n_tracks -= 1

    def move_track(self, from_slot, to_slot):
        self._ck(self.L.wbxh_move_track(self.h, from_slot, to_slot))

    def solo_track(self, track):
        self._ck(self.L.wbxh_solo_track(self.h, track))

    def set_clip_gain(self, track, clip, gain):
        self._ck(self.L.wbxh_set_clip_gain(self.h, track, clip, gain))

    def clip_count(self, track):
        return self._ck(self.L.wbxh_clip_count(self.h, track))

    def clip_range(self, track, clip):
        a, b = C.c_double(), C.c_double()
        self._ck(self.L.wbxh_clip_range(self.h, track, clip, C.byref(a), C.byref(b)))
        return a.value, b.value

    def move_clip(self, track, clip, relative_pos):
        return self._ck(self.L.wbxh_move_clip(self.h, track, clip, relative_pos))

    def resize_clip(self, track, clip, relative_pos, resize_limit, min_length, left_side, shift=False, stretch=False):
        return self._ck(self.L.wbxh_resize_clip(self.h, track, clip, relative_pos, resize_limit, min_length, int(left_side),
                                                int(shift), int(stretch)))

    def delete_clip(self, track, clip):
        return self._ck(self.L.wbxh_delete_clip(self.h, track, clip))

    def duplicate_clip(self, track, clip, min_beat, max_beat):
        return self._ck(self.L.wbxh_duplicate_clip(self.h, track, clip, min_beat, max_beat))

    def delete_region(self, track, min_beat, max_beat):
        return self._ck(self.L.wbxh_delete_region(self.h, track, min_beat, max_beat))

    def set_effects(self, track, params):
        """params: EffectParams (see effect_params()) or None to remove the chain."""
        self._ck(self.L.wbxh_set_effects(self.h, track, C.byref(params) if params is not None else None))

    def set_impulse_response(self, h):
        """Convolution-reverb impulse response (f32 array) shared by every chain with reverb=True; None removes it."""
        if h is None:
            return self._ck(self.L.wbxh_set_impulse_response(self.h, None, 0))
        h = np.ascontiguousarray(h, np.float32)
        return self._ck(self.L.wbxh_set_impulse_response(self.h, h.ctypes.data, h.size))

    def set_resampler(self, mode):
        """0 = linear (the reference's only resampler), 1 = polyphase windowed sinc (extension)."""
        self.L.wbxh_set_resampler(self.h, mode)

    def set_bpm(self, bpm):
        self._bpm = bpm
        self.L.wbxh_set_bpm(self.h, bpm)

    def set_playhead(self, beat):
        self.L.wbxh_set_playhead(self.h, beat)

    def play(self):
        self.L.wbxh_play(self.h)

    def stop(self):
        self.L.wbxh_stop(self.h)

    def render(self, n_blocks, want_peaks=True, out=None, want_bus=True):
        """-> (bus [C][n_blocks*B], peaks [n_blocks][N][2]) from one device launch. `out` may be a
        caller-owned [C][n_blocks*B] f32 array (e.g. PinnedArray(...).array: the kernel then writes it directly).
        want_bus=False: ranks > 0 of a sharded setup, which do not receive the master bus."""
        if out is None and want_bus:
            out = np.empty((self.C, n_blocks * self.B), np.float32)
        peaks = np.zeros((n_blocks, self.n_tracks, 2), np.float32) if want_peaks else None
        self._ck(self.L.wbxh_render(self.h, n_blocks, _chan_ptrs(out) if want_bus else None,
                                    peaks.ctypes.data if (peaks is not None and self.n_tracks) else None))
        if self.dev is not None:  # keep the device view's shape in step (fetch / fetch_interleaved after render)
            self.dev.C, self.dev.B, self.dev.n_tracks, self.dev.n_blocks = self.C, self.B, self.n_tracks, n_blocks
        return out, peaks

    def render_begin(self, n_blocks):
        """First half of render() for lock-step sharded rendering (whitebox_b200.shard.ShardedEngine)."""
        self._ck(self.L.wbxh_render_begin(self.h, n_blocks))
        self.dev.C, self.dev.B, self.dev.n_tracks, self.dev.n_blocks = self.C, self.B, self.n_tracks, n_blocks

    def render_end(self, n_blocks, want_bus=True, want_peaks=True):
        out = np.empty((self.C, n_blocks * self.B), np.float32) if want_bus else None
        peaks = np.zeros((n_blocks, self.n_tracks, 2), np.float32) if want_peaks else None
        self._ck(self.L.wbxh_render_end(self.h, _chan_ptrs(out) if want_bus else None,
                                        peaks.ctypes.data if (peaks is not None and self.n_tracks) else None))
        return out, peaks

    def process(self, n_blocks):
        """Scenario API shared with the CPU checkers: -> (out [K][C][B], peaks [K][N][2])."""
        if self.batched:
            out, peaks = self.render(n_blocks)
            return np.ascontiguousarray(out.reshape(self.C, n_blocks, self.B).transpose(1, 0, 2)), peaks
        outs, pks = [], []
        for _ in range(n_blocks):
            o, p = self.render(1)
            outs.append(o.reshape(1, self.C, self.B))
            pks.append(p)
        return np.concatenate(outs, axis=0), np.concatenate(pks, axis=0)

    def schedule(self, n_blocks):
        """Host scheduling only -> (segments structured array copy, gains [N][2] copy)."""
        segs, n, gains = C.c_void_p(), C.c_uint32(), C.c_void_p()
        self._ck(self.L.wbxh_schedule(self.h, n_blocks, C.byref(segs), C.byref(n), C.byref(gains)))
        if n.value:
            buf = (C.c_char * (n.value * SEGMENT_DTYPE.itemsize)).from_address(segs.value)
            s = np.frombuffer(buf, dtype=SEGMENT_DTYPE).copy()
        else:
            s = np.zeros(0, SEGMENT_DTYPE)
        if self.n_tracks:
            gb = (C.c_float * (self.n_tracks * 2)).from_address(gains.value)
            g = np.frombuffer(gb, dtype=np.float32).reshape(self.n_tracks, 2).copy()
        else:
            g = np.zeros((0, 2), np.float32)
        return s, g

    def sampler_offset(self, t):
        return self.L.wbxh_sampler_offset(self.h, t)

    def sample_position(self):
        return self.L.wbxh_sample_position(self.h)

    def playhead(self):
        return self.L.wbxh_playhead(self.h)

    def level(self, t, c, reset=False):
        return self.L.wbxh_level(self.h, t, c, int(reset))
