"""numpy statement of the wbx_segment semantics documented in include/wbx.h. TEST INFRASTRUCTURE ONLY.

Lets the CPU test-suite check the product's HOST logic (whitebox_b200 scheduling-only engine -> segment
table) bit-for-bit against the reference's golden vectors without a GPU: segments -> numpy render with the
reference's operation order (dsp/sampler.cpp:34-59,88-210; dsp/dsp_ops.h:27-31; engine/vu_meter.h:20-30;
core/audio_buffer.h:73-82; engine/engine.cpp:1627-1636). Never used by the product or by -m gpu parity."""
import numpy as np

FMT_I16, FMT_I24, FMT_I32, FMT_F32 = 3, 5, 7, 9
f32 = np.float32


def _unity(data, fmt, idx):
    v = data[idx]
    if fmt == FMT_F32:
        return v.astype(f32)
    if fmt == FMT_I16:
        x = v.astype(f32) * (f32(1.0) / f32(32767.0))
        m = np.where(x < f32(1.0), x, f32(1.0))
        return np.where(m > f32(-1.0), m, f32(-1.0)).astype(f32)
    norm = 1.0 / 8388607.0 if fmt == FMT_I24 else 1.0 / 2147483647.0
    x = v.astype(np.float64) * norm
    m = np.where(x < 1.0, x, 1.0)
    return np.where(m > -1.0, m, -1.0).astype(f32)


def _lin(data, fmt, idx):
    v = data[idx]
    if fmt == FMT_F32:
        return v.astype(f32)
    if fmt == FMT_I16:
        return (f32(1.0 / 32767.0) * v.astype(f32)).astype(f32)
    norm = 1.0 / 8388607.0 if fmt == FMT_I24 else 1.0 / 2147483647.0
    return (norm * v.astype(np.float64)).astype(f32)


def _env(n, fin, fout, ln):
    """fade extension (include/wbx.h): envelope of clip-relative output frames n (float64 array)."""
    e = np.ones_like(n)
    if fin > 0.0:
        r = n / fin
        e = np.where(r < 1.0, r, 1.0)
    if fout > 0.0:
        r = (ln - n) / fout
        r = np.where(r > 0.0, r, 0.0)
        r = np.where(r < 1.0, r, 1.0)
        e = e * r
    return e.astype(f32)


def render(segs, gains, samples, C, B, n_blocks, n_tracks, clamp=True):
    """samples: {id: (data[ch][frames], fmt)} -> (out [K][C][B], peaks [K][N][2])"""
    mix = np.zeros((n_blocks, n_tracks, C, B), f32)
    for s in segs:
        data, fmt = samples[int(s["sample_id"])]
        count = data.shape[1]
        pad = np.zeros((data.shape[0], 32), data.dtype)
        data = np.concatenate([data, pad], axis=1)
        pos, speed, length = float(s["src_pos"]), float(s["speed"]), int(s["length"])
        gain = f32(s["gain"])
        for b in range(int(s["n_blocks"])):
            k = int(s["block"]) + b
            if pos >= float(count):
                break
            n_act = min(length, int(np.ceil((float(count) - pos) / speed)))
            j = np.arange(n_act)
            for c in range(C):
                ch = data[c % data.shape[0]]
                if speed == 1.0:
                    v = _unity(ch, fmt, (int(pos) & 0xFFFFFFFF) + j)
                else:
                    x = pos + j.astype(np.float64) * speed
                    ix = x.astype(np.int64)
                    fx = (x - ix.astype(np.float64)).astype(f32)
                    a, bb = _lin(ch, fmt, ix), _lin(ch, fmt, ix + 1)
                    v = (a + fx * (bb - a)).astype(f32)
                d0 = int(s["dst_offset"])
                m = (v * gain).astype(f32)
                if int(s["flags"]) & 1:
                    nn = float(s["clip_frame"]) + float(b) * float(length) + j.astype(np.float64)
                    m = (m * _env(nn, float(s["fade_in_frames"]), float(s["fade_out_frames"]),
                                  float(s["clip_len_frames"]))).astype(f32)
                mix[k, int(s["track"]), c, d0:d0 + n_act] += m
            pos = pos + float(length) * speed
    g = np.asarray(gains, f32).reshape(n_tracks, 2)
    out = np.zeros((n_blocks, C, B), f32)
    peaks = np.zeros((n_blocks, n_tracks, 2), f32)
    for t in range(n_tracks):
        for c in range(C):
            term = (mix[:, t, c, :] * g[t, c]).astype(f32)
            peaks[:, t, c] = np.abs(term).max(axis=1) if B else 0
            out[:, c, :] = (out[:, c, :] + term).astype(f32)
    if clamp:
        out = np.where(out > f32(1.0), f32(1.0), np.where(out < f32(-1.0), f32(-1.0), out)).astype(f32)
    return out, peaks


class ScheduleOnlyEngine:
    """whitebox_b200's host engine without a device (scheduling only) + the numpy render above: exposes the
    scenario API so tests/scenarios.py can drive the product's host logic on a CPU-only box."""

    def __init__(self, C, B, rate, bpm, batched=True, clamp=True):
        import whitebox_b200 as wb
        self.eng = wb.Engine(C, B, rate, bpm, device=-1)
        self.C, self.B, self.batched, self.clamp = C, B, batched, clamp
        self.samples = {}
        self.n_tracks = 0

    def add_track(self, *a):
        self.n_tracks += 1
        return self.eng.add_track(*a)

    def add_sample(self, data, rate, fmt=FMT_F32):
        sid = self.eng.add_sample(data, rate, fmt)
        self.samples[sid] = (np.array(data), fmt)
        return sid

    def delete_track(self, t):
        self.n_tracks -= 1
        return self.eng.delete_track(t)

    def configure(self, C, B, rate):
        self.eng.configure(C, B, rate)
        self.C, self.B = C, B

    def __getattr__(self, name):
        return getattr(self.eng, name)

    def process(self, n_blocks):
        chunks = [n_blocks] if self.batched else [1] * n_blocks
        outs, pks = [], []
        for n in chunks:
            segs, gains = self.eng.schedule(n)
            o, p = render(segs, gains, self.samples, self.C, self.B, n, self.n_tracks, self.clamp)
            outs.append(o)
            pks.append(p)
        return np.concatenate(outs), np.concatenate(pks)
