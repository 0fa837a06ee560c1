"""Track sharding across the GPUs of one box: tracks are independent until the bus sum
(engine/engine.cpp:1600-1617), so each rank owns a contiguous track range and its samples, mixes its shard
UNCLAMPED, one all-reduce (f32 sum) adds the partial buses, and the clamp (engine.cpp:1627-1636) runs after the
reduce. torch.distributed supplies the communicator (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import numpy as np


def track_range(n_tracks, rank, world):
    """Contiguous shard [lo, hi) of rank `rank`; earlier ranks take the remainder."""
    per, rem = divmod(n_tracks, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def owner_blocks(n_blocks, rank, world):
    """Callbacks [lo, hi) whose bus rank `rank` reduces in the peer-memory exchange (include/wbx.h, "sharded render"):
    callback k belongs to rank k // ceil(n_blocks / world); ranks past the end own nothing."""
    per = (n_blocks + world - 1) // world
    lo = min(rank * per, n_blocks)
    return lo, min(lo + per, n_blocks)


def reduce_owned(partials, rank, world):
    """What an owner computes for its callbacks: the partial buses [K][C][B] of ranks 0..world-1 added in rank order
    (f32, one rounding per add), then the clamp. `partials` is indexed by source rank. numpy statement of
    shard_reduce_kernel, used by the CPU tests of the exchange's host logic."""
    lo, hi = owner_blocks(partials[0].shape[0], rank, world)
    acc = partials[0][lo:hi].astype(np.float32).copy()
    for r in range(1, world):
        acc = (acc + partials[r][lo:hi]).astype(np.float32)
    return clamp_bus(acc)


class _DevPtr:
    """Raw device pointer -> torch tensor without a copy (__cuda_array_interface__)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def bus_tensor(dev):
    """The engine's device bus [C * n_blocks * B] as a torch tensor aliasing the same memory."""
    import torch
    ptr, n = dev.device_bus()
    return torch.as_tensor(_DevPtr(ptr, n), device="cuda")


def mix_sharded(dev, dist, world):
    """After dev.submit(...): mix this rank's shard, reduce the bus across ranks, clamp. All on dev's stream."""
    from . import MIX_NO_CLAMP
    dev.mix(MIX_NO_CLAMP if world > 1 else 0)
    if world > 1:
        ptr, n = dev.device_bus()
        dist.all_reduce(bus_tensor(dev))  # the single exchange step: sum of the partial buses
        dev.clamp_device(ptr, n)


def clamp_bus(x):
    """engine.cpp:1627-1636 on a host array (used by the CPU tests of the sharded path)."""
    return np.where(x > np.float32(1.0), np.float32(1.0), np.where(x < np.float32(-1.0), np.float32(-1.0), x)).astype(np.float32)


def mix_sharded_lockstep(devs):
    """One thread driving all ranks' engines (same process): the three phases of the collective in lock step."""
    for phase in (0, 1, 2):
        for d in devs:
            d.mix_sharded(phase)


class ShardedEngine:
    """One process, several GPUs (or several shards on one GPU): the host engine API of whitebox_b200.Engine in front of
    `len(devices)` engines. Track i lives on shard i % W; a sample is uploaded to a shard the first time one of that
    shard's clips uses it; every render runs the peer-memory bus exchange (include/wbx.h "sharded render") in lock step
    from this thread. Python twin of include/wbx_sharded.hpp, with the scenario API the parity tests drive."""

    def __init__(self, devices, out_channels=2, block=512, rate=48000, bpm=120.0, max_blocks=64, sum_mode=None):
        from . import Engine, SUM_EXACT
        self.C, self.B, self.rate = out_channels, block, rate
        self.W = len(devices)
        self.shards = [Engine(out_channels, block, rate, bpm, device=d, sum_mode=SUM_EXACT if sum_mode is None else sum_mode)
                       for d in devices]
        self.max_blocks = max_blocks
        for r, e in enumerate(self.shards):
            e.dev.shard_init(r, self.W, max_blocks)
        devs = [e.dev for e in self.shards]
        for e in self.shards:
            e.dev.shard_connect_local(devs)
        self.tracks = []      # global track index -> (shard, local index)
        self.solo = []        # Engine::solo_track's UI flag per global track
        self.samples = []     # global sample id -> (data, rate, fmt, {shard: local id})

    def close(self):
        for e in self.shards:
            e.close()

    def add_track(self, volume_db=0.0, pan=0.0, mute=False):
        s = len(self.tracks) % self.W
        self.tracks.append((s, self.shards[s].add_track(volume_db, pan, mute)))
        self.solo.append(False)
        return len(self.tracks) - 1

    def add_sample(self, data, rate, fmt=None):
        from . import FMT_F32
        self.samples.append((data, rate, FMT_F32 if fmt is None else fmt, {}))
        return len(self.samples) - 1

    def _resident(self, sample, shard):
        data, rate, fmt, where = self.samples[sample]
        if shard not in where:
            where[shard] = self.shards[shard].add_sample(data, rate, fmt)
        return where[shard]

    def add_clip(self, track, sample, min_beat, max_beat, start_offset=0.0, speed=1.0, gain=1.0, fade_start=0.0,
                 fade_end=0.0):
        s, t = self.tracks[track]
        return self.shards[s].add_clip(t, self._resident(sample, s), min_beat, max_beat, start_offset, speed, gain,
                                       fade_start, fade_end)

    def _track(self, track):
        s, t = self.tracks[track]
        return self.shards[s], t

    def set_volume(self, t, db):
        e, i = self._track(t)
        e.set_volume(i, db)

    def set_pan(self, t, pan):
        e, i = self._track(t)
        e.set_pan(i, pan)

    def set_mute(self, t, m):
        e, i = self._track(t)
        e.set_mute(i, m)

    def set_effects(self, t, params):
        e, i = self._track(t)
        e.set_effects(i, params)

    def set_clip_gain(self, t, clip, gain):
        e, i = self._track(t)
        e.set_clip_gain(i, clip, gain)

    def set_plugin(self, t, present=True):
        e, i = self._track(t)
        e.set_plugin(i, present)

    def configure(self, out_channels, block, rate):
        """Engine::set_audio_channel_config again: every shard is reconfigured and the exchange (whose buffers are sized by
        block and channel count) is set up anew; tracks, clips, resident samples and the transport persist."""
        for e in self.shards:
            e.dev.shard_close()
            e.configure(out_channels, block, rate)
        self.C, self.B, self.rate = out_channels, block, rate
        for r, e in enumerate(self.shards):
            e.dev.shard_init(r, self.W, self.max_blocks)
        devs = [e.dev for e in self.shards]
        for e in self.shards:
            e.dev.shard_connect_local(devs)

    def delete_track(self, t):
        """Engine::delete_track: the track leaves its shard; later tracks of that shard move down one local slot."""
        s, i = self.tracks.pop(t)
        self.shards[s].delete_track(i)
        self.tracks = [(ss, ii - 1 if (ss == s and ii > i) else ii) for ss, ii in self.tracks]
        self.solo.pop(t)

    def move_track(self, from_slot, to_slot):
        """Engine::move_track: session order only (which slot a track's peaks are reported in); the bus of a sharded
        session is summed shard by shard either way."""
        self.tracks.insert(to_slot, self.tracks.pop(from_slot))
        self.solo.insert(to_slot, self.solo.pop(from_slot))

    def solo_track(self, slot):
        """Engine::solo_track (engine/engine.cpp:245-262) across shards."""
        mute = False
        if self.solo[slot]:
            self.solo[slot] = False
        else:
            self.solo[slot] = True
            self.set_mute(slot, False)
            mute = True
        for i in range(len(self.tracks)):
            if i == slot:
                continue
            self.solo[i] = False
            self.set_mute(i, mute)

    def clip_count(self, t):
        e, i = self._track(t)
        return e.clip_count(i)

    def clip_range(self, t, clip):
        e, i = self._track(t)
        return e.clip_range(i, clip)

    def move_clip(self, t, clip, relative_pos):
        e, i = self._track(t)
        return e.move_clip(i, clip, relative_pos)

    def resize_clip(self, t, clip, *a, **k):
        e, i = self._track(t)
        return e.resize_clip(i, clip, *a, **k)

    def delete_clip(self, t, clip):
        e, i = self._track(t)
        return e.delete_clip(i, clip)

    def duplicate_clip(self, t, clip, min_beat, max_beat):
        e, i = self._track(t)
        return e.duplicate_clip(i, clip, min_beat, max_beat)

    def delete_region(self, t, min_beat, max_beat):
        e, i = self._track(t)
        return e.delete_region(i, min_beat, max_beat)

    def set_impulse_response(self, h):
        for e in self.shards:
            e.set_impulse_response(h)

    def set_resampler(self, mode):
        for e in self.shards:
            e.set_resampler(mode)

    def set_bpm(self, bpm):
        for e in self.shards:
            e.set_bpm(bpm)

    def set_playhead(self, beat):
        for e in self.shards:
            e.set_playhead(beat)

    def play(self):
        for e in self.shards:
            e.play()

    def stop(self):
        for e in self.shards:
            e.stop()

    def sampler_offset(self, t):
        e, i = self._track(t)
        return e.sampler_offset(i)

    def sample_position(self):
        return self.shards[0].sample_position()

    def playhead(self):
        return self.shards[0].playhead()

    def level(self, t, c, reset=False):
        e, i = self._track(t)
        return e.level(i, c, reset)

    def render(self, n_blocks):
        """-> (master bus [C][n_blocks*B] from rank 0, peaks [n_blocks][N][2] re-assembled in session track order)."""
        assert n_blocks <= self.max_blocks
        for e in self.shards:
            e.render_begin(n_blocks)
        mix_sharded_lockstep([e.dev for e in self.shards])
        out = None
        peaks = np.zeros((n_blocks, len(self.tracks), 2), np.float32)
        for r, e in enumerate(self.shards):
            o, p = e.render_end(n_blocks, want_bus=(r == 0))
            if r == 0:
                out = o
            for g, (s, t) in enumerate(self.tracks):
                if s == r:
                    peaks[:, g, :] = p[:, t, :]
        return out, peaks

    def process(self, n_blocks):
        """Scenario API shared with the CPU checkers: -> (out [K][C][B], peaks [K][N][2])."""
        out, peaks = self.render(n_blocks)
        return np.ascontiguousarray(out.reshape(self.C, n_blocks, self.B).transpose(1, 0, 2)), peaks
