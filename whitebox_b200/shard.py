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
