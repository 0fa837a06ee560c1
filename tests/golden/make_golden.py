"""Regenerates tests/golden/*.npz from the reference's OWN Engine::process (oracle/_ref/libwbref.so, built
from /root/reference/src by oracle/Makefile). Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden.py

Each .npz holds the arrays a scenario of tests/scenarios.py returns (clamped bus, per-callback VU peaks,
sampler offsets, transport) bit-for-bit as the reference produced them, plus scalar known answers for the
host-side pan / dB math. The GPU box has no /root/reference; these vectors are how parity travels.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_api as o  # noqa: E402
import scenarios as sc  # noqa: E402


def main():
    assert o.have_ref(), "build oracle/_ref/libwbref.so first (make -C oracle ref)"
    mk = lambda C, B, r, bpm: o.Session("reference", C, B, r, bpm)  # noqa: E731
    for name, fn in sc.ALL.items():
        res = fn(mk)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)
        print(name, {k: v.shape for k, v in res.items()})
    for seed in range(4):
        res = sc.fuzz(mk, seed)
        np.savez_compressed(os.path.join(HERE, "fuzz%d.npz" % seed), **res)
    # waveform mip-maps (gfx/waveform_visual.cpp) of three small samples, both qualities
    mip = {}
    for name, (fmt, frames, ch) in sc.MIP_CASES.items():
        s = o.Session("reference")
        sid = s.add_sample(sc.mip_source(fmt, frames, ch), 48000, fmt)
        for q in (0, 1):
            for lv, a in enumerate(s.mipmaps(sid, q)):
                mip["%s_q%d_l%d" % (name, q, lv)] = a
    np.savez_compressed(os.path.join(HERE, "mipmaps.npz"), **mip)
    pans = np.linspace(-1, 1, 41).astype(np.float32)
    dbs = np.array([0, -6, -12, 6, -72, -71.9, -3.3, 12, -40.5, -100], np.float32)
    pc = np.array([o.panning_coefs("reference", float(p)) for p in pans], np.float32)
    dl = np.array([o.db_to_linear("reference", float(d)) for d in dbs], np.float32)
    rng = np.random.RandomState(8)
    planar = np.concatenate([rng.uniform(-1, 1, (2, 500)).astype(np.float32),
                             np.array([[1, -1, 0, 0.5, -0.5, 1e-9], [-1, 1, -0.0, 0.999999, -0.999999, -1e-9]],
                                      np.float32)], axis=1)
    conv = {("conv_%d" % f): o.interleave("reference", planar, f) for f in (o.FMT_I16, o.FMT_I24, o.FMT_I24_X8, o.FMT_I32, o.FMT_F32)}
    np.savez_compressed(os.path.join(HERE, "scalars.npz"), pans=pans, pan_coefs=pc, dbs=dbs, db_lin=dl,
                        planar=planar, **conv)


if __name__ == "__main__":
    main()
