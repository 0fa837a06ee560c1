"""Development tool: times the waveform mip-map kernels (SURVEY f-4) on one large resident sample; run under
`ncu --metrics gpu__time_duration.sum` to get per-level device times."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whitebox_b200 as wb  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 32 * 1024 * 1024
dev = wb.DeviceEngine(0)
dev.configure(2, 512, 48000)
x = (np.random.default_rng(0).random((2, frames), dtype=np.float32) * 2 - 1).astype(np.float32)
sid = dev.sample_upload(x, 48000)
for q in (1, 0):
    t0 = time.perf_counter()
    levels = dev.sample_mipmaps(sid, q, 2)
    dt = time.perf_counter() - t0
    print("quality %d: %d levels of a %d-frame stereo f32 sample in %.1f ms incl. D2H (%s)" %
          (q, len(levels), frames, dt * 1e3, [a.shape[1] for a in levels[:4]]), flush=True)
