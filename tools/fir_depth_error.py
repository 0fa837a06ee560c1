#!/usr/bin/env python
"""Error of the tensor-core convolution reverb at full accumulation depth (cfg 5 shape: 65536 taps, 128 signals, 81920
frames) against the f64 specification, per third of the render. python tools/fir_depth_error.py [taps] [tracks] [blocks]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenarios as sc  # noqa: E402
import whitebox_b200 as wb  # noqa: E402

taps = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
tracks = int(sys.argv[2]) if len(sys.argv) > 2 else 64
blocks = int(sys.argv[3]) if len(sys.argv) > 3 else 160
os.environ.setdefault("WBX_FIR", "tc")
res = sc.reverb_full_depth(lambda C, B, r, bpm: wb.Engine(C, B, r, bpm, device=0, sum_mode=wb.SUM_EXACT), wb.effect_params,
                           taps=taps, n_tracks=tracks, n_blocks=blocks)
want, want_peaks = sc.reverb_f64_expected(res, wb.panning_coefs, wb.db_to_linear)
peak = np.abs(want).max(axis=(1, 2), keepdims=True)
err = (np.abs(res["out"].astype(np.float64) - want) / peak).max(axis=(1, 2))
third = max(1, blocks // 3)
print("taps %d, %d signals, %d callbacks: max error of block peak %.3g (first third %.3g, middle %.3g, last %.3g); split factor %d" %
      (taps, 2 * tracks, blocks, err.max(), err[:third].max(), err[third:2 * third].max(), err[2 * third:].max(),
       wb.lib().wbx_fir_split_factor()))
pk = float(np.abs(want_peaks).max())
print("VU peaks: max error %.3g of the largest peak" % (float(np.abs(res["peaks"].astype(np.float64) - want_peaks).max()) / pk))
