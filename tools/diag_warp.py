"""Device time of the warp kernel per geometry / mode, HBM-cold: every timed launch is preceded by a write of a
buffer larger than L2 (the flush alone is timed too and subtracted).

    python tools/diag_warp.py [case ...]        case = 1080p-nv12 | 4k-nv12 | 4k-p010 | 8k-p010 | 8k-nv12 | 1080p-p010
    HR_CUDA_LIB=tools/variants/libhr_X.so python tools/diag_warp.py ...   (a kernel variant)
Prints one line per (case, mode, batch): us per output frame, GB/s of algorithmic bytes, fraction of the measured peak.
"""
import json
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

import hr_pkg

hr = hr_pkg.load()
from hopperrender_b200 import synth

CASES = {"1080p-nv12": (1920, 1080, 0), "1080p-p010": (1920, 1080, 1), "4k-nv12": (3840, 2160, 0), "4k-p010": (3840, 2160, 1),
         "8k-nv12": (7680, 4320, 0), "8k-p010": (7680, 4320, 1)}
names = [a for a in sys.argv[1:] if a in CASES] or ["1080p-nv12", "4k-p010", "8k-p010"]
modes = [int(a[5:]) for a in sys.argv[1:] if a.startswith("mode=")] or [2]
batches = [int(a[6:]) for a in sys.argv[1:] if a.startswith("batch=")] or [1]
peak = 6552.6
try:
    peak = json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass

stream = torch.cuda.Stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name in names:
    w, h, pf = CASES[name]
    bps = 2 if pf else 1
    tdt = torch.uint16 if pf else torch.uint8
    c = synth.MovingTextureClip(w, h, pixfmt=pf)
    g = hr.HrCuda(h, w, w, pf)
    g.set_stream(stream.cuda_stream)
    if c is None:      # 8K: a cheap pair (noise shifted by a few pixels), the warp's time does not depend on the picture
        f0 = synth.noise_frame(h, w, 3, pf)
        f1 = (np.roll(f0[0], (16, 48), (0, 1)), np.roll(f0[1], (8, 48), (0, 1)))
    else:
        f0, f1 = c.frame(0), c.frame(1)
    g.update_frame(*f0)
    g.update_frame(*f1)
    g.calc_flow(5)
    lw, lh = g.info.lowWidth, g.info.lowHeight
    alg = 3 * int(1.5 * w * h * bps) + 4 * lw * lh
    nout = 8
    outs = [(torch.empty((h, w), dtype=tdt, device="cuda"), torch.empty((h // 2, w), dtype=tdt, device="cuda")) for _ in range(nout)]
    for mode in modes:
        for nb in batches:
            ts = [0.2 + 0.1 * i for i in range(nb)]

            def run(with_warp, reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    e0.record(stream)
                    for i in range(reps):
                        flush.fill_(i & 255)
                        if with_warp:
                            g.warp_batch(ts, outs[:nb], mode)
                    e1.record(stream)
                stream.synchronize()
                return e0.elapsed_time(e1) * 1e3 / reps

            run(True, 5)
            reps = 30
            t_both = min(run(True, reps) for _ in range(3))
            t_flush = min(run(False, reps) for _ in range(3))
            us = (t_both - t_flush) / nb
            gbs = alg / us * 1e-3
            print("%-11s mode %d batch %d: %7.2f us per output  %7.1f GB/s  %.3f of %.0f GB/s   (flush %.1f us)" % (name, mode, nb, us, gbs, gbs / peak, peak, t_flush), flush=True)
    g.close()
