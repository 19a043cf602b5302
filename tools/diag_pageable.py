"""End-to-end rate of the filter's call sequence (C host layer, hrReplay.c) with pageable planes, by number of copying
threads and chunk size of the staging ring (csrc/hr_staging.h), beside pinned planes.
python tools/diag_pageable.py [W H PIXFMT]"""
import os, sys, time, ctypes, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import hr_pkg
hr = hr_pkg.load()
from hopperrender_b200 import synth, pacing

w, h, pf = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080, 0)
npdt = np.uint16 if pf else np.uint8
tdt = torch.uint16 if pf else torch.uint8
clip = synth.MovingTextureClip(w, h, pixfmt=pf)
base = [clip.frame(k) for k in range(6)]
p = pacing.Pacer(24.0, 60.0); p.next_source_frame()
ts = [p.next_source_frame() for _ in range(2000)]
lib = hr.load_ofc_library()


def planes(pinned, src=None):
    if pinned:
        both = torch.zeros((h + h // 2, w), dtype=tdt).pin_memory()
        y, uv = both[:h], both[h:]
        if src is not None:
            y.numpy().view(npdt)[:] = src[0]; uv.numpy().view(npdt)[:] = src[1]
        return (y, uv), both
    both = np.zeros((h + h // 2, w), npdt)
    y, uv = both[:h], both[h:]
    if src is not None:
        y[:] = src[0]; uv[:] = src[1]
    return (y, uv), both


def leg(pinned, threads, chunk):
    os.environ["HR_STAGE_THREADS"] = str(threads)
    os.environ["HR_STAGE_CHUNK_KB"] = str(chunk)
    keep = [planes(pinned, f) for f in base]
    ring = [k[0] for k in keep]
    out, outkeep = planes(pinned)
    c = hr.COpticalFlowCalc()
    c.pixelFormat = pf
    assert not lib.initOpticalFlowCalc(ctypes.byref(c), h, w, w)
    hr.replay_stream_c(c, ring, 5, [[]], 2, out)
    hr.replay_stream_c(c, ring, 0, ts[:20], 2, out)
    n = 300 if w <= 1920 else 80
    t0 = time.perf_counter()
    got = hr.replay_stream_c(c, ring, 2, ts[20:20 + n], 2, out)
    dt = time.perf_counter() - t0
    lib.freeOFC(ctypes.byref(c))
    fb = 1.5 * w * h * (2 if pf else 1)
    return got / dt, (n + got) * fb / dt / 1e9


print("%dx%d pf=%d" % (w, h, pf))
print("pinned planes:                 %8.0f frames/s  %5.1f GB/s over PCIe" % leg(True, 4, 512))
print("pageable, driver's own path:   %8.0f frames/s  %5.1f GB/s" % leg(False, 0, 512))
for threads in (2, 3, 4, 6):
    for chunk in (512, 1024, 2048):
        print("pageable, %d threads, %4d KB: %8.0f frames/s  %5.1f GB/s" % ((threads, chunk) + leg(False, threads, chunk)))
