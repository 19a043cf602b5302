"""End-to-end rate of the filter's call sequence through the compiled C host layer (hrReplay.c), pinned planes, a few
repetitions: python tools/diag_e2e_c.py [W H PIXFMT] [REPS]"""
import sys, time, ctypes, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import hr_pkg
hr = hr_pkg.load()
from hopperrender_b200 import synth, pacing

w, h, pf = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080, 0)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
npdt = np.uint16 if pf else np.uint8
tdt = torch.uint16 if pf else torch.uint8
clip = synth.MovingTextureClip(w, h, pixfmt=pf)
base = [clip.frame(k) for k in range(6)]
p = pacing.Pacer(24.0, 60.0); p.next_source_frame()
ts = [p.next_source_frame() for _ in range(2000)]
lib = hr.load_ofc_library()
keep = []
ring = []
for f in base:
    both = torch.zeros((h + h // 2, w), dtype=tdt).pin_memory()
    both[:h].numpy().view(npdt)[:] = f[0]; both[h:].numpy().view(npdt)[:] = f[1]
    keep.append(both); ring.append((both[:h], both[h:]))
ob = torch.zeros((h + h // 2, w), dtype=tdt).pin_memory()
out = (ob[:h], ob[h:])
c = hr.COpticalFlowCalc()
c.pixelFormat = pf
assert not lib.initOpticalFlowCalc(ctypes.byref(c), h, w, w)
hr.replay_stream_c(c, ring, 5, [[]], 2, out)
hr.replay_stream_c(c, ring, 0, ts[:20], 2, out)
n = 600 if w <= 1920 else 100
rates = []
for r in range(reps):
    t0 = time.perf_counter()
    got = hr.replay_stream_c(c, ring, 2, ts[20:20 + n], 2, out)
    dt = time.perf_counter() - t0
    rates.append(got / dt)
lib.freeOFC(ctypes.byref(c))
print("%dx%d pf=%d: %s frames/s (median %.0f), %.1f us per source frame" % (w, h, pf, " ".join("%.0f" % r for r in rates), sorted(rates)[len(rates) // 2], 1e6 * 2.5 / sorted(rates)[len(rates) // 2]))
