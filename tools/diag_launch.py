"""GPU diagnostics: CPU enqueue cost vs device time of each C-ABI call."""
import sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import hr_pkg
hr = hr_pkg.load()
from hopperrender_b200 import synth

w, h = 1920, 1080
c = synth.MovingTextureClip(w, h)
g = hr.HrCuda(h, w, w)
g.update_frame(*c.frame(0)); g.update_frame(*c.frame(1))
for name, fn in (("calc_flow R=5", lambda: g.calc_flow(5, blocking=False)), ("calc_flow R=16", lambda: g.calc_flow(16, blocking=False)), ("warp", lambda: g.warp(0.4, 2))):
    for _ in range(20): fn()
    g.synchronize()
    n = 300
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter()
    g.synchronize()
    t2 = time.perf_counter()
    print("%-16s cpu enqueue %.1f us/call, total %.1f us/call" % (name, (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
g.set_profiling(True)
g.calc_flow(5, blocking=False); g.warp(0.4, 2)
print("event-timed single launches:", {k: round(v * 1e6, 1) for k, v in g.kernel_times().items()})
