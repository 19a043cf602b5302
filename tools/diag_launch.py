"""GPU diagnostics: CPU enqueue cost vs device time of each C-ABI call, per geometry."""
import sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import hr_pkg
hr = hr_pkg.load()
from hopperrender_b200 import synth

cases = [(1920, 1080, 0)] if len(sys.argv) < 2 else [(1920, 1080, 0), (3840, 2160, 0), (3840, 2160, 1), (7680, 4320, 1)]
for w, h, pf in cases:
    c = synth.MovingTextureClip(w, h, pixfmt=pf)
    g = hr.HrCuda(h, w, w, pf)
    f0, f1 = c.frame(0), c.frame(1)
    g.update_frame(*f0); g.update_frame(*f1)
    print("---- %dx%d %s" % (w, h, "P010" if pf else "NV12"))
    import torch
    dy, duv = torch.from_numpy(f1[0]).cuda(), torch.from_numpy(f1[1]).cuda()
    torch.cuda.synchronize()
    for name, fn in (("calc_flow R=5", lambda: g.calc_flow(5, blocking=False)), ("calc_flow R=16", lambda: g.calc_flow(16, blocking=False)),
                     ("warp mode 2", lambda: g.warp(0.4, 2)), ("pack (device frame)", lambda: g.update_frame_device(dy, duv, borrow=True))):
        for _ in range(20): fn()
        g.synchronize()
        n = 200
        t0 = time.perf_counter()
        for _ in range(n): fn()
        t1 = time.perf_counter()
        g.synchronize()
        t2 = time.perf_counter()
        print("%-20s cpu enqueue %.1f us/call, total %.1f us/call" % (name, (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
    g.close()
