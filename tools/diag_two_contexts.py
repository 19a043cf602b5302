"""Aggregate device-resident throughput of N independent streams (contexts) on ONE GPU, pipelined mode.
python tools/diag_two_contexts.py [NCTX] [RADIUS] [STEPS]   (HR_CUDA_LIB selects a library variant)"""
import sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import hr_pkg
hr = hr_pkg.load()
from hopperrender_b200 import synth, pacing

nctx = int(sys.argv[1]) if len(sys.argv) > 1 else 2
radius = int(sys.argv[2]) if len(sys.argv) > 2 else 5
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 300
w, h, pf = 1920, 1080, 0
fb = int(1.5 * w * h)
nring = 86
c = synth.MovingTextureClip(w, h)
base = [c.frame(k) for k in range(8)]
ring = [(torch.from_numpy(base[k % 8][0]).cuda(), torch.from_numpy(base[k % 8][1]).cuda()) for k in range(nring)]
nout = 48
outs = [(torch.empty((h, w), dtype=torch.uint8, device="cuda"), torch.empty((h // 2, w), dtype=torch.uint8, device="cuda")) for _ in range(nout)]
p = pacing.Pacer(24.0, 60.0)
p.next_source_frame()
ts = [p.next_source_frame() for _ in range(steps + 60)]
ctxs = []
for k in range(nctx):
    st = torch.cuda.Stream()
    g = hr.HrCuda(h, w, w, pf)
    g.set_stream(st.cuda_stream)
    g.set_pipeline(True)
    g.step_device(*ring[nring - 1 - k], [], [])
    ctxs.append((g, st))
oi = 0
CH = 10
def step(i):
    global oi
    if i % CH:
        return 0
    n = 0
    for k, (g, st) in enumerate(ctxs):
        tl = ts[i:i + CH]
        m = sum(len(t) for t in tl)
        o = [outs[(oi + j) % nout] for j in range(m)]
        oi += m
        g.steps_device([ring[(i + j + 7 * k) % nring] for j in range(CH)], tl, o, radius=radius)
        n += m
    return n
for i in range(20):
    step(i)
for g, st in ctxs:
    g.synchronize()
torch.cuda.synchronize()
t0 = time.perf_counter()
n = sum(step(20 + i) for i in range(steps))
t1 = time.perf_counter()
for g, st in ctxs:
    g.synchronize()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("%d contexts R=%d: %.1f us per source frame (aggregate), %.0f frames/s, cpu enqueue %.1f us/frame" % (nctx, radius, (t2 - t0) / (steps * nctx) * 1e6, n / (t2 - t0), (t1 - t0) / (steps * nctx) * 1e6))
for g, st in ctxs:
    g.close()
