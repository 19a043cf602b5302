"""Device-resident throughput of one stream, serial vs pipelined, through hr_step_device.
python tools/diag_pipeline.py [W H PIXFMT] [RADIUS] [STEPS]   (HR_CUDA_LIB selects a library variant)"""
import sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import hr_pkg
hr = hr_pkg.load()
from hopperrender_b200 import synth, pacing

w, h, pf = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080, 0)
radius = int(sys.argv[4]) if len(sys.argv) > 4 else 5
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 300
tdt = torch.uint16 if pf else torch.uint8
fb = int(1.5 * w * h * (2 if pf else 1))
nring = min(96, max(8, (2 * 126 * 2**20) // fb + 2))
c = synth.MovingTextureClip(w, h, pixfmt=pf)
base = [c.frame(k) for k in range(8 if fb < (64 << 20) else 4)]
ring = [(torch.from_numpy(base[k % len(base)][0]).cuda().view(tdt), torch.from_numpy(base[k % len(base)][1]).cuda().view(tdt)) for k in range(nring)]
nout = max(8, (126 * 2**20) // fb + 2)
outs = [(torch.empty((h, w), dtype=tdt, device="cuda"), torch.empty((h // 2, w), dtype=tdt, device="cuda")) for _ in range(nout)]
p = pacing.Pacer(24.0, 60.0)
p.next_source_frame()
ts = [p.next_source_frame() for _ in range(steps + 80)]
stream = torch.cuda.Stream()
for pipe in (0, 1, 2):          # 2: pipelined, 25 source frames per call (hr_steps_device)
    g = hr.HrCuda(h, w, w, pf)
    g.set_stream(stream.cuda_stream)
    g.set_pipeline(bool(pipe))
    g.step_device(*ring[nring - 1], [], [])
    oi = 0
    def step(i):
        global oi
        if pipe == 2:
            if i % 25:
                return 0
            tl = ts[i:i + 25]
            n = sum(len(t) for t in tl)
            o = [outs[(oi + j) % nout] for j in range(n)]
            oi += n
            g.steps_device([ring[(i + j) % nring] for j in range(25)], tl, o, radius=radius)
            return n
        o = [outs[(oi + j) % nout] for j in range(len(ts[i]))]
        oi += len(ts[i])
        g.step_device(*ring[i % nring], ts[i], o, radius=radius)
        return len(ts[i])
    with torch.cuda.stream(stream):
        for i in range(25):
            step(i)
        g.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        n = sum(step(25 + i) for i in range(steps))
        t1 = time.perf_counter()
        g.pipeline_join()
        e1.record(stream)
        g.synchronize()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%dx%d pf=%d R=%d pipeline=%d: %.1f us/step device, %.0f frames/s, cpu enqueue %.1f us/step" % (w, h, pf, radius, pipe, ms / steps * 1e3, n / (ms * 1e-3), (t1 - t0) / steps * 1e6))
    g.close()
