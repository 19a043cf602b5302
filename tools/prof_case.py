"""A few steps of one geometry for ncu: python tools/prof_case.py W H PIXFMT [RADIUS] [MODE] [STEPS]"""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import hr_pkg
hr = hr_pkg.load()
from hopperrender_b200 import synth

w, h, pf = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
radius = int(sys.argv[4]) if len(sys.argv) > 4 else 5
mode = int(sys.argv[5]) if len(sys.argv) > 5 else 2
steps = int(sys.argv[6]) if len(sys.argv) > 6 else 4
c = synth.MovingTextureClip(w, h, pixfmt=pf)
g = hr.HrCuda(h, w, w, pf)
tdt = torch.uint16 if pf else torch.uint8
frames = [c.frame(k) for k in range(4)]
dev = [(torch.from_numpy(y).cuda().view(tdt), torch.from_numpy(uv).cuda().view(tdt)) for y, uv in frames]
torch.cuda.synchronize()
g.update_frame_device(*dev[0], borrow=True)
for i in range(steps):
    g.update_frame_device(*dev[(i + 1) % 4], borrow=True)
    g.calc_flow(radius, 8, 6, blocking=False)
    for t in (0.2, 0.6):
        g.warp(t, mode)
g.synchronize()
print("ok")
g.close()
