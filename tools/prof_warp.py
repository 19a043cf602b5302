"""A few warp launches of one geometry for ncu: python tools/prof_warp.py CASE [MODE] [BATCH]   (cases of tools/diag_warp.py)"""
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
w, h, pf = CASES[sys.argv[1]]
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 1
tdt = torch.uint16 if pf else torch.uint8
g = hr.HrCuda(h, w, w, pf)
c = synth.MovingTextureClip(w, h, pixfmt=pf)
f0, f1 = c.frame(0), c.frame(1)
g.update_frame(*f0)
g.update_frame(*f1)
g.calc_flow(5)
outs = [(torch.empty((h, w), dtype=tdt, device="cuda"), torch.empty((h // 2, w), dtype=tdt, device="cuda")) for _ in range(nb)]
for i in range(4):
    g.warp_batch([0.2 + 0.1 * j for j in range(nb)], outs, mode)
g.synchronize()
print("ok")
g.close()
