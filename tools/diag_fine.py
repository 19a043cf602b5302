"""Where the cycles of a warp-local search step go (thread 0 of every CTA, DBG kernel): python tools/diag_fine.py"""
import sys, pathlib
sys.path.insert(0, "/root/repo")
import numpy as np
import hr_pkg
hr = hr_pkg.load()
from hopperrender_b200 import synth
w, h = 1920, 1080
c = synth.MovingTextureClip(w, h)
g = hr.HrCuda(h, w, w)
g.update_frame(*c.frame(0)); g.update_frame(*c.frame(1))
for _ in range(5): g.calc_flow(5)
g.set_timeline(True)
g.calc_flow(5)
tl = g.get_timeline()
# the three warp-local levels (windows 8, 4, 2): slots 40 + level * 12 + axis * 6
for it in (5, 6, 7):
    for ax in (0, 1):
        b = 40 + (it - 5) * 12 + ax * 6
        d = tl[:, b:b + 6]
        dd = np.diff(d, axis=1)
        print("ws %d axis %d: " % (256 >> it, ax) + "  ".join("%s %5d" % (n, np.median(dd[:, k])) for k, n in enumerate(["addr+issue", "neighbours", "loads arrive", "sad+reduce+score", "tail"])) + "   total %d" % np.median(d[:, 5] - d[:, 0]))
