"""Per-step SM-clock timeline of the search kernel (developer tool, GPU box only)."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import hr_pkg
hr = hr_pkg.load()
from hopperrender_b200 import synth

R = int(sys.argv[1]) if len(sys.argv) > 1 else 5
w, h = 1920, 1080
c = synth.MovingTextureClip(w, h)
g = hr.HrCuda(h, w, w)
g.update_frame(*c.frame(0)); g.update_frame(*c.frame(1))
for _ in range(5): g.calc_flow(R)
g.set_timeline(True)
g.calc_flow(R)
tl = g.get_timeline()
n = int((tl[0, :120] != 0).sum())
g0, g1 = tl[:, 126], tl[:, 127]
print("globaltimer: CTA start skew %d ns, end skew %d ns, first start -> last end %d ns, median CTA life %d ns" % (g0.max() - g0.min(), g1.max() - g1.min(), g1.max() - g0.min(), np.median(g1 - g0)))
print("stamps per CTA:", n, "ctas", tl.shape[0])
d = np.diff(tl[:, :n], axis=1)
print("total cycles (median over CTAs): %d" % np.median(tl[:, n - 1] - tl[:, 0]))
for i in range(n - 1):
    print("seg %2d: median %6d  min %6d  max %6d" % (i, np.median(d[:, i]), d[:, i].min(), d[:, i].max()))
