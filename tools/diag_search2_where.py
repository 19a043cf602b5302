"""Step timeline of a generation-2 search launch (clock stamps of thread 0 of every CTA); with --host the stamps live in
mapped host memory and are read without waiting: where is every warp of a launch that does not return standing?
python tools/diag_search2_where.py [--host] [W H PIXFMT] [RADIUS]"""
import os, sys, time, pathlib
HOST = "--host" in sys.argv   # stamps in mapped host memory: readable while the launch hangs, but every stamp crosses PCIe
if HOST:
    sys.argv.remove("--host")
    os.environ["HR_TIMELINE_HOST"] = "1"
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import hr_pkg
hr = hr_pkg.load()
from hopperrender_b200 import synth
w, h, pf = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080, 0)
R = int(sys.argv[4]) if len(sys.argv) > 4 else 5
c = synth.MovingTextureClip(w, h, pixfmt=pf)
g = hr.HrCuda(h, w, w, pf)
GEN = int(os.environ.get('GEN', '2'))
g.set_search_generation(GEN)
g.set_search_staged(os.environ.get('STAGED', '1') == '1')
g.set_timeline(True)
g.update_frame(*c.frame(2)); g.update_frame(*c.frame(3))
for _ in range(3):
    g.calc_flow(R, blocking=False)
if HOST:
    time.sleep(2.0)
    t = g.peek_timeline()
else:
    t = g.get_timeline()
prog = t[:, 104:120]
done = t[:, 101] != 0
print("CTAs done: %d of %d" % (int(done.sum()), len(done)))
if not done.all():
    vals, counts = np.unique(prog[~done], return_counts=True)
    for v, n in zip(vals, counts):
        print("  step %2d phase %d: %d warps" % (v // 8, v % 8, n))
    tx = g.info.lowWidth // 32 + (1 if g.info.lowWidth % 32 else 0)
    for cta in np.nonzero(~done)[0][:12]:
        print("  cta %3d (tile %d,%d): %s" % (cta, cta % tx, cta // tx, " ".join("%d.%d" % (v // 8, v % 8) for v in prog[cta])))
    os._exit(1)
seg = np.diff(t[:, 1:1 + 4 * 2 * g.info.iterations + 1].astype(np.int64), axis=1)
med = np.median(seg, axis=0)
tot = np.median(t[:, 101] - t[:, 0])
print("median CTA: %d cycles launch entered -> blur done; search done at %d" % (tot, np.median(t[:, 100] - t[:, 0])))
for st in range(2 * g.info.iterations):
    a = med[4 * st:4 * st + 4]
    print("step %2d (window %3d, %s): loads issued %5d  totals %5d  winner %5d  tail %5d   = %5d" % (st, g.info.firstWindow >> (st // 2), "xy"[st & 1], a[0], a[1], a[2], a[3] if len(a) > 3 else 0, a.sum()))
print("blur: %d" % np.median(t[:, 101] - t[:, 100]))
print("wall (globaltimer) first start -> last end: %.1f us" % ((t[:, 127].max() - t[:, 126].min()) / 1e3))
