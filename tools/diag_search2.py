"""Search kernel generations side by side: per-launch device time (CUDA events around the launch, hr_set_profiling) and
equality of the offsets. One subprocess per case with its own timeout, so that a hang costs seconds.
python tools/diag_search2.py [W H PIXFMT] [radii...]"""
import subprocess, sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

CASE = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import hr_pkg
hr = hr_pkg.load()
from hopperrender_b200 import synth
w, h, pf, R, gen, staged = %d, %d, %d, %d, %d, %d
c = synth.MovingTextureClip(w, h, pixfmt=pf)
g = hr.HrCuda(h, w, w, pf)
g.set_search_generation(gen)
g.set_search_staged(bool(staged))
g.set_profiling(True)
g.update_frame(*c.frame(2)); g.update_frame(*c.frame(3))
ts = []
for k in range(30):
    g.calc_flow(R)
    ts.append(g.kernel_times()["search"] * 1e6)
raw, blur = g.get_offsets()
import hashlib
print("gen %%d%%s R %%2d: search %%6.1f us (min %%6.1f)  sha %%s" %% (g.last_search_generation(), " staged  " if g.last_search_staged() else " unstaged", R, float(np.median(ts[5:])), min(ts), hashlib.sha1(raw.tobytes() + blur.tobytes()).hexdigest()[:12]))
'''
w, h, pf = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080, 0)
radii = [int(a) for a in sys.argv[4:]] or [5, 8, 16]
for R in radii:
    for gen, staged in ((1, 0), (2, 0), (2, 1), (3, 0)):
        try:
            r = subprocess.run([sys.executable, "-c", CASE % (str(ROOT), w, h, pf, R, gen, staged)], capture_output=True, text=True, timeout=40)
            print((r.stdout.strip() or r.stderr.strip()[-400:]), flush=True)
        except subprocess.TimeoutExpired:
            print("gen %d staged %d R %2d: TIMEOUT (hang)" % (gen, staged, R), flush=True)
