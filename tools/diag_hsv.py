"""Mode 3 (HSV flow) differences: reference OpenCL vs oracle (NVCL) vs CUDA, per flow vector (GPU box only)."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import hr_pkg
hr = hr_pkg.load()
from hopperrender_b200 import synth
from oracle import ref_opencl, hr_oracle_py as O
ok, why = ref_opencl.available(); print(ok, why)
w, h = 1920, 1080
c = synth.MovingTextureClip(w, h)
r = ref_opencl.Reference(h, w, w); o = O.Oracle(h, w, w); g = hr.HrCuda(h, w, w)
for x in (r, o, g):
    x.update_frame(*c.frame(2)); x.update_frame(*c.frame(3)); x.calc_flow(8)
blur = r.get_offsets()[1]
r.warp(0.0, 3); ry, ruv = r.download()
o.warp(0.0, 3); oy, ouv = o.download()
g.warp(0.0, 3); gy, guv, _ = g.download()
s = 2
for nm, a in (("oracle", (oy, ouv)), ("cuda", (gy, guv))):
    for pl, (A, B) in enumerate(((a[0], ry), (a[1], ruv))):
        d = A.astype(int) - B.astype(int)
        print(nm, "plane", pl, "differ %.3f%%" % (100 * (d != 0).mean()), "hist", {int(k): int(v) for k, v in zip(*np.unique(d, return_counts=True))})
# per-vector table from the chroma plane (U,V depend on the vector only)
seen = {}
H2 = h // 2
for cy in range(0, H2, 4):
    for cx in range(0, w, 8):
        ly, lx = (cy >> s) << 1, (cx >> s) & ~1
        v = (int(blur[0, ly, lx]), int(blur[1, ly, lx]))
        if v in seen: continue
        seen[v] = (tuple(int(t) for t in ruv[cy, cx:cx + 2]), tuple(int(t) for t in ouv[cy, cx:cx + 2]), tuple(int(t) for t in guv[cy, cx:cx + 2]))
bad = {k: v for k, v in seen.items() if v[0] != v[1] or v[0] != v[2]}
print("distinct vectors", len(seen), "with differing UV", len(bad))
for k in list(bad)[:25]:
    print("flow", k, "ref UV", bad[k][0], "oracle", bad[k][1], "cuda", bad[k][2])
