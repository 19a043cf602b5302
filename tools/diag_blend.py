"""Which float formula does the NVIDIA OpenCL build of the reference warp kernel use for the blend?"""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import hr_pkg
hr = hr_pkg.load()
from hopperrender_b200 import synth
from oracle import ref_opencl
ok, why = ref_opencl.available(); print(ok, why)
w, h = 1920, 1080
c = synth.MovingTextureClip(w, h)
r = ref_opencl.Reference(h, w, w)
r.update_frame(*c.frame(2)); r.update_frame(*c.frame(3)); r.calc_flow(8)
f32 = np.float32
def fma(a, b, cc): return (a.astype(np.float64) * np.float64(b) + cc.astype(np.float64)).astype(np.float32)
for t in (0.4, 0.6, 0.8, 0.2):
    t12 = f32(t); t21 = f32(1.0) - t12
    r.warp(t, 0); a, auv = r.download()
    r.warp(t, 1); b, buv = r.download()
    r.warp(t, 2); m, muv = r.download()
    for nm, A, B, M in (("Y", a, b, m), ("UV", auv, buv, muv)):
        A = A.astype(f32); B = B.astype(f32)
        cands = {
            "unfused": np.trunc(A * t21 + B * t12),
            "fma(a,t21,b*t12)": np.trunc(fma(A, t21, B * t12)),
            "fma(b,t12,a*t21)": np.trunc(fma(B, t12, A * t21)),
        }
        print("t=%.1f %s:" % (t, nm), {k: int((v.astype(np.int64) != M.astype(np.int64)).sum()) for k, v in cands.items()})
# levels: which formula for (v-black)/(white-black)*255 ?
r.warp(0.4, 2, 0.0, 255.0); m0, muv0 = r.download()
r.warp(0.4, 2, 16.0, 219.0); m1, muv1 = r.download()
v = m0.astype(f32)
den = f32(219.0) - f32(16.0)
exact = np.trunc(np.clip((v - f32(16)) / den * f32(255), 0, 255))
recip = np.trunc(np.clip((v - f32(16)) * (f32(1) / den) * f32(255), 0, 255))
print("levels Y: ieee-div mismatches", int((exact != m1).sum()), " reciprocal-mul mismatches", int((recip != m1).sum()))
vu = muv0.astype(f32)
exact = np.trunc(np.clip((vu - f32(128)) / f32(219) * f32(255) + f32(128), 0, 255))
fm = np.trunc(np.clip(fma((vu - f32(128)) / f32(219), f32(255), np.full_like(vu, 128)), 0, 255))
print("levels UV: unfused mismatches", int((exact != muv1).sum()), " fma mismatches", int((fm != muv1).sum()))
