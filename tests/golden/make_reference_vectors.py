#!/usr/bin/env python
"""Generates tests/golden/reference_vectors.npz ON THE GPU BOX from the REFERENCE ITSELF: the unmodified
opticalFlowCalc.c + .cl kernels of /root/reference (built into oracle/_ref by oracle/build_ref.py)
executed on the B200 through the NVIDIA OpenCL ICD (oracle/ref_opencl.py).

For a few small seeded cases (sizes the CPU oracle finishes in seconds) it stores the reference's raw and
blurred offsets, the window sums of the last search step, and SHA-256 digests + thumbnails of the
output planes for every output mode. tests/test_oracle_golden_cpu.py then pins the oracle against these
without a GPU.

  gpurun -- 'python tests/golden/make_reference_vectors.py'  ->  gpurun_out/reference_vectors.npz (copy here)
"""
import hashlib
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import hr_pkg

hr_pkg.load()
from hopperrender_b200 import synth
from oracle import ref_opencl

CASES = [
    # name, w, h, stride, radius, deltaScalar, neighborBiasScalar, frames (k1, k2)
    ("s1_640x360_r5", 640, 360, 640, 5, 8, 6, (2, 3)),
    ("s1_padded_854x480_r8", 854, 480, 896, 8, 8, 6, (1, 2)),
    ("s0_480x270_r5", 480, 270, 480, 5, 8, 6, (0, 1)),
    ("s2_1280x720_r16", 1280, 720, 1280, 16, 8, 6, (4, 5)),
    ("s1_640x360_r9_scalars", 640, 360, 640, 9, 12, 10, (2, 3)),
]
WARPS = [(0.0, 0.0, 255.0), (0.4, 0.0, 255.0), (0.8, 16.0, 219.0)]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ok, why = ref_opencl.available()
    if not ok:
        raise SystemExit("reference OpenCL run unavailable: " + why)
    out = {}
    for name, w, h, stride, R, dS, nS, (k1, k2) in CASES:
        c = synth.MovingTextureClip(w, h, stride=stride)
        r = ref_opencl.Reference(h, stride, w)
        r.update_frame(*c.frame(k1))
        r.update_frame(*c.frame(k2))
        r.calc_flow(R, dS, nS)
        raw, blur = r.get_offsets()
        out[name + "/raw"] = raw
        out[name + "/blurred"] = blur
        sums = r.get_last_sums(R)
        lh2 = r.lh - (r.lh % 2)
        out[name + "/last_sums_w2"] = sums[:, 0:lh2:2, ::2].copy()
        for mode in range(7):
            for t, black, white in WARPS:
                assert not r.warp(t, mode, black, white)
                y, uv = r.download()
                key = "%s/m%d_t%.1f_%g_%g" % (name, mode, t, black, white)
                out[key + "/sha"] = np.frombuffer((sha(y[:, :w]) + sha(uv[:, :w])).encode(), np.uint8)
                out[key + "/thumb_y"] = y[::15, 0:w:16].copy()
                out[key + "/thumb_uv"] = uv[::15, 0:w:16].copy()
        r.close()
        print(name, "done")
    d = ROOT / "gpurun_out"
    d.mkdir(exist_ok=True)
    np.savez_compressed(d / "reference_vectors.npz", **out)
    print("wrote", d / "reference_vectors.npz")


if __name__ == "__main__":
    main()
