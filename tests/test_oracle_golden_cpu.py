"""Pins the CPU oracle against OUTPUTS OF THE REFERENCE ITSELF, without a GPU.

tests/golden/reference_vectors.npz was produced on the B200 by tests/golden/make_reference_vectors.py:
the reference's unmodified opticalFlowCalc.c and .cl kernels (oracle/_ref) run through the NVIDIA OpenCL
ICD on seeded moving-texture clips. Here the oracle recomputes every case on the same inputs:
raw and blurred offsets and the last step's window sums must be bit-exact; the output planes of all
seven modes must hash to the reference's digests when the oracle uses the arithmetic of that device
(HRO_ARITH_NVCL) — HSV (mode 3) goes through libm's atan2f/fmodf and is held to +-1 on the thumbnails.
"""
import hashlib
import pathlib

import numpy as np
import pytest

GOLD = pathlib.Path(__file__).resolve().parent / "golden" / "reference_vectors.npz"

CASES = [
    ("s1_640x360_r5", 640, 360, 640, 5, 8, 6, (2, 3)),
    ("s1_padded_854x480_r8", 854, 480, 896, 8, 8, 6, (1, 2)),
    ("s0_480x270_r5", 480, 270, 480, 5, 8, 6, (0, 1)),
    ("s2_1280x720_r16", 1280, 720, 1280, 16, 8, 6, (4, 5)),
    ("s1_640x360_r9_scalars", 640, 360, 640, 9, 12, 10, (2, 3)),
]
WARPS = [(0.0, 0.0, 255.0), (0.4, 0.0, 255.0), (0.8, 16.0, 219.0)]


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def gold():
    assert GOLD.exists(), "tests/golden/reference_vectors.npz is missing"
    return np.load(GOLD)


@pytest.mark.parametrize("name,w,h,stride,R,dS,nS,ks", CASES)
def test_oracle_reproduces_the_reference_run(oracle, synth, gold, name, w, h, stride, R, dS, nS, ks):
    assert oracle.have_nvcl()
    c = synth.MovingTextureClip(w, h, stride=stride)
    o = oracle.Oracle(h, stride, w, arith=oracle.ARITH_NVCL)
    o.update_frame(*c.frame(ks[0]))
    o.update_frame(*c.frame(ks[1]))
    o.calc_flow(R, dS, nS)
    raw, blur = o.get_offsets()
    assert np.array_equal(raw, gold[name + "/raw"]), "raw offsets differ from the reference run"
    assert np.array_equal(blur, gold[name + "/blurred"]), "blurred offsets differ from the reference run"
    lh2 = o.lh - (o.lh % 2)
    assert np.array_equal(o.get_last_sums()[:, 0:lh2:2, ::2], gold[name + "/last_sums_w2"]), "last-step window sums differ"
    for mode in range(7):
        for t, black, white in WARPS:
            assert o.warp(t, mode, black, white) == 0
            y, uv = o.download()
            key = "%s/m%d_t%.1f_%g_%g" % (name, mode, t, black, white)
            ty, tuv = gold[key + "/thumb_y"], gold[key + "/thumb_uv"]
            if mode == 3:
                dy = np.abs(y[::15, 0:w:16].astype(int) - ty.astype(int))
                duv = np.abs(uv[::15, 0:w:16].astype(int) - tuv.astype(int))
                assert dy.max() <= 1 and duv.max() <= 1, (key, int(dy.max()), int(duv.max()))
                continue
            want = bytes(gold[key + "/sha"]).decode()
            got = _sha(y[:, :w]) + _sha(uv[:, :w])
            if got != want:   # say where, using the thumbnails
                dy = np.argwhere(y[::15, 0:w:16] != ty)
                pytest.fail("%s: output planes differ from the reference run (thumbnail mismatches: %d)" % (key, len(dy)))
