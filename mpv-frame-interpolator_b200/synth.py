"""Seeded synthetic moving-texture clips: the bench / parity input.

Base luma = uniform uint8 noise (PCG64 seed 1234) box-filtered 8x8 and stretched to [16,235],
periodic; chroma from seeds 1235 / 1236 at half resolution, interleaved as NV12. Frame k is the
texture translated by k*(vx,vy) plus a foreground rectangle (a quarter of the frame, its own
texture) moving by k*(fx,fy). P010 = the 10-bit version of the same picture, << 6.
"""
import numpy as np


def _texture(h, w, seed, lo=16, hi=235):
    rng = np.random.Generator(np.random.PCG64(seed))
    n = rng.integers(0, 256, size=(h, w), dtype=np.uint8).astype(np.float64)
    # periodic 8x8 box filter through cumulative sums
    for axis in (0, 1):
        pad = np.concatenate([n, np.take(n, range(8), axis=axis)], axis=axis)
        c = np.cumsum(pad, axis=axis)
        c = np.concatenate([np.zeros_like(np.take(c, [0], axis=axis)), c], axis=axis)
        n = (np.take(c, range(8, 8 + n.shape[axis]), axis=axis) - np.take(c, range(0, n.shape[axis]), axis=axis)) / 8.0
    mn, mx = n.min(), n.max()
    return (lo + (n - mn) * ((hi - lo) / max(mx - mn, 1e-9)))


class MovingTextureClip:
    def __init__(self, width=1920, height=1080, stride=None, pixfmt=0, velocity=None, fg_velocity=None, bits=None, seed=1234):
        self.w, self.h = width, height
        self.stride = stride or width
        self.pixfmt = pixfmt
        scale = max(1, round(height / 1080)) if height >= 1080 else 1
        self.v = velocity if velocity is not None else (12 * scale, 4 * scale)
        self.fv = fg_velocity if fg_velocity is not None else (-8 * scale, 6 * scale)
        self.bits = bits or (10 if pixfmt == 1 else 8)
        top = float((1 << self.bits) - 1) / 255.0
        self.dtype = np.uint16 if pixfmt == 1 else np.uint8
        self.shift = 6 if pixfmt == 1 else 0
        cw, ch = width // 2, height // 2
        self.bgY = np.rint(_texture(height, width, seed) * top).astype(np.uint16)
        self.bgU = np.rint(_texture(ch, cw, seed + 1, 64, 192) * top).astype(np.uint16)
        self.bgV = np.rint(_texture(ch, cw, seed + 2, 64, 192) * top).astype(np.uint16)
        fh, fw = (height // 4) * 2, (width // 4) * 2
        self.fgY = np.rint(_texture(fh, fw, seed + 3) * top).astype(np.uint16)
        self.fgU = np.rint(_texture(fh // 2, fw // 2, seed + 4, 64, 192) * top).astype(np.uint16)
        self.fgV = np.rint(_texture(fh // 2, fw // 2, seed + 5, 64, 192) * top).astype(np.uint16)
        self.fh, self.fw = fh, fw

    def frame(self, k):
        """Returns (Y [h, stride], UV [h/2, stride]) of source frame k."""
        vx, vy = self.v
        y = np.roll(self.bgY, (k * vy, k * vx), axis=(0, 1)).copy()
        u = np.roll(self.bgU, ((k * vy) // 2, (k * vx) // 2), axis=(0, 1)).copy()
        v = np.roll(self.bgV, ((k * vy) // 2, (k * vx) // 2), axis=(0, 1)).copy()
        # foreground rectangle, top-left corner kept even so chroma stays aligned
        fx0 = (self.w // 4 + k * self.fv[0]) % max(self.w - self.fw, 2)
        fy0 = (self.h // 4 + k * self.fv[1]) % max(self.h - self.fh, 2)
        fx0 &= ~1
        fy0 &= ~1
        y[fy0:fy0 + self.fh, fx0:fx0 + self.fw] = self.fgY
        u[fy0 // 2:fy0 // 2 + self.fh // 2, fx0 // 2:fx0 // 2 + self.fw // 2] = self.fgU
        v[fy0 // 2:fy0 // 2 + self.fh // 2, fx0 // 2:fx0 // 2 + self.fw // 2] = self.fgV
        Y = np.zeros((self.h, self.stride), self.dtype)
        UV = np.zeros((self.h // 2, self.stride), self.dtype)
        Y[:, :self.w] = (y << self.shift).astype(self.dtype)
        UV[:, 0:self.w:2] = (u << self.shift).astype(self.dtype)
        UV[:, 1:self.w:2] = (v << self.shift).astype(self.dtype)
        if self.stride > self.w:  # padding columns are part of the search lattice; fill deterministically
            Y[:, self.w:] = Y[:, self.w - 1:self.w]
            UV[:, self.w:] = np.tile(UV[:, self.w - 2:self.w], (1, (self.stride - self.w + 1) // 2))[:, :self.stride - self.w]
        return Y, UV


def noise_frame(height, stride, seed, pixfmt=0):
    """Full-range white noise (adversarial: uint32 window-sum wrap at large windows)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if pixfmt == 1:
        y = (rng.integers(0, 1024, size=(height, stride), dtype=np.uint16) << 6).astype(np.uint16)
        uv = (rng.integers(0, 1024, size=(height // 2, stride), dtype=np.uint16) << 6).astype(np.uint16)
    else:
        y = rng.integers(0, 256, size=(height, stride), dtype=np.uint8)
        uv = rng.integers(0, 256, size=(height // 2, stride), dtype=np.uint8)
    return y, uv
