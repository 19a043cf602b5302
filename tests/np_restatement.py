"""Independent, vectorised numpy restatement of the reference's flow search + blur.

Written from the prose of SURVEY.md Appendix A (A1, A2), not from oracle/hr_oracle.c: a second
formulation (whole-array gathers, uint64 sums reduced mod 2^32 at the end, add.reduceat window
sums) used only to cross-check the C oracle in the CPU test-suite.
Reference: video/filter/HopperRender/Kernels/calcDeltaSumsKernel.cl:34-189,
determineLowestLayerKernel.cl:2-22, adjustOffsetArrayKernel.cl:2-18, blurFlowKernel.cl:5-12,77-88,
opticalFlowCalc.c:126-203.
"""
import math

import numpy as np


def lattice(H, W):
    s = 0
    while (H >> s) > 270:
        s += 1
    return s, math.ceil(W / 2 ** s), math.ceil(H / 2 ** s)


def first_window(lw, lh):
    m = max(lw, lh)
    if m & (m - 1) == 0:
        ws = m
    else:
        while m & (m - 1):
            m &= m - 1
        ws = m << 1
    return ws // 2


def _mirror(p, D):
    p = np.where(p >= D, 2 * D - p - 1, np.where(p < 0, -p - 1, p))
    return np.clip(p, 0, D - 1)


def _window_sum(v, ws):
    """v: uint64 [lh, lw] -> [nwy, nwx] sums over ws x ws windows aligned at the origin."""
    lh, lw = v.shape
    r = np.add.reduceat(v, np.arange(0, lh, ws), axis=0)
    return np.add.reduceat(r, np.arange(0, lw, ws), axis=1)


def calc_flow(y1, uv1, y2, uv2, R=5, dS=8, nS=6, top8=False, return_layers=False):
    """Frames: (Y [H,W], UV [H/2,W]) arrays; frame1 = previous, frame2 = newest."""
    if top8:
        y1, uv1, y2, uv2 = (a >> 8 for a in (y1, uv1, y2, uv2))
    y1, uv1, y2, uv2 = (a.astype(np.int64) for a in (y1, uv1, y2, uv2))
    H, W = y1.shape
    s, lw, lh = lattice(H, W)
    ws = first_window(lw, lh)
    iters = int(math.log2(ws))
    off = np.zeros((2, lh, lw), np.int64)
    cy, cx = np.meshgrid(np.arange(lh), np.arange(lw), indexing="ij")
    sy, sx = cy << s, cx << s
    Y2 = y2[sy, sx]
    U2 = uv2[sy >> 1, sx & ~1]
    V2 = uv2[sy >> 1, (sx & ~1) + 1]
    layers = []
    for it in range(iters):
        for step in range(2):
            S = []
            for z in range(R):
                rel = z - R // 2
                c = rel * abs(rel)
                ox = off[0] + (c if step == 0 else 0)
                oy = off[1] + (c if step == 1 else 0)
                ox = ((ox + 32768) % 65536) - 32768
                oy = ((oy + 32768) % 65536) - 32768
                nx, ny = _mirror(sx + ox, W), _mirror(sy + oy, H)
                d = np.abs(y1[ny, nx] - Y2) + np.abs(uv1[ny >> 1, nx & ~1] - U2) + np.abs(uv1[ny >> 1, (nx & ~1) + 1] - V2)
                tot = (d.astype(np.uint64) << np.uint64(dS))
                own = ox if step == 0 else oy
                tot = tot + np.abs(own).astype(np.uint64)
                if it >= 4:
                    a = off[step]
                    nb = np.zeros_like(own)
                    for dx, dy in ((0, 2 * ws), (2 * ws, 0), (-2 * ws, 0), (0, -2 * ws)):
                        nb = nb + np.abs(a[np.clip(cy + dy, 0, lh - 1), np.clip(cx + dx, 0, lw - 1)] - own)
                    tot = tot + (nb.astype(np.uint64) << np.uint64(nS))
                S.append(_window_sum(tot, ws) & np.uint64(0xFFFFFFFF))
            S = np.stack(S)                       # [R, nwy, nwx]
            win = np.argmin(S, axis=0)            # first minimum
            layers.append(np.repeat(np.repeat(win, ws, axis=0), ws, axis=1)[:lh, :lw].astype(np.uint8))
            rel = win - R // 2
            upd = rel * np.abs(rel)
            full = np.repeat(np.repeat(upd, ws, axis=0), ws, axis=1)[:lh, :lw]
            off[step] = ((off[step] + full + 32768) % 65536) - 32768
        ws = max(ws >> 1, 1)
    raw = off.astype(np.int16)
    if return_layers:
        return raw, blur(raw), layers
    return raw, blur(raw)


def blur(raw):
    raw = raw.astype(np.int64)
    _, lh, lw = raw.shape
    out = np.zeros_like(raw)
    ys = np.arange(lh)[:, None]
    xs = np.arange(lw)[None, :]
    for ky in range(-4, 4):
        my = np.clip(_blur_mirror(ys + ky, lh), 0, lh - 1)
        for kx in range(-4, 4):
            mx = np.clip(_blur_mirror(xs + kx, lw), 0, lw - 1)
            out += raw[:, my, mx]
    # C division truncates toward zero
    return (np.sign(out) * (np.abs(out) // 64)).astype(np.int16)


def _blur_mirror(p, D):
    return np.where(p >= D, 2 * D - p - 1, np.where(p < 0, -p - 1, p))
