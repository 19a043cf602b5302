"""Host-visible cost of each call of the reference-facing interface (pinned planes), per geometry."""
import sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import hr_pkg
hr = hr_pkg.load()
from hopperrender_b200 import synth, pacing

cases = [(1920, 1080, 0), (3840, 2160, 1)] if len(sys.argv) < 2 else [(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))]
for w, h, pf in cases:
    tdt = torch.uint16 if pf else torch.uint8
    c = synth.MovingTextureClip(w, h, pixfmt=pf)
    fr = [c.frame(k) for k in range(4)]
    hin = [(torch.from_numpy(y).pin_memory(), torch.from_numpy(uv).pin_memory()) for y, uv in fr]
    hout = (torch.empty((h, w), dtype=tdt).pin_memory(), torch.empty((h // 2, w), dtype=tdt).pin_memory())
    ofc = hr.OpticalFlowCalc()
    assert not hr.initOpticalFlowCalc(ofc, h, w, w, pf)
    p = pacing.Pacer(24.0, 60.0); p.next_source_frame()
    ts = [p.next_source_frame() for _ in range(140)]
    acc = {"update": 0.0, "flow": 0.0, "warp": 0.0, "download": 0.0}
    cnt = {"update": 0, "flow": 0, "warp": 0, "download": 0}
    hr.updateFrame(ofc, list(hin[0]))
    for i in range(120):
        t0 = time.perf_counter(); hr.updateFrame(ofc, list(hin[(i + 1) % 4])); t1 = time.perf_counter()
        hr.calculateOpticalFlow(ofc); t2 = time.perf_counter()
        if i >= 20:
            acc["update"] += t1 - t0; acc["flow"] += t2 - t1; cnt["update"] += 1; cnt["flow"] += 1
        for t in ts[i]:
            t3 = time.perf_counter(); hr.warpFrames(ofc, t, 2); t4 = time.perf_counter()
            hr.downloadFrame(ofc, list(hout)); t5 = time.perf_counter()
            if i >= 20:
                acc["warp"] += t4 - t3; acc["download"] += t5 - t4; cnt["warp"] += 1; cnt["download"] += 1
    fb = 1.5 * w * h * (2 if pf else 1)
    print("%dx%d pf=%d: " % (w, h, pf) + ", ".join("%s %.1f us" % (k, acc[k] / cnt[k] * 1e6) for k in acc) +
          "; PCIe-only time of one frame at 55 GB/s: %.1f us; device flow %.1f us, warp+download %.1f us" % (fb / 55e3, ofc.ofcCalcTime * 1e6, ofc.warpCalcTime * 1e6))
    hr.freeOFC(ofc)
