"""Search launches only, back to back, pipelined (two lanes) vs serial: what two co-resident launches cost one another.
python tools/diag_search_only.py [RADIUS]  (HR_SEARCH_GEN forces a generation)"""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import hr_pkg
hr = hr_pkg.load()
from hopperrender_b200 import synth
R = int(sys.argv[1]) if len(sys.argv) > 1 else 5
w, h = 1920, 1080
c = synth.MovingTextureClip(w, h)
stream = torch.cuda.Stream()
for gen in (1, 3):
    for pipe in (0, 1):
        g = hr.HrCuda(h, w, w, 0)
        g.set_stream(stream.cuda_stream)
        g.set_search_generation(gen)
        g.update_frame(*c.frame(0)); g.update_frame(*c.frame(1))
        g.set_pipeline(bool(pipe))
        with torch.cuda.stream(stream):
            for _ in range(50): g.calc_flow(R, blocking=False)
            g.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            n = 2000
            for _ in range(n): g.calc_flow(R, blocking=False)
            if pipe: g.pipeline_join()
            e1.record(stream)
            g.synchronize(); torch.cuda.synchronize()
        print("generation %d R %d %s: %.1f us per search" % (gen, R, "two lanes " if pipe else "one stream", e0.elapsed_time(e1) * 1e3 / n))
        g.set_pipeline(False)
        g.close()
