"""Replay of the filter's pacing arithmetic (host double precision).

video/filter/HopperRender/vf_HopperRender.c:357-375 (blendingScalar advance), :481
(numIntFrames), :490-501 (first source frame produces no warp). SURVEY.md Appendix D.
"""
import math


class Pacer:
    def __init__(self, source_fps=24.0, display_fps=60.0, speed=1.0):
        self.targetFrameTime = 1.0 / display_fps                 # vf_HopperRender.c:682
        self.sourceFrameTime = 1.0 / (source_fps * speed)        # :428
        self.blendingScalar = 0.0                                # :696
        self.sourceFrameNum = 0

    def reset(self):                                             # :562-567
        self.sourceFrameNum = 0
        self.blendingScalar = 0.0

    def next_source_frame(self):
        """Returns the list of float32-bound blending scalars warped for this source frame."""
        self.sourceFrameNum += 1
        ratio = self.targetFrameTime / self.sourceFrameTime
        n = int(max(math.ceil((1.0 - self.blendingScalar) / ratio), 1.0))   # :481
        if self.sourceFrameNum < 2:
            return []                                            # :490-495
        ts = []
        for _ in range(n):
            ts.append(self.blendingScalar)                       # warpFrames(ofc, blendingScalar, mode)  :361
            self.blendingScalar += ratio                         # :371
            if self.blendingScalar >= 1.0:
                self.blendingScalar -= 1.0                       # :372-374
        return ts


def schedule(num_source_frames, source_fps=24.0, display_fps=60.0):
    p = Pacer(source_fps, display_fps)
    return [p.next_source_frame() for _ in range(num_source_frames)]
