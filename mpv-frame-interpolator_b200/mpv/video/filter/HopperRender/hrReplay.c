/*
 * hrReplay.c — the filter's per-frame call sequence as one C loop, for measurements without mpv.
 *
 * For every source frame: updateFrame, calculateOpticalFlow (vf_HopperRender_process_new_source_frame,
 * reference vf_HopperRender.c:412-505), then for every blend scalar the pacing rule asks for (:371-374, :481):
 * warpFrames + downloadFrame (vf_HopperRender_interpolate_frame, :357-375). Only the six functions of
 * opticalFlowCalc.h are called, with host planes, exactly as the filter calls them; bench.py times this loop for
 * its end-to-end figure so that the interpreter is not part of the measurement.
 */
#include "opticalFlowCalc.h"

/* framePlanes: 2 * nFrames pointers (Y, UV of frame 0, Y, UV of frame 1, ...), used round-robin from
 * frame `firstFrame`; tsStart[i] .. tsStart[i + 1] index into ts[] for step i. Returns the number of output
 * frames delivered, or -1 when one of the calls failed. */
long long hrReplayStream(struct OpticalFlowCalc *ofc, unsigned char **framePlanes, int nFrames, int firstFrame, int nSteps, const float *ts,
                         const int *tsStart, int frameOutputMode, unsigned char **outputPlanes) {
    long long delivered = 0;
    for (int i = 0; i < nSteps; ++i) {
        unsigned char **planes = framePlanes + 2 * ((firstFrame + i) % nFrames);
        if (updateFrame(ofc, planes)) return -1;
        if (calculateOpticalFlow(ofc)) return -1;
        for (int k = tsStart[i]; k < tsStart[i + 1]; ++k) {
            if (warpFrames(ofc, ts[k], frameOutputMode)) return -1;
            if (downloadFrame(ofc, outputPlanes)) return -1;
            ++delivered;
        }
    }
    return delivered;
}
