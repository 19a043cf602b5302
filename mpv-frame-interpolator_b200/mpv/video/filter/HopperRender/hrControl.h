/*
 * hrControl.h — headless control surface for the filter.
 *
 * The reference steers the filter at run time through a GTK tray applet that writes integer codes into a pipe
 * (video/filter/HopperRender/vf_HopperRender.c:112-180, HopperRenderSettingsApplet.py); with INC_APP_IND 0 nothing
 * is left to steer it with, and with it the filter blocks at start-up when GTK is missing (README.md:51). This
 * module keeps the applet's protocol — the same codes with the same meaning — and drops the applet: codes arrive as
 * text lines from any file descriptor (a FIFO, a socket) or from mpv's `vf-command` string, and are applied to the
 * same state. Two codes are added for reproducible runs: a fixed search radius and its release.
 *
 *   0 / 1        interpolation off (counters restart) / on                      reference :127-136
 *   2 .. 8       frame output mode = code - 2 (enum FrameOutput)                :138-158
 *   9, 10, 11    level presets 0/255, 10/219, 16/219                            :159-170
 *   100 .. 355   black level = code - 100                                       :173-174
 *   400 .. 655   white level = code - 400                                       :175-176
 *   700 .. 731   deltaScalar = code - 700                                       :177-178
 *   800 .. 831   neighborBiasScalar = code - 800                                :179-180
 *   900 .. 932   (added) search radius pinned to code - 900, 900 = follow the timing again
 */
#ifndef HRCONTROL_H
#define HRCONTROL_H

#include <stddef.h>

#include "opticalFlowCalc.h"

/* the part of the filter's private state that the codes touch (struct priv, reference :29-72) */
typedef struct HrControlState {
    int interpolationActive; /* interpolationState: Active (1) / Deactivated (0)                      */
    int frameOutputMode;     /* enum FrameOutput, 0 .. 6                                               */
    int restartCounters;     /* set when code 0 arrives: sourceFrameNum, interpolatedFrameNum, blendingScalar := 0 */
    int pinnedRadius;        /* 0: the filter's timing rule moves the radius; otherwise keep this one  */
    /* hrControlPoll: a line cut in two by the end of a read is kept here until its rest arrives        */
    int pendingLength;
    char pending[28];
} HrControlState;

/* first decimal integer of a text line, as the applet channel reads it (a line that does not start with a digit
 * carries no code); -1 when there is none */
int hrControlParse(const char *text);
/* apply one code; returns 0 when the code is known, 1 otherwise (state untouched) */
int hrControlApply(struct OpticalFlowCalc *ofc, HrControlState *st, int code);
/* read whatever is waiting on fd (non-blocking descriptors welcome), apply every complete line; an unfinished last
 * line waits in the state for the next call (at end of file it is applied as it is); returns the number of codes
 * applied, -1 on a read error other than "nothing there" */
int hrControlPoll(int fd, struct OpticalFlowCalc *ofc, HrControlState *st);
/* the status text the applet's widget shows (reference :188-207), for whoever wants to display it; returns the length */
int hrControlStatus(char *buf, size_t size, const struct OpticalFlowCalc *ofc, double targetFrameTime, double sourceFrameTime, double playbackSpeed,
                    double totalWarpDuration);

#endif /* HRCONTROL_H */
