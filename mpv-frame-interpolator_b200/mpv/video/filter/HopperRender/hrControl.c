/* hrControl.c — see hrControl.h. Host-only C, no CUDA. */
#include "hrControl.h"

#include <ctype.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

int hrControlParse(const char *text) {
    if (!text || !isdigit((unsigned char)text[0])) return -1;
    return atoi(text);
}

struct Range {
    int first, last;
};
static int in(struct Range r, int code) { return code >= r.first && code <= r.last; }

int hrControlApply(struct OpticalFlowCalc *ofc, HrControlState *st, int code) {
    static const struct Range modes = {2, 8}, black = {100, 355}, white = {400, 655}, delta = {700, 731}, bias = {800, 831}, radius = {900, 932};
    static const float presets[3][2] = {{0.f, 255.f}, {10.f, 219.f}, {16.f, 219.f}};
    if (!ofc || !st) return 1;
    if (code == 0) {
        st->interpolationActive = 0;
        st->restartCounters = 1;
    } else if (code == 1) {
        st->interpolationActive = 1;
    } else if (in(modes, code)) {
        st->frameOutputMode = code - modes.first;
    } else if (code >= 9 && code <= 11) {
        ofc->outputBlackLevel = presets[code - 9][0];
        ofc->outputWhiteLevel = presets[code - 9][1];
    } else if (in(black, code)) {
        ofc->outputBlackLevel = (float)(code - black.first);
    } else if (in(white, code)) {
        ofc->outputWhiteLevel = (float)(code - white.first);
    } else if (in(delta, code)) {
        ofc->deltaScalar = code - delta.first;
    } else if (in(bias, code)) {
        ofc->neighborBiasScalar = code - bias.first;
    } else if (in(radius, code)) {
        const int r = code - radius.first;
        if (r != 0 && r < 2) return 1;
        st->pinnedRadius = r;
        if (r) ofc->opticalFlowSearchRadius = r;
    } else {
        return 1;
    }
    return 0;
}

/* one line of the channel: complete lines are applied; `last` without a terminator is parked in the state */
static int control_line(struct OpticalFlowCalc *ofc, HrControlState *st, const char *line, size_t len) {
    char text[sizeof(st->pending) + 4];
    if (len >= sizeof(text)) len = sizeof(text) - 1; /* a code has at most six digits: the tail of a longer line carries nothing */
    memcpy(text, line, len);
    text[len] = '\0';
    const int code = hrControlParse(text);
    return code >= 0 && hrControlApply(ofc, st, code) == 0;
}

int hrControlPoll(int fd, struct OpticalFlowCalc *ofc, HrControlState *st) {
    char buf[512 + sizeof(st->pending)];
    int applied = 0;
    if (!ofc || !st) return -1;
    if (st->pendingLength < 0 || st->pendingLength > (int)sizeof(st->pending)) st->pendingLength = 0;
    for (;;) {
        /* what the previous read left unfinished goes in front of what this one brings */
        size_t have = (size_t)st->pendingLength;
        memcpy(buf, st->pending, have);
        const ssize_t n = read(fd, buf + have, 512);
        if (n < 0) return (errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR) ? applied : -1;
        if (n == 0) { /* end of file: the writer will not finish the line */
            if (have) applied += control_line(ofc, st, buf, have);
            st->pendingLength = 0;
            return applied;
        }
        have += (size_t)n;
        size_t start = 0;
        for (size_t i = 0; i < have; ++i) {
            if (buf[i] == '\n' || buf[i] == '\r') {
                if (i > start) applied += control_line(ofc, st, buf + start, i - start);
                start = i + 1;
            }
        }
        size_t rest = have - start;
        if (rest > sizeof(st->pending)) { /* no code is that long: keep its head, which is all hrControlParse looks at */
            rest = sizeof(st->pending);
        }
        memcpy(st->pending, buf + start, rest);
        st->pendingLength = (int)rest;
        if (n < 512) return applied;
    }
}

int hrControlStatus(char *buf, size_t size, const struct OpticalFlowCalc *ofc, double targetFrameTime, double sourceFrameTime, double playbackSpeed,
                    double totalWarpDuration) {
    const double total = ofc->ofcCalcTime + totalWarpDuration;
    return snprintf(buf, size,
                    "Search Radius: %d\nCalc Res: %dx%d\nTarget Time: %06.2f ms (%.1f fps)\nFrame Time: %06.2f ms (%.3f fps | %.2fx)\n"
                    "Total Time: %06.2f ms (%.0f fps > %.3f fps)\nOFC Time: %06.2f ms (%.0f fps > %.3f fps)\nWarp Time: %06.2f ms (%.0f fps > %.3f fps)",
                    ofc->opticalFlowSearchRadius, ofc->frameWidth >> ofc->opticalFlowResScalar, ofc->frameHeight >> ofc->opticalFlowResScalar,
                    targetFrameTime * 1e3, 1.0 / targetFrameTime, sourceFrameTime * 1e3, 1.0 / sourceFrameTime, playbackSpeed, total * 1e3, 1.0 / total,
                    1.0 / sourceFrameTime, ofc->ofcCalcTime * 1e3, 1.0 / ofc->ofcCalcTime, 1.0 / sourceFrameTime, totalWarpDuration * 1e3,
                    1.0 / totalWarpDuration, 1.0 / sourceFrameTime);
}
