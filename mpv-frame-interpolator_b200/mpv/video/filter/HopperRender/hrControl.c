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

int hrControlPoll(int fd, struct OpticalFlowCalc *ofc, HrControlState *st) {
    char buf[512];
    int applied = 0;
    for (;;) {
        const ssize_t n = read(fd, buf, sizeof(buf) - 1);
        if (n < 0) return (errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR) ? applied : -1;
        if (n == 0) return applied;
        buf[n] = '\0';
        for (char *line = strtok(buf, "\r\n"); line; line = strtok(NULL, "\r\n")) {
            const int code = hrControlParse(line);
            if (code >= 0 && hrControlApply(ofc, st, code) == 0) ++applied;
        }
        if ((size_t)n < sizeof(buf) - 1) return applied;
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
