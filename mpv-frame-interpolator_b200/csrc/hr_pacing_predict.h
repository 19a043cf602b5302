/*
 * hr_pacing_predict.h — which blending scalar will the filter ask for next? (host code only)
 *
 * The filter keeps a double, adds targetFrameTime / sourceFrameTime after every output and subtracts 1.0 when it
 * reaches 1.0 (video/filter/HopperRender/vf_HopperRender.c:371-374); a source frame gets the outputs up to the wrap
 * (:481). The library only ever sees that double rounded to float (warpFrames' argument, opticalFlowCalc.c:205). To
 * warp an output ahead of the call that asks for it, the float has to be guessed BIT FOR BIT — a warp for any other
 * scalar is thrown away (hr_warp compares with memcmp), so a wrong guess costs time, never correctness.
 *
 * "last scalar + last float increment" is right two times out of three at 24 -> 60 and one in five at 24 -> 144 (float
 * subtraction loses what the sum needs). Two trackers instead:
 *  - an interval tracker: every float seen pins the filter's double to the interval that rounds to it; intervals are
 *    carried forward by the ratio (itself an interval, narrowed by every observation against the narrowest one seen)
 *    and intersected. Scalars just after a wrap are tiny and pin the double to ~1e-23, so both converge quickly;
 *  - an exact follower: a scalar below 2^-30 (0.0 at the start of a stream, the accumulated rounding after a wrap) is
 *    the filter's double itself. From such an anchor candidate ratios are run through the filter's own two lines over
 *    the floats seen since and dropped at the first one they do not reproduce: every double the ratio interval still
 *    allows once it is a few ulps wide, or else the doubles around the simplest fraction inside it (frame rates are
 *    ratios of small integers; the filter's quotient of two reciprocals lands within an ulp or two of it). What
 *    survives reproduces the filter's arithmetic including its rounding: the residue the next tiny scalar consists
 *    of, and whether 6 x (1/6) reaches 1.0 and wraps or stops one ulp short and is passed on as 1.0f.
 * Anything irregular (a seek, a speed change, calls out of order) empties the interval or kills the candidates and the
 * trackers start over.
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

struct HrPacingPredictor {
    static constexpr int kCand = 33;
    static constexpr int kHist = 2048;
    /* interval tracker */
    int seen;         /* observations since the last reset */
    double lo, hi;    /* the filter's scalar after the last observation */
    double rLo, rHi;  /* the ratio */
    double aLo, aHi;  /* the narrowest observation so far (unwrapped: + wraps) and its index */
    long long aIdx, idx, wraps;
    float lastT;
    /* exact follower */
    int nCand;
    double candR[kCand], candB[kCand];
    bool candAlive[kCand];
    bool haveAnchor;
    double anchorB;    /* the filter's double at the anchor */
    int nHist, nextTry;
    float hist[kHist]; /* what was asked for since */

    void reset() { memset(this, 0, sizeof(*this)); }

    /* the doubles that round to t (round to nearest), as a closed interval inside [0, 1) */
    static void bounds_of(float t, double *lo, double *hi) {
        const float dn = std::nextafterf(t, -1.0f), up = std::nextafterf(t, 2.0f);
        *lo = t <= 0.0f ? 0.0 : 0.5 * ((double)dn + (double)t);
        *hi = t >= 1.0f ? std::nextafter(1.0, 0.0) : 0.5 * ((double)t + (double)up);
    }
    static double step(double b, double r, bool *wrapped) {
        b += r; /* vf_HopperRender.c:371 */
        *wrapped = b >= 1.0;
        if (b >= 1.0) b -= 1.0; /* :372-374 */
        return b;
    }
    /* the simplest fraction inside [a, b], 0 < a <= b < 1 (Stern-Brocot descent, long runs taken at once) */
    static bool simplest_fraction(double a, double b, double *p, double *q) {
        double pl = 0, ql = 1, ph = 1, qh = 1;
        for (int it = 0; it < 64; ++it) {
            const double pm = pl + ph, qm = ql + qh;
            if (pm < a * qm) { /* mediant below the interval: as many steps towards the high end as stay below */
                double k = std::floor((a * ql - pl) / (ph - a * qh));
                if (!(k >= 1)) k = 1;
                if (k > 1e9) return false;
                pl += k * ph, ql += k * qh;
                if (pl >= a * ql && pl <= b * ql) {
                    *p = pl, *q = ql;
                    return true;
                }
            } else if (pm > b * qm) {
                double k = std::floor((ph - b * qh) / (b * ql - pl));
                if (!(k >= 1)) k = 1;
                if (k > 1e9) return false;
                ph += k * pl, qh += k * ql;
                if (ph >= a * qh && ph <= b * qh) {
                    *p = ph, *q = qh;
                    return true;
                }
            } else {
                *p = pm, *q = qm;
                return true;
            }
        }
        return false;
    }
    void try_ratio(double r) {
        if (nCand >= kCand || !(r > 0.0)) return;
        for (int i = 0; i < nCand; ++i)
            if (candR[i] == r) return;
        double b = anchorB;
        for (int k = 0; k < nHist; ++k) {
            bool w;
            b = step(b, r, &w);
            if ((float)b != hist[k]) return;
        }
        candR[nCand] = r, candB[nCand] = b, candAlive[nCand] = true;
        ++nCand;
    }
    /* candidates for the exact follower, run from the anchor through everything seen since */
    void seed() {
        nCand = 0;
        const double mid = 0.5 * (rLo + rHi);
        if (!(mid > 0.0 && mid < 1.0)) return;
        const double ulp = std::nextafter(mid, 2.0) - mid;
        if (rHi - rLo < 24.0 * ulp) {
            double r = rLo;
            for (int k = 0; k < 4; ++k) r = std::nextafter(r, 0.0);
            for (; r <= rHi + 4.0 * ulp; r = std::nextafter(r, 2.0)) try_ratio(r);
            return;
        }
        double p, q;
        if (!simplest_fraction(rLo > 0.0 ? rLo : ulp, rHi, &p, &q)) return;
        double r = p / q;
        for (int k = 0; k < 4; ++k) r = std::nextafter(r, 0.0);
        for (int k = 0; k < 9; ++k, r = std::nextafter(r, 2.0)) try_ratio(r);
    }

    /* the filter asked for t */
    void observe(float t) {
        if (!(t >= 0.0f && t <= 1.0f)) { /* (1.0f is the double one ulp below 1.0, which does not wrap) */
            reset();
            return;
        }
        double oLo, oHi;
        bounds_of(t, &oLo, &oHi);
        if (seen == 0) {
            lo = oLo, hi = oHi;
            rLo = 0.0, rHi = 1.0;
            aLo = oLo, aHi = oHi, aIdx = 0, idx = 0, wraps = 0;
            nCand = 0;
        } else {
            ++idx;
            if (t < lastT) ++wraps;
            /* the ratio, against the narrowest observation (the filter's additions round: up to 1.1e-16 each, so at
             * most that on the mean) */
            const double n = (double)(idx - aIdx), uLo = oLo + (double)wraps, uHi = oHi + (double)wraps;
            const double qLo = (uLo - aHi) / n - 2.3e-16, qHi = (uHi - aLo) / n + 2.3e-16;
            if (qLo > rLo) rLo = qLo;
            if (qHi < rHi) rHi = qHi;
            /* the scalar: carried forward, then cut by what was seen */
            double pLo = lo + rLo, pHi = hi + rHi;
            if (t < lastT) pLo -= 1.0, pHi -= 1.0;
            pLo -= 4e-16, pHi += 4e-16; /* the filter's own rounding on the way */
            lo = pLo > oLo ? pLo : oLo;
            hi = pHi < oHi ? pHi : oHi;
            if (!(rLo <= rHi) || !(lo <= hi)) { /* not the sequence it was: start over from this observation */
                const float keep = t;
                reset();
                observe(keep);
                return;
            }
            if (oHi - oLo < (aHi - aLo) * 0.5) aLo = uLo, aHi = uHi, aIdx = idx;
            /* exact follower: advance, drop what does not reproduce the float */
            bool any = false;
            for (int i = 0; i < nCand; ++i) {
                if (!candAlive[i]) continue;
                bool w;
                candB[i] = step(candB[i], candR[i], &w);
                if ((float)candB[i] != t) candAlive[i] = false;
                any = any || candAlive[i];
            }
            if (!any) nCand = 0;
        }
        if (t < 9.3e-10f) { /* below 2^-30: the float is the filter's double */
            haveAnchor = true;
            anchorB = (double)t;
            nHist = 0;
        } else if (haveAnchor) {
            if (nHist < kHist) hist[nHist++] = t;
            else haveAnchor = false;
        }
        lastT = t;
        ++seen;
        if (nCand == 0 && haveAnchor && seen >= 4 && seen >= nextTry) {
            seed();
            if (nCand == 0) nextTry = seen + 24;
        }
    }

    /* the scalar of the next output and whether it belongs to the next source frame; false: no idea yet */
    bool predict(float *t, bool *nextFrame) const {
        if (seen < 2) return false;
        for (int i = 0; i < nCand; ++i) {
            if (!candAlive[i]) continue;
            *t = (float)step(candB[i], candR[i], nextFrame);
            return true;
        }
        *t = (float)step(0.5 * (lo + hi), 0.5 * (rLo + rHi), nextFrame);
        return *t >= 0.0f && *t <= 1.0f;
    }
};
