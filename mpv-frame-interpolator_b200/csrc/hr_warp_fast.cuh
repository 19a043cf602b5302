/*
 * hr_warp_fast.cuh — K5 (warpFrameKernel.cl:114-182) for the cases a player produces: output modes 0..5, resolution
 * scalar >= 2 (every frame taller than 540 lines), planes and strides aligned to the access width, level knobs for
 * which the reference's division is MUFU.RCP * x. Everything else runs warp_generic_kernel (hr_warp.cuh), which
 * computes the same numbers one sample at a time.
 *
 * HBM-bound by nature (2 frames read, 1 written, 518 KB of flow), issue- and load-request-bound in practice. The
 * kernel is organised around both counts:
 *   - thread = UW samples x ROWS rows inside ONE lattice cell, UW chosen per resolution scalar so that a unit is as
 *     wide as one displacement allows: a cell is 4 samples wide at 1080p, 8 at 4K, 16 at 8K. A unit row is one access
 *     of UB = UW * sizeof(sample) bytes: 32 bits (NV12 1080p), 64 bits (P010 1080p, NV12 4K) or 128 bits (P010 4K and 8K,
 *     NV12 8K); the flow vector, the flipped vector (warpFrameKernel.cl:155-156), the four roundings and every bounds
 *     test are done once per unit. The flow comes as one packed (x | y << 16) word per cell.
 *   - a source run = UW consecutive samples at an arbitrary displacement = TWO aligned UB-byte loads (the vector that
 *     holds the run's first sample and the next one) and a register barrel: word select by the offset's word part,
 *     funnel shift by its byte part. The row pitch is a multiple of UB, so the barrel setting is the same for every
 *     row of the unit. Chroma with an odd displacement (warpFrameKernel.cl:171 takes U from column c+d-1 and V from
 *     c+d+1) reads one word more from the same two vectors and fixes the pairs with one byte permute per word.
 *   - conversions through the 2^23 magic number, with the subtractions folded into the arithmetic where
 *     that is exact:  fl(b * t) == fma(2^23 + b, t, -(2^23 * t))   (one rounding of the exact product),
 *                     v - sub   == (2^23 + v) - (2^23 + sub)       (integers below 2^24),
 *     all of it two samples per instruction on the packed fp32 pipe (FADD2 / FMUL2 / FFMA2);
 *   - output modes: 0 / 1 copy a run, 2 / 5 blend + levels, 3 (HSV flow) adds the cell's colour from the per-flow
 *     table (hr_warp.cuh) to half of the blend — its chroma plane is a per-cell constant and reads no picture at
 *     all —, 4 (grey flow) is a per-cell constant;
 *   - several output frames of one frame pair (same flow, same sources, different blend scalars) in ONE launch:
 *     grid.z = output (WarpBatch), so a 24->60 source frame costs one warp launch instead of two or three.
 * The float expressions are those of hr_warp.cuh (header there): bit-identical to the reference kernel
 * as the NVIDIA OpenCL compiler builds it.
 */
#pragma once
#include "hr_warp.cuh"
#include <type_traits>

/* Constants of the packed blend + level arithmetic as fp32 pairs (both halves equal), prepared by the host so that the
 * kernel reads them as 64-bit constant-bank operands instead of building and holding them in registers. */
struct LevelConsts {           /* per plane kind */
    F2 rcp, mul, add, negSub, negMsub;
    float lo, hi;
};
#define HR_WARP_BATCH 8
struct WarpBatch {
    int n;
    float t12[HR_WARP_BATCH], t21[HR_WARP_BATCH];
    F2 t12x2[HR_WARP_BATCH], t21x2[HR_WARP_BATCH], negMt12[HR_WARP_BATCH]; /* (t, t), (1-t, 1-t), -(2^23 * t) twice */
    void *outY[HR_WARP_BATCH], *outUV[HR_WARP_BATCH];
};

struct WarpFastArgs {
    int lumaGroups, lumaG0, chromaG0, chromaGN; /* row groups of the launch: luma first, then chroma       */
    int unitsX, edgeBlocks, coreBlocksX;        /* unit columns; CTAs of the two edge columns; CTAs per row of the rest */
    /* level map, per plane kind [0] luma [1] chroma. 8-bit (NV12, and the HSV mode of either format): sub = black /
     * 128, rcp = MUFU.RCP of den (read back from the device by the host once per knob setting). 16-bit: sub = b16 /
     * 32768, rcp = correctly rounded reciprocal (host). */
    float sub[2], den[2], rcp[2];
    int clampNeeded[2];    /* the map can leave [0, max]: clamp in float before the truncation             */
    int subIsInt[2];       /* sub is an integer below 2^22: (2^23 + v) - (2^23 + sub) is exact              */
    const uint32_t *colours; /* HSV mode: flow colour per lattice cell (flow_colour_kernel)                 */
    float white;           /* HSV mode: the chroma constants go through the 8-bit level map per unit        */
    LevelConsts lc[2];     /* the same level map as packed constants ([0] luma, [1] chroma)                 */
    F2 negM, M;            /* (-2^23, -2^23), (2^23, 2^23)                                                   */
};

/* ---- aligned vectors of UB bytes -------------------------------------------------------------------------- */
template <int UB>
struct Vec {
    uint32_t w[UB / 4];
};
template <int UB>
__device__ __forceinline__ Vec<UB> vec_load(const unsigned char *p);
template <>
__device__ __forceinline__ Vec<4> vec_load<4>(const unsigned char *p) {
    Vec<4> v;
    v.w[0] = __ldg(reinterpret_cast<const uint32_t *>(p));
    return v;
}
template <>
__device__ __forceinline__ Vec<8> vec_load<8>(const unsigned char *p) {
    const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
    Vec<8> v;
    v.w[0] = t.x;
    v.w[1] = t.y;
    return v;
}
template <>
__device__ __forceinline__ Vec<16> vec_load<16>(const unsigned char *p) {
    const uint4 t = __ldg(reinterpret_cast<const uint4 *>(p));
    Vec<16> v;
    v.w[0] = t.x;
    v.w[1] = t.y;
    v.w[2] = t.z;
    v.w[3] = t.w;
    return v;
}
__device__ __forceinline__ void vec_store(unsigned char *p, const Vec<4> &v) { *reinterpret_cast<uint32_t *>(p) = v.w[0]; }
__device__ __forceinline__ void vec_store(unsigned char *p, const Vec<8> &v) { *reinterpret_cast<uint2 *>(p) = make_uint2(v.w[0], v.w[1]); }
__device__ __forceinline__ void vec_store(unsigned char *p, const Vec<16> &v) { *reinterpret_cast<uint4 *>(p) = make_uint4(v.w[0], v.w[1], v.w[2], v.w[3]); }

/* bytes per source load by unit size (tools/diag_warp.py): HR_WARP_LG8 in {4, 8}, HR_WARP_LG16 in {4, 8, 16} */
#ifndef HR_WARP_LG8
#define HR_WARP_LG8 4
#endif
#ifndef HR_WARP_LG16
#define HR_WARP_LG16 8
#endif
__host__ __device__ constexpr int warp_load_bytes(int unitBytes) { return unitBytes >= 16 ? HR_WARP_LG16 : (unitBytes >= 8 ? HR_WARP_LG8 : 4); }

/* ---- source runs ----------------------------------------------------------------------------------------- */
/* Where a unit's source block starts, and how it is fetched: aligned loads of LG bytes (LG <= UB; LG = UB: two loads per
 * row, LG = 4: one per word) from the one that holds the run's first sample; the run's word and bit offset inside it.
 * An odd chroma displacement starts one sample earlier (pair-aligned) and reads one word more. */
template <int UB, int LG, bool IS16>
struct RunSource {
    static constexpr int NW = UB / 4;          /* words per run                          */
    static constexpr int LW = LG / 4;          /* words per load                         */
    static constexpr int NL = UB / LG + 1;     /* loads that can be needed               */
    const unsigned char *base;
    unsigned wordOff, bitOff;
    bool last, odd;
    size_t pitch;
    __device__ __forceinline__ void set(const void *plane, long long firstSample, size_t pitchBytes, bool oddDisplacement) {
        odd = oddDisplacement;
        if (odd) firstSample -= 1;
        const uintptr_t a = (uintptr_t)plane + (uintptr_t)(firstSample * (IS16 ? 2 : 1));
        const unsigned k = (unsigned)(a & (LG - 1));
        base = reinterpret_cast<const unsigned char *>(a - k);
        wordOff = k >> 2;
        bitOff = (k & 3u) * 8u;
        last = k != 0 || odd; /* the run ends in the last load */
        pitch = pitchBytes;
    }
    /* the UW samples of row r, in output order */
    template <bool CHROMA>
    __device__ __forceinline__ Vec<UB> row(int r) const {
        const unsigned char *p = base + (size_t)r * pitch;
        uint32_t w[NL * LW + 2];
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            Vec<LG> v;
#pragma unroll
            for (int i = 0; i < LW; ++i) v.w[i] = 0u;
            if (l + 1 < NL || last) v = vec_load<LG>(p + l * LG);
#pragma unroll
            for (int i = 0; i < LW; ++i) w[l * LW + i] = v.w[i];
        }
        w[NL * LW] = w[NL * LW + 1] = 0u;
        /* barrel: shift the words down by wordOff (one conditional move per word and offset bit) ... */
        if (LW >= 4) {
            const bool by2 = wordOff & 2u;
#pragma unroll
            for (int i = 0; i < NL * LW; ++i) w[i] = by2 ? w[i + 2] : w[i];
        }
        if (LW >= 2) {
            const bool by1 = wordOff & 1u;
#pragma unroll
            for (int i = 0; i < NL * LW; ++i) w[i] = by1 ? w[i + 1] : w[i];
        }
        /* ... and by bitOff inside the words */
        uint32_t o[NW + 1];
#pragma unroll
        for (int i = 0; i < NW + 1; ++i) o[i] = __funnelshift_r(w[i], w[i + 1], bitOff);
        Vec<UB> out;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            /* odd chroma displacement: sample j of the output pair (U, V) comes from run sample j (U) and j + 2 (V) */
            if (CHROMA) out.w[i] = __byte_perm(o[i], o[i + 1], odd ? (IS16 ? 0x7610u : 0x5230u) : 0x3210u);
            else out.w[i] = o[i];
        }
        return out;
    }
};

/* UW samples at a frame border: each through reflect_inner() and, for chroma, the pair rule */
template <int UB, bool IS16>
__device__ __forceinline__ Vec<UB> border_row(const void *rowBase, int cx0, int dx, int aW, bool chroma) {
    constexpr int UW = UB / (IS16 ? 2 : 1);
    Vec<UB> v;
#pragma unroll
    for (int i = 0; i < UB / 4; ++i) v.w[i] = 0u;
#pragma unroll
    for (int k = 0; k < UW; ++k) {
        const int x = reflect_inner(cx0 + k + dx, aW);
        const int col = chroma ? (x & ~1) + (k & 1) : x;
        if (IS16) v.w[k >> 1] |= (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(rowBase) + col) << (16 * (k & 1));
        else v.w[k >> 2] |= (uint32_t)__ldg(reinterpret_cast<const uint8_t *>(rowBase) + col) << (8 * (k & 3));
    }
    return v;
}

/* ---- blend + levels, two samples per instruction ----------------------------------------------------- */
/* what a unit needs of the constants: references into the kernel's parameter space plus its own HSV term */
struct BlendK {
    const LevelConsts &L;
    const F2 &t12, &t21, &negMt12, &negM, &M;
    uint32_t lumaColour; /* HSV mode: 0x4B000000 + the cell's halved colour luma */
};
__device__ __forceinline__ uint32_t f2_lo_bits(F2 v) { return (uint32_t)v; }
__device__ __forceinline__ uint32_t f2_hi_bits(F2 v) { return (uint32_t)(v >> 32); }

/* one pair: A, B = the bit patterns 0x4B000000 | sample of the frame-1 / frame-2 samples; returns the bit
 * patterns 0x4B000000 | output of the pair. HSV (luma only): the blend's 8-bit reading is halved and the cell's
 * colour added before the level map (hr_warp.cuh, blend_finish); PICSHIFT = 1 (8-bit samples) or 9 (16-bit). */
template <bool CHROMA, bool CLAMP, bool SUBINT, int PICSHIFT>
__device__ __forceinline__ F2 blend_pair(const BlendK &K, F2 A, F2 B) {
    const LevelConsts &L = K.L;
    const F2 a = f2_add(A, K.negM);                         /* (float)f1                                  */
    const F2 p = f2_fma(B, K.t12, K.negMt12);               /* (float)f2 * t12, one rounding               */
    const F2 bl = f2_fma(a, K.t21, p);                      /* fma(f1, t21, f2 * t12)                      */
    F2 vM = f2_add_rz(bl, K.M);                             /* 2^23 + trunc(blend)                         */
    if (PICSHIFT) {
        vM = f2_bits(((f2_lo_bits(vM) >> PICSHIFT) & 0x7Fu) + K.lumaColour, ((f2_hi_bits(vM) >> PICSHIFT) & 0x7Fu) + K.lumaColour);
    }
    const F2 d = SUBINT ? f2_add(vM, L.negMsub) : f2_add(f2_add(vM, K.negM), L.negSub); /* v - sub            */
    F2 x = f2_mul(d, L.rcp);
    x = CHROMA ? f2_fma(x, L.mul, L.add) : f2_mul(x, L.mul);
    if (CLAMP) x = f2_make(fmaxf(fminf(f2_lo(x), L.hi), L.lo), fmaxf(fminf(f2_hi(x), L.hi), L.lo));
    return f2_add_rz(x, K.M);
}

/* one 32-bit word of samples. 8-bit: four samples in, four out. */
template <bool CHROMA, bool CLAMP, bool SUBINT, bool HSV>
__device__ __forceinline__ uint32_t blend_word8(const BlendK &K, uint32_t wa, uint32_t wb) {
    constexpr int PS = HSV ? 1 : 0;
    const F2 r0 = blend_pair<CHROMA, CLAMP, SUBINT, PS>(K, f2_bits(__byte_perm(wa, 0x4B000000u, 0x7540u), __byte_perm(wa, 0x4B000000u, 0x7541u)),
                                                        f2_bits(__byte_perm(wb, 0x4B000000u, 0x7540u), __byte_perm(wb, 0x4B000000u, 0x7541u)));
    const F2 r1 = blend_pair<CHROMA, CLAMP, SUBINT, PS>(K, f2_bits(__byte_perm(wa, 0x4B000000u, 0x7542u), __byte_perm(wa, 0x4B000000u, 0x7543u)),
                                                        f2_bits(__byte_perm(wb, 0x4B000000u, 0x7542u), __byte_perm(wb, 0x4B000000u, 0x7543u)));
    return __byte_perm(__byte_perm(f2_lo_bits(r0), f2_hi_bits(r0), 0x0040), __byte_perm(f2_lo_bits(r1), f2_hi_bits(r1), 0x0040), 0x5410);
}
/* 16-bit: two samples per word. The low 16 bits of each result hold trunc(x) < 65536; clamp to 65472, round to the
 * nearest 10-bit code ((v + 32) & 0xFFC0) on both halves at once. HSV: the 8-bit result, << 8. */
template <bool CHROMA, bool CLAMP, bool SUBINT, bool HSV>
__device__ __forceinline__ uint32_t blend_word16(const BlendK &K, uint32_t wa, uint32_t wb) {
    constexpr int PS = HSV ? 9 : 0;
    const F2 r = blend_pair<CHROMA, CLAMP, SUBINT, PS>(K, f2_bits(__byte_perm(wa, 0x4B000000u, 0x7510u), __byte_perm(wa, 0x4B000000u, 0x7532u)),
                                                       f2_bits(__byte_perm(wb, 0x4B000000u, 0x7510u), __byte_perm(wb, 0x4B000000u, 0x7532u)));
    if (HSV) return __byte_perm(f2_lo_bits(r), f2_hi_bits(r), 0x4101u); /* bytes: 0, lo, 0, hi */
    const uint32_t p = __vminu2(__byte_perm(f2_lo_bits(r), f2_hi_bits(r), 0x5410), 0xFFC0FFC0u);
    return (p + 0x00200020u) & 0xFFC0FFC0u;
}
template <bool IS16, int UB, bool CHROMA, bool CLAMP, bool SUBINT, bool HSV>
__device__ __forceinline__ Vec<UB> blend_vec(const BlendK &K, const Vec<UB> &a, const Vec<UB> &b) {
    Vec<UB> o;
#pragma unroll
    for (int i = 0; i < UB / 4; ++i)
        o.w[i] = IS16 ? blend_word16<CHROMA, CLAMP, SUBINT, HSV>(K, a.w[i], b.w[i]) : blend_word8<CHROMA, CLAMP, SUBINT, HSV>(K, a.w[i], b.w[i]);
    return o;
}

/* rare units: a partial column / row group, or the column that straddles the middle in mode 5 */
template <typename T>
__device__ __noinline__ void unit_by_samples(const WarpParams<T> &P, float t12, float t21, T *out, int cx0, int cy0, bool chroma, int ncols, int nrows) {
    for (int r = 0; r < nrows; ++r)
        for (int k = 0; k < ncols; ++k)
            if (cx0 + k < P.aW) out[(size_t)(cy0 + r) * P.W + cx0 + k] = (T)sample_value(P, t12, t21, cx0 + k, cy0 + r, chroma);
}

/* A unit whose source blocks touch a frame border: every sample through the reflection, chroma through the pair rule;
 * the blend in its general form (float clamp, unfolded subtraction: the same numbers as the specialised forms). Out of
 * line, so that the interior path does not carry its registers. */
template <typename T, int ROWS, int UW, bool CHROMA>
__device__ __noinline__ void border_unit(const WarpParams<T> &P, const WarpFastArgs &A, const WarpBatch &B, int z, uint32_t lumaColour, unsigned char *po, int cx0,
                                         Shift d, int b12, int b21) {
    constexpr bool is16 = SampleTraits<T>::is16;
    const BlendK K = {A.lc[CHROMA ? 1 : 0], B.t12x2[z], B.t21x2[z], B.negMt12[z], A.negM, A.M, lumaColour};
    constexpr int UB = UW * (int)sizeof(T);
    typedef Vec<UB> Run;
    const int planeH = CHROMA ? (P.H >> 1) : P.H;
    const T *s12 = CHROMA ? P.f1uv : P.f1y;
    const T *s21 = CHROMA ? P.f2uv : P.f2y;
    const size_t pitch = (size_t)P.W * sizeof(T);
    const int mode = P.mode;
    /* every row's samples first (all loads in flight together: one memory round trip per unit, like the interior
     * path), then the arithmetic */
    Run ra[ROWS], rb[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        ra[r] = Run();
        rb[r] = Run();
        if (mode != HR_MODE_WARPED_21) ra[r] = border_row<UB, is16>(s12 + (size_t)reflect_inner(b12 + r, planeH) * P.W, cx0, d.x1, P.aW, CHROMA);
        if (mode != HR_MODE_WARPED_12) rb[r] = border_row<UB, is16>(s21 + (size_t)reflect_inner(b21 + r, planeH) * P.W, cx0, d.x2, P.aW, CHROMA);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const Run a = ra[r], b = rb[r];
        Run o;
        if (mode == HR_MODE_WARPED_12) o = a;
        else if (mode == HR_MODE_WARPED_21) o = b;
        else if (mode == HR_MODE_HSV_FLOW) o = blend_vec<is16, UB, false, true, false, true>(K, a, b); /* luma only: chroma never gets here */
        else o = blend_vec<is16, UB, CHROMA, true, false, false>(K, a, b);
        vec_store(po + r * pitch, o);
    }
}

template <typename T, int ROWS, int UW, bool CHROMA>
__device__ __forceinline__ void warp_unit(const WarpParams<T> &P, const WarpFastArgs &A, const WarpBatch &B, int z, T *outPlane, int cx0, int cy0) {
    const float t12 = B.t12[z], t21 = B.t21[z];
    constexpr bool is16 = SampleTraits<T>::is16;
    constexpr int UB = UW * (int)sizeof(T);
    typedef Vec<UB> Run;
    constexpr int cz = CHROMA ? 1 : 0;
    const int planeH = CHROMA ? (P.H >> 1) : P.H;
    const T *s12 = CHROMA ? P.f1uv : P.f1y;
    const T *s21 = CHROMA ? P.f2uv : P.f2y;
    const size_t pitch = (size_t)P.W * sizeof(T);
    unsigned char *po = reinterpret_cast<unsigned char *>(outPlane + (size_t)cy0 * P.W + cx0);
    const int mode = P.mode;

    if (cx0 + UW > P.aW || cy0 + ROWS > planeH) {
        unit_by_samples(P, t12, t21, outPlane, cx0, cy0, CHROMA, UW, hr_min(ROWS, planeH - cy0));
        return;
    }
    if (mode == HR_MODE_SIDE_BY_SIDE_1) {
        const int half = P.aW >> 1;
        if (cx0 + UW <= half) { /* left half: frame 1 as it is (warpFrameKernel.cl:131-133) */
            const unsigned char *ps = reinterpret_cast<const unsigned char *>(s12 + (size_t)cy0 * P.W + cx0);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) vec_store(po + r * pitch, vec_load<UB>(ps + r * pitch));
            return;
        }
        if (cx0 < half) {
            unit_by_samples(P, t12, t21, outPlane, cx0, cy0, CHROMA, UW, ROWS);
            return;
        }
    }

    /* the cell's flow and its flip (warpFrameKernel.cl:151-156) */
    const int2 cell = lattice_cell(cx0, cy0, P.s, CHROMA);
    if (mode == HR_MODE_HSV_FLOW && CHROMA) {
        /* flow colours: the chroma plane of a cell is one (U, V) pair, through the 8-bit level map */
        const uint32_t colour = __ldg(A.colours + cell.y * P.lw + cell.x);
        const unsigned u = levels_uv8((float)((colour >> 8) & 255u), A.white), v = levels_uv8((float)((colour >> 16) & 255u), A.white);
        Run fill;
#pragma unroll
        for (int i = 0; i < UB / 4; ++i) fill.w[i] = is16 ? (u << 8) | (v << 24) : u | (v << 8) | (u << 16) | (v << 24);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) vec_store(po + r * pitch, fill);
        return;
    }
    const VectorPair v = vector_pair(P, cell);
    if (mode == HR_MODE_GREY_FLOW) {
        const unsigned len4 = (unsigned)(abs(v.fx) + abs(v.fy)) << 2;
        const unsigned g8 = CHROMA ? 128u : (len4 < 255u ? len4 : 255u);
        Run fill;
#pragma unroll
        for (int i = 0; i < UB / 4; ++i) fill.w[i] = is16 ? (g8 << 8) * 0x00010001u : g8 * 0x01010101u;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) vec_store(po + r * pitch, fill);
        return;
    }

    /* displacements (warpFrameKernel.cl:165-168) */
    const Shift d = shift_for(v, t12, t21, CHROMA);
    const int a12 = cx0 + d.x1, a21 = cx0 + d.x2, b12 = cy0 + d.y1, b21 = cy0 + d.y2;
    /* all four source blocks inside [1, aW-2] x [1, planeH-2]: reflect_inner() is the identity, rows and columns are
     * consecutive */
    const bool interior = (unsigned)(a12 - 1) <= (unsigned)(P.aW - UW - 2) && (unsigned)(a21 - 1) <= (unsigned)(P.aW - UW - 2) &&
                          (unsigned)(b12 - 1) <= (unsigned)(planeH - ROWS - 2) && (unsigned)(b21 - 1) <= (unsigned)(planeH - ROWS - 2);
    const bool hsv = mode == HR_MODE_HSV_FLOW;
    uint32_t lumaColour = 0x4B000000u;
    if (hsv) lumaColour += __ldg(A.colours + cell.y * P.lw + cell.x) & 255u;
    if (!interior) {
        border_unit<T, ROWS, UW, CHROMA>(P, A, B, z, lumaColour, po, cx0, d, b12, b21);
        return;
    }
    const BlendK K = {A.lc[cz], B.t12x2[z], B.t21x2[z], B.negMt12[z], A.negM, A.M, lumaColour};
    /* one output row from its two source runs. VAR 0: WarpedFrame12 / WarpedFrame21, the run as it is
     * (warpFrameKernel.cl:170-173); 1..3: blend + levels (no clamp / clamp / clamp and non-integer black); 4: HSV luma */
    const int var = mode < 2 ? 0 : (hsv ? 4 : (A.subIsInt[cz] ? (A.clampNeeded[cz] ? 2 : 1) : 3));
    auto rows = [&](auto varTag) {
        constexpr int VAR = decltype(varTag)::value;
        Run ra[ROWS], rb[ROWS];
        RunSource<UB, warp_load_bytes(UB), is16> A12, A21;
        if (VAR != 0 || mode == 0) {
            A12.set(s12, (long long)b12 * P.W + a12, pitch, CHROMA && (d.x1 & 1));
#pragma unroll
            for (int r = 0; r < ROWS; ++r) ra[r] = A12.template row<CHROMA>(r);
        }
        if (VAR != 0 || mode == 1) {
            A21.set(s21, (long long)b21 * P.W + a21, pitch, CHROMA && (d.x2 & 1));
#pragma unroll
            for (int r = 0; r < ROWS; ++r) rb[r] = A21.template row<CHROMA>(r);
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            if (VAR == 0) vec_store(po + r * pitch, mode == 0 ? ra[r] : rb[r]);
            else if (VAR == 4) vec_store(po + r * pitch, blend_vec<is16, UB, CHROMA, true, false, true>(K, ra[r], rb[r]));
            else vec_store(po + r * pitch, blend_vec<is16, UB, CHROMA, VAR >= 2, VAR != 3, false>(K, ra[r], rb[r]));
        }
    };
    switch (var) {
        case 0: rows(std::integral_constant<int, 0>()); break;
        case 1: rows(std::integral_constant<int, 1>()); break;
        case 2: rows(std::integral_constant<int, 2>()); break;
        case 3: rows(std::integral_constant<int, 3>()); break;
        default:
            if (!CHROMA) rows(std::integral_constant<int, 4>());
            break;
    }
}

/* grid: x = CTAs (edge columns first, see the kernel), z = output frame of the batch; row groups of the luma plane
 * first, then of the chroma plane.
 * HR_WARP_MINBLOCKS 128-thread CTAs per SM: the kernel is latency-bound at 4K and above, resident warps are what hides
 * its three dependent round trips (flow word -> flipped flow word -> samples). */
/* rows per unit and CTAs per SM by unit size in bytes (tools/diag_warp.py sweeps them through tools/build_variant.py) */
#ifndef HR_WARP_ROWS4
#define HR_WARP_ROWS4 4
#endif
#ifndef HR_WARP_ROWS8
#define HR_WARP_ROWS8 4
#endif
#ifndef HR_WARP_ROWS16
#define HR_WARP_ROWS16 4
#endif
#ifndef HR_WARP_MB4
#define HR_WARP_MB4 12
#endif
#ifndef HR_WARP_MB8
#define HR_WARP_MB8 12
#endif
#ifndef HR_WARP_MB16
#define HR_WARP_MB16 8
#endif
__host__ __device__ constexpr int warp_rows(int unitBytes) { return unitBytes >= 16 ? HR_WARP_ROWS16 : (unitBytes >= 8 ? HR_WARP_ROWS8 : HR_WARP_ROWS4); }
__host__ __device__ constexpr int warp_min_blocks(int unitBytes) { return unitBytes >= 16 ? HR_WARP_MB16 : (unitBytes >= 8 ? HR_WARP_MB8 : HR_WARP_MB4); }
template <typename T, int UW>
__global__ void __launch_bounds__(128, warp_min_blocks(UW * (int)sizeof(T)))
    warp_fast_kernel(const __grid_constant__ WarpParams<T> P, const __grid_constant__ WarpFastArgs A, const __grid_constant__ WarpBatch B) {
    constexpr int ROWS = warp_rows(UW * (int)sizeof(T));
    /* The first and the last unit column always touch the picture's border (an undisplaced column 0 reads column 1,
     * hr_warp.cuh reflect_inner): their units take the reflected path, which is several times longer than the interior
     * one, and a single such unit would drag its whole warp through it. They get CTAs of their own — the first
     * A.edgeBlocks — as single-row units, lane = (row, side), so that the long path is spread over ROWS times as many
     * threads and is not what the launch waits for. The other columns follow, 32 units x 4 row groups per CTA. */
    const int z = blockIdx.z;
    if ((int)blockIdx.x < A.edgeBlocks) {
        const int e = blockIdx.x * 128 + threadIdx.y * 32 + threadIdx.x;
        const int row = e >> 1;                       /* row of the launch: luma rows first, then chroma rows */
        const int cx0 = (e & 1) ? (A.unitsX - 1) * UW : 0;
        const int lumaRows = A.lumaGroups * ROWS;
        if (row < lumaRows) {
            const int cy0 = A.lumaG0 * ROWS + row;
            if (cy0 < P.H) warp_unit<T, 1, UW, false>(P, A, B, z, (T *)B.outY[z], cx0, cy0);
        } else if (row - lumaRows < A.chromaGN * ROWS) {
            const int cy0 = A.chromaG0 * ROWS + row - lumaRows;
            if (cy0 < (P.H >> 1)) warp_unit<T, 1, UW, true>(P, A, B, z, (T *)B.outUV[z], cx0, cy0);
        }
        return;
    }
    const int b = blockIdx.x - A.edgeBlocks;
    const int by = b / A.coreBlocksX;
    const int ux = 1 + (b - by * A.coreBlocksX) * 32 + threadIdx.x;
    if (ux >= A.unitsX - 1) return;
    const int rg = by * 4 + threadIdx.y;
    const int cx0 = ux * UW;
    if (rg < A.lumaGroups) {
        warp_unit<T, ROWS, UW, false>(P, A, B, z, (T *)B.outY[z], cx0, (A.lumaG0 + rg) * ROWS);
    } else if (rg - A.lumaGroups < A.chromaGN) {
        const int cy0 = (A.chromaG0 + rg - A.lumaGroups) * ROWS;
        if (cy0 < (P.H >> 1)) warp_unit<T, ROWS, UW, true>(P, A, B, z, (T *)B.outUV[z], cx0, cy0);
    }
}

/* packed copy of a planar blurred flow (parity tap hr_set_blurred_offsets) */
__global__ void pack_flow_kernel(const int16_t *__restrict__ flow, uint32_t *__restrict__ flowXY, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flowXY[i] = (uint32_t)(uint16_t)flow[i] | ((uint32_t)(uint16_t)flow[n + i] << 16);
}
