/*
 * hr_oracle.h — CPU ORACLE for the HopperRender hot path. TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library. The product path (libhopperrender_cuda.so) never links,
 * loads or calls anything in oracle/.
 *
 * It is a plain-C restatement, one loop iteration per OpenCL work-item, of the
 * reference's five kernels and their driver loop:
 *   video/filter/HopperRender/Kernels/calcDeltaSumsKernel.cl:34-189          (K1)
 *   video/filter/HopperRender/Kernels/determineLowestLayerKernel.cl:2-22     (K2)
 *   video/filter/HopperRender/Kernels/adjustOffsetArrayKernel.cl:2-18        (K3)
 *   video/filter/HopperRender/Kernels/blurFlowKernel.cl:5-12,77-88           (K4)
 *   video/filter/HopperRender/Kernels/warpFrameKernel.cl:1-182               (K5)
 *   video/filter/HopperRender/opticalFlowCalc.c:96-107,126-234,323-364       (driver)
 *
 * PARITY PIN STATUS: pinned. The reference ships no golden vectors, known-answer tests or
 * fixtures for this path, so the pin is the reference ITSELF: its unmodified
 * opticalFlowCalc.c + .cl kernels are built from /root/reference by oracle/build_ref.py into
 * oracle/_ref/ and executed on the B200 through the NVIDIA OpenCL ICD (oracle/ref_opencl.py).
 *   - tests/golden/reference_vectors.npz (made on the GPU box by tests/golden/make_reference_vectors.py)
 *     holds that run's offsets, window sums and output digests for five seeded cases;
 *     tests/test_oracle_golden_cpu.py checks this restatement against them bit for bit, no GPU needed.
 *   - tests/test_gpu_vs_reference_opencl.py compares oracle, reference and CUDA path live on the GPU box.
 *
 * Deliberate deviations from reference UB, identical in the
 * CUDA path:
 *   - K1's single mirror can leave the coordinate out of range when |offset| >= dim;
 *     the oracle clamps to [0, dim-1] afterwards instead of reading out of bounds.
 *   - K2 is evaluated only for in-lattice rows/cols (the reference's padded grid rows
 *     read/write past the arrays at window size 2).
 */
#ifndef HR_ORACLE_H
#define HR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HRO_PIXFMT_NV12 0 /* 8-bit, Y plane + interleaved UV plane                    */
#define HRO_PIXFMT_P010 1 /* 16-bit LE, 10 significant bits MSB-aligned (by construction) */

#define HRO_MAX_STEPS 32

/* How K5's float expressions are rounded. OpenCL C leaves this to the device compiler (contraction
 * allowed, '/' accurate to 2.5 ulp), so "the reference's result" exists per device:
 *   HRO_ARITH_IEEE  the literal reading: every operation rounded on its own, correctly rounded '/'.
 *   HRO_ARITH_NVCL  what the NVIDIA OpenCL compiler makes of the unmodified warpFrameKernel.cl
 *                   (PTX dumped on the B200 by tools/dump_ref_ptx.py): a*b+c*d -> fma(a, b, c*d),
 *                   x/y*255+128 -> fma(x/y, 255, 128), '/' -> div.full.f32 = x * MUFU.RCP(y).
 *                   Needs the MUFU.RCP table (hro_set_rcp_table; tests/golden/mufu_rcp_table.npy was
 *                   read back from the B200). Bit-identical to the reference run on that GPU for
 *                   modes 0,1,2,4,5,6; mode 3 within +-1 (the hue uses libm's atan2f/fmodf).
 * Integer results (flow, positions) do not depend on it. */
#define HRO_ARITH_IEEE 0
#define HRO_ARITH_NVCL 1

typedef struct HrOracle HrOracle;

/* frameHeight, frameWidth (= stride in samples), actualWidth: opticalFlowCalc.c:323-336 */
HrOracle *hro_create(int frameHeight, int frameWidth, int actualWidth, int pixfmt);
void hro_destroy(HrOracle *o);

int hro_low_width(const HrOracle *o);
int hro_low_height(const HrOracle *o);
int hro_res_scalar(const HrOracle *o);
int hro_num_steps(const HrOracle *o); /* 2 * iterations */

/* opticalFlowCalc.c:96-107: copy into slot 0 then swap, so slot 1 = newest frame. */
int hro_update_frame(HrOracle *o, const void *y, const void *uv);
/* opticalFlowCalc.c:126-203. Returns 0 on success. */
int hro_calc_flow(HrOracle *o, int searchRadius, int deltaScalar, int neighborBiasScalar);
/* table[i] = MUFU.RCP((float)i), 0 <= i < n; the pointer is kept, not copied */
void hro_set_rcp_table(const float *table, int n);
/* Returns 1 if NVCL is asked for without a table. Default: HRO_ARITH_IEEE. */
int hro_set_arith(HrOracle *o, int arith);
/* opticalFlowCalc.c:205-234 + warpFrameKernel.cl. Returns 1 when t > 1. */
int hro_warp(HrOracle *o, float t, int mode, float black, float white);
/* opticalFlowCalc.c:109-124 */
int hro_download(HrOracle *o, void *y, void *uv);

/* Parity taps. raw/blurred: int16 [2][lowH][lowW] (X plane then Y plane). */
int hro_get_offsets(const HrOracle *o, int16_t *raw, int16_t *blurred);
int hro_set_blurred_offsets(HrOracle *o, const int16_t *blurred);
/* lowestLayerArray as left by search step `step` (0..num_steps-1): uint8 [lowH][lowW],
 * valid at window representatives only. */
int hro_get_step_layers(const HrOracle *o, int step, uint8_t *layers);
/* summedDeltaValuesArray as left by the LAST search step: uint32 [radius][lowH][lowW]. */
int hro_get_last_sums(const HrOracle *o, uint32_t *sums);

/* Individual kernels on caller-owned arrays (used by unit tests). */
void hro_blur_flow(const int16_t *in, int16_t *out, int lowH, int lowW);

int hro_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
