/*
 * opticalFlowCalc.c — the host side of the optical-flow calculator, in C like the reference's
 * video/filter/HopperRender/opticalFlowCalc.c, but it only marshals into the CUDA C-ABI library
 * (include/hopperrender_cuda.h): no OpenCL, no runtime kernel compilation, no $HOME lookup
 * (reference :57, :366-378), no CPU fallback.
 */
#include "opticalFlowCalc.h"

#include <stdio.h>

#include "hopperrender_cuda.h"

#define CHECK_ERROR(err)                                                                                   \
    if (err) {                                                                                             \
        fprintf(stderr, "HopperRender CUDA error occurred in function: %s (%s)\n", __func__,               \
                hr_last_error((const HrContext *)ofc->impl));                                              \
        return 1;                                                                                          \
    }

/* reference :96-107 */
bool updateFrame(struct OpticalFlowCalc *ofc, unsigned char **inputPlanes) {
    if (!ofc->isInitialized) return 1;
    CHECK_ERROR(hr_update_frame((HrContext *)ofc->impl, inputPlanes[0], inputPlanes[1]));
    return 0;
}

/* reference :109-124 */
bool downloadFrame(struct OpticalFlowCalc *ofc, unsigned char **outputPlanes) {
    if (!ofc->isInitialized) return 1;
    double seconds = 0.0;
    CHECK_ERROR(hr_download((HrContext *)ofc->impl, outputPlanes[0], outputPlanes[1], &seconds));
    ofc->warpCalcTime = seconds;
    return 0;
}

/* reference :126-203: the search radius and both scalars are read from the struct at every call */
bool calculateOpticalFlow(struct OpticalFlowCalc *ofc) {
    if (!ofc->isInitialized) return 1;
    double seconds = 0.0;
    CHECK_ERROR(hr_calc_flow((HrContext *)ofc->impl, ofc->opticalFlowSearchRadius, ofc->deltaScalar, ofc->neighborBiasScalar, &seconds));
    ofc->ofcCalcTime = seconds;
    return 0;
}

/* reference :205-234 */
bool warpFrames(struct OpticalFlowCalc *ofc, const float blendingScalar, const int frameOutputMode) {
    if (!ofc->isInitialized) return 1;
    CHECK_ERROR(hr_warp((HrContext *)ofc->impl, blendingScalar, frameOutputMode, ofc->outputBlackLevel, ofc->outputWhiteLevel));
    return 0;
}

/* ---- device-resident frames (no reference counterpart) ---- */
bool updateFrameDevice(struct OpticalFlowCalc *ofc, unsigned char **devicePlanes) {
    if (!ofc->isInitialized) return 1;
    CHECK_ERROR(hr_update_frame_device((HrContext *)ofc->impl, devicePlanes[0], devicePlanes[1], 1));
    return 0;
}

bool warpFramesToDevice(struct OpticalFlowCalc *ofc, const float blendingScalar, const int frameOutputMode, unsigned char **devicePlanes) {
    if (!ofc->isInitialized) return 1;
    HrContext *ctx = (HrContext *)ofc->impl;
    CHECK_ERROR(hr_set_output_device(ctx, devicePlanes[0], devicePlanes[1]));
    const int failed = hr_warp(ctx, blendingScalar, frameOutputMode, ofc->outputBlackLevel, ofc->outputWhiteLevel);
    hr_set_output_device(ctx, NULL, NULL);
    CHECK_ERROR(failed);
    return 0;
}

bool finishFrames(struct OpticalFlowCalc *ofc) {
    if (!ofc->isInitialized) return 1;
    double seconds = 0.0;
    CHECK_ERROR(hr_finish((HrContext *)ofc->impl, &seconds));
    ofc->warpCalcTime = seconds;
    return 0;
}

unsigned char *allocHostPlanes(size_t bytes) {
    void *p = NULL;
    return hr_host_alloc(&p, bytes) ? NULL : (unsigned char *)p;
}

void freeHostPlanes(void *opaque, unsigned char *data) {
    (void)opaque;
    hr_host_free(data);
}

/* reference :236-253 (the struct itself belongs to the filter) */
void freeOFC(struct OpticalFlowCalc *ofc) {
    if (ofc->impl) hr_destroy((HrContext *)ofc->impl);
    ofc->impl = NULL;
    ofc->isInitialized = false;
}

/* reference :323-442 */
bool initOpticalFlowCalc(struct OpticalFlowCalc *ofc, const int frameHeight, const int frameWidth, const int actualWidth) {
    ofc->frameWidth = frameWidth;
    ofc->frameHeight = frameHeight;
    ofc->actualWidth = actualWidth;
    ofc->outputBlackLevel = 0.0f;
    ofc->outputWhiteLevel = 255.0f;
    ofc->opticalFlowSearchRadius = MIN_SEARCH_RADIUS;
    ofc->ofcCalcTime = 0.0;
    ofc->warpCalcTime = 0.0;
    ofc->deltaScalar = 8;
    ofc->neighborBiasScalar = 6;

    HrContext *ctx = NULL;
    /* the filter calloc()s the struct (vf_HopperRender.c:709): pixelFormat 0 = NV12, and
     * cudaDevice 0 means "not set" = current device; set cudaDevice = ordinal + 1 to pin one */
    if (hr_create(&ctx, frameHeight, frameWidth, actualWidth, ofc->pixelFormat, ofc->cudaDevice > 0 ? ofc->cudaDevice - 1 : -1)) {
        fprintf(stderr, "HopperRender CUDA error occurred in function: %s (%s)\n", __func__, hr_last_error(NULL));
        return 1;
    }
    /* the packed copy of a new frame is only needed by the NEXT pair's search: let it be built beside this pair's
     * search instead of in front of it (results are identical, include/hopperrender_cuda.h) */
    if (hr_set_pipeline(ctx, 1)) {
        fprintf(stderr, "HopperRender CUDA error occurred in function: %s (%s)\n", __func__, hr_last_error(ctx));
        hr_destroy(ctx);
        return 1;
    }
    HrInfo info;
    hr_get_info(ctx, &info);
    ofc->opticalFlowResScalar = info.resScalar;
    ofc->opticalFlowFrameWidth = info.lowWidth;
    ofc->opticalFlowFrameHeight = info.lowHeight;
    printf("[HopperRender] Using CUDA device %d and %zu MB of VRAM\n", info.device, info.deviceBytes / 1024 / 1024);
    ofc->impl = ctx;
    ofc->isInitialized = true;
    return 0;
}
