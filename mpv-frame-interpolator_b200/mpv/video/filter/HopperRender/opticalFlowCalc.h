/*
 * opticalFlowCalc.h — drop-in replacement for the reference's
 * video/filter/HopperRender/opticalFlowCalc.h:1-126.
 *
 * Same six functions, same argument meaning, same inverted-bool convention (0/false = success,
 * 1/true = failure), same public struct fields by name. The OpenCL members (cl_device_id ...
 * cl_kernel, reference :30-64) are replaced by one opaque handle to the CUDA C-ABI library
 * (include/hopperrender_cuda.h); <CL/cl.h> is no longer needed.
 */
#ifndef OPTICALFLOWCALC_H
#define OPTICALFLOWCALC_H

#include <stdbool.h>
#include <stddef.h>

#include "config.h"

typedef struct OpticalFlowCalc {
    /* Video properties (reference :12-17) */
    bool isInitialized;
    int frameWidth;          /* stride of the frame, in samples */
    int frameHeight;
    int actualWidth;         /* width as encoded */
    float outputBlackLevel;
    float outputWhiteLevel;

    /* Optical flow calculation (reference :20-27) */
    int opticalFlowResScalar;
    int opticalFlowFrameWidth;
    int opticalFlowFrameHeight;
    int opticalFlowSearchRadius;
    double ofcCalcTime;
    double warpCalcTime;
    int deltaScalar;
    int neighborBiasScalar;

    /* CUDA implementation (replaces reference :30-64) */
    int pixelFormat;         /* 0 = NV12 (what the reference negotiates), 1 = P010; set before init */
    int cudaDevice;          /* 0 = current device, n = device ordinal n-1; set before init */
    void *impl;              /* HrContext* */
} OpticalFlowCalc;

bool initOpticalFlowCalc(struct OpticalFlowCalc *ofc, const int frameHeight, const int frameWidth, const int actualWidth);
void freeOFC(struct OpticalFlowCalc *ofc);
bool updateFrame(struct OpticalFlowCalc *ofc, unsigned char **inputPlanes);
bool downloadFrame(struct OpticalFlowCalc *ofc, unsigned char **outputPlanes);
bool calculateOpticalFlow(struct OpticalFlowCalc *ofc);
bool warpFrames(struct OpticalFlowCalc *ofc, const float blendingScalar, const int frameOutputMode);

/* Frames that live in device memory (IMGFMT_CUDA in and out, video/img_format.h:274; no reference counterpart — the
 * reference moves every frame through host memory, reference :98-100, :112-114). Same convention: 0 = success.
 * devicePlanes[0] = Y, [1] = interleaved UV, CUDA device pointers with the stride given at init.
 *   updateFrameDevice   the new source frame is used in place: the caller keeps it alive and unchanged until two further
 *                       updateFrame* calls have been made (the filter holds a reference to the last two source images)
 *   warpFramesToDevice  warpFrames into a device image (not the source images of the pair); enqueue-only
 *   finishFrames        waits until every frame warped so far is complete; sets warpCalcTime like downloadFrame does */
bool updateFrameDevice(struct OpticalFlowCalc *ofc, unsigned char **devicePlanes);
bool warpFramesToDevice(struct OpticalFlowCalc *ofc, const float blendingScalar, const int frameOutputMode, unsigned char **devicePlanes);
bool finishFrames(struct OpticalFlowCalc *ofc);

/* Page-locked host memory for the images the filter allocates itself (its output pool, reference vf_HopperRender.c:385,
 * :699): downloadFrame then copies straight into the image instead of through a staging buffer. Shaped for
 * mp_image_from_buffer (video/mp_image.h:139-142): allocHostPlanes returns NULL when no page-locked memory is to be had
 * (the caller falls back to mp_image_alloc), freeHostPlanes has the signature of its free callback and may run on any
 * thread. */
unsigned char *allocHostPlanes(size_t bytes);
void freeHostPlanes(void *opaque, unsigned char *data);

#endif /* OPTICALFLOWCALC_H */
