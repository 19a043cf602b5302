/*
 * hr_p010_patch.h — force-included (after this repository's opticalFlowCalc.h) when oracle/build_ref.py compiles the
 * reference's vf_HopperRender.c for the P010 build of the filter harness. It states, as a macro seen by that
 * translation unit only, the change a maintainer makes in the filter for P010:
 * at vf_HopperRender.c:446 set `ofc->pixelFormat = 1` before the call and pass the stride in samples (mpv's
 * mp_image.stride is in bytes, two per P010 sample). Format negotiation (:385, :668) is the harness's business.
 * Test infrastructure only.
 */
#ifndef HR_P010_PATCH_H
#define HR_P010_PATCH_H
#define initOpticalFlowCalc(o, h, s, w) ((o)->pixelFormat = 1, initOpticalFlowCalc((o), (h), (s) / 2, (w)))
#endif
