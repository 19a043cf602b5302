/*
 * hr_mpv_shim.h — the slice of mpv's filter API that video/filter/HopperRender/vf_HopperRender.c uses,
 * re-declared from scratch so that the REFERENCE FILTER can be compiled UNMODIFIED outside an mpv
 * build (this image has no meson, ffmpeg, libplacebo ...). Test infrastructure only (oracle/).
 *
 * Names and the fields the filter touches follow mpv (filters/filter.h, filters/filter_internal.h,
 * filters/frame.h, filters/f_autoconvert.h, filters/user_filters.h, video/mp_image.h,
 * video/mp_image_pool.h, options/m_option.h); everything else is reduced to what
 * oracle/filter_host_sim.c needs to drive the filter: single-slot pins, a pass-through autoconvert,
 * malloc-backed images.
 */
#ifndef HR_MPV_SHIM_H
#define HR_MPV_SHIM_H

/* The reference filter includes its own config.h (same directory) just before the first mpv header, so
 * this is the place where the harness overrides two of its knobs: no GTK applet (it forks a Python
 * script and blocks on a FIFO, vf_HopperRender.c:633-656) and no timing-driven search-radius drift
 * (vf_HopperRender.c:326-345) — otherwise no two runs would produce the same frames. */
#undef INC_APP_IND
#define INC_APP_IND 0
#undef AUTO_SEARCH_RADIUS_ADJUST
#define AUTO_SEARCH_RADIUS_ADJUST 0

#include <math.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ---- logging, talloc -------------------------------------------------------------------------- */
#define MP_ERR(f, ...) do { (void)(f); fprintf(stderr, "[vf_HopperRender] error: " __VA_ARGS__); } while (0)
#define MP_WARN(f, ...) do { (void)(f); fprintf(stderr, "[vf_HopperRender] warning: " __VA_ARGS__); } while (0)
void talloc_free(void *p);
#define talloc_steal(ctx, p) (p)

/* ---- images ------------------------------------------------------------------------------------ */
#define IMGFMT_NV12 1
#define IMGFMT_P010 2
#define IMGFMT_CUDA 3 /* planes[] are CUDA device pointers, params.hw_subfmt the layout (video/img_format.h:274) */
struct AVBufferRef {
    unsigned char *data;
};
typedef struct AVHWFramesContext { /* libavutil/hwcontext.h: the fields the patched filter reads */
    struct AVBufferRef *device_ref;
    int sw_format, width, height;
} AVHWFramesContext;
typedef struct AVFrame {
    unsigned char *data[4];
    int linesize[4];
    int width, height;
    struct AVBufferRef *hw_frames_ctx;
} AVFrame;
struct mp_image_params {
    int hw_subfmt;
};
struct mp_image {
    int w, h;
    int imgfmt;
    struct mp_image_params params;
    struct AVBufferRef *hwctx; /* IMGFMT_CUDA: the frames context the image belongs to */
    unsigned char *planes[4];
    int stride[4];
    double pts;
    double nominal_fps;
    int *refcount;          /* shared between references                      */
    unsigned char *storage; /* freed with the last reference (host images)    */
    void *deviceStorage;    /* cudaFree()d with the last reference            */
    unsigned char *buffer;  /* mp_image_from_buffer: handed to bufferFree with the last reference */
    void *bufferOpaque;
    void (*bufferFree)(void *opaque, unsigned char *data);
};
#define MP_IMAGE_BYTE_ALIGN 64 /* video/mp_image.h:35 */
int mp_image_get_alloc_size(int imgfmt, int w, int h, int stride_align); /* video/mp_image.h:138 */
struct mp_image *mp_image_from_buffer(int imgfmt, int w, int h, int stride_align, unsigned char *buffer, int buffer_size, void *free_opaque,
                                      void (*free)(void *opaque, unsigned char *data)); /* video/mp_image.h:139-142 */
struct mp_image *mp_image_alloc(int fmt, int w, int h);                  /* video/mp_image.h:144 */
AVFrame *av_frame_alloc(void);
void av_frame_free(AVFrame **frame);
int av_hwframe_get_buffer(struct AVBufferRef *hwframe_ctx, AVFrame *frame, int flags);
void av_buffer_unref(struct AVBufferRef **buf);
struct mp_image *mp_image_from_av_frame(AVFrame *src);
bool mp_update_av_hw_frames_pool(struct AVBufferRef **hw_frames_ctx, struct AVBufferRef *hw_device_ctx, int imgfmt, int sw_imgfmt, int w, int h,
                                 bool disable_multiplane); /* video/mp_image_pool.h:37 */
struct mp_image *mp_image_new_ref(struct mp_image *img);
void mp_image_unrefp(struct mp_image **img);
void mp_image_copy_attributes(struct mp_image *dst, struct mp_image *src);
struct mp_image_pool;
struct mp_image_pool *mp_image_pool_new(void *tparent);
struct mp_image *mp_image_pool_get(struct mp_image_pool *pool, int fmt, int w, int h);
void mp_image_pool_clear(struct mp_image_pool *pool);
typedef struct mp_image *(*mp_image_allocator)(void *data, int fmt, int w, int h);                    /* video/mp_image_pool.h:20 */
void mp_image_pool_set_allocator(struct mp_image_pool *pool, mp_image_allocator cb, void *cb_data); /* video/mp_image_pool.h:21-22 */

/* ---- frames ------------------------------------------------------------------------------------ */
enum mp_frame_type { MP_FRAME_NONE = 0, MP_FRAME_VIDEO, MP_FRAME_AUDIO, MP_FRAME_PACKET, MP_FRAME_EOF };
struct mp_frame {
    enum mp_frame_type type;
    void *data;
};
#define MAKE_FRAME(t, d) ((struct mp_frame){(t), (d)})
bool mp_frame_is_signaling(struct mp_frame frame);

/* ---- filters and pins ---------------------------------------------------------------------------- */
enum mp_pin_dir { MP_PIN_INVALID = 0, MP_PIN_IN, MP_PIN_OUT };
struct mp_pin;
struct mp_filter;
enum mp_filter_command_type { MP_FILTER_COMMAND_NONE = 0, MP_FILTER_COMMAND_TEXT };
struct mp_filter_command { /* filters/filter.h:377-392 */
    enum mp_filter_command_type type;
    const char *target, *cmd, *arg;
    void *res;
    double speed;
    bool is_active;
};
struct mp_filter_info {
    const char *name;
    int priv_size;
    void (*process)(struct mp_filter *f);
    void (*reset)(struct mp_filter *f);
    void (*destroy)(struct mp_filter *f);
    bool (*command)(struct mp_filter *f, struct mp_filter_command *cmd);
};
struct mp_filter {
    const struct mp_filter_info *info;
    void *priv;
    struct mp_pin **pins;  /* as seen from outside */
    struct mp_pin **ppins; /* as seen by the filter */
    int num_pins;
    struct mp_filter_sim *sim;
};
struct mp_filter *mp_filter_create(struct mp_filter *parent, const struct mp_filter_info *info);
struct mp_pin *mp_filter_add_pin(struct mp_filter *f, enum mp_pin_dir dir, const char *name);
void mp_filter_internal_mark_progress(struct mp_filter *f);
void mp_filter_internal_mark_failed(struct mp_filter *f);
bool mp_pin_can_transfer_data(struct mp_pin *dst, struct mp_pin *src);
bool mp_pin_in_needs_data(struct mp_pin *p);
bool mp_pin_in_write(struct mp_pin *p, struct mp_frame frame);
struct mp_frame mp_pin_out_read(struct mp_pin *p);

struct mp_stream_info {
    double (*get_display_fps)(struct mp_stream_info *i);
    double display_fps;
};
struct mp_stream_info *mp_filter_find_stream_info(struct mp_filter *f);

struct mp_autoconvert {
    struct mp_filter *f;
};
struct mp_autoconvert *mp_autoconvert_create(struct mp_filter *parent);
void mp_autoconvert_add_imgfmt(struct mp_autoconvert *c, int imgfmt, int subfmt);

/* ---- options / registration -------------------------------------------------------------------- */
typedef struct m_option {
    const char *name;
    int offset;
    int min, max;
    int defval;
} m_option_t;
#define OPT_INT(field) offsetof(OPT_BASE_STRUCT, field)
#define M_RANGE(a, b) (a), (b)
#define OPTDEF_INT(v) (v)
struct m_obj_desc {
    const char *name;
    const char *description;
    int priv_size;
    const m_option_t *options;
};
struct mp_user_filter_entry {
    struct m_obj_desc desc;
    struct mp_filter *(*create)(struct mp_filter *parent, void *options);
};

#endif
