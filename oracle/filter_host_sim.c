/*
 * filter_host_sim.c — drives the REFERENCE FILTER (video/filter/HopperRender/vf_HopperRender.c, compiled
 * unmodified from /root/reference by oracle/build_ref.py) on top of THIS repository's optical-flow-calc
 * layer (mpv-frame-interpolator_b200/mpv/video/filter/HopperRender/opticalFlowCalc.c -> the CUDA C ABI),
 * outside an mpv build. It implements the slice of mpv's filter runtime declared in
 * mpv_shim/hr_mpv_shim.h and a small push/run/pop API for the tests (tests/test_gpu_filter_host.py):
 * that is the drop-in claim made executable — the reference's own filter code, its six calls into
 * opticalFlowCalc.h, our library underneath. Test infrastructure only (oracle/).
 */
#include "mpv_shim/hr_mpv_shim.h"

extern const struct mp_user_filter_entry vf_HopperRender; /* vf_HopperRender.c:719-720 */

/* ---- talloc / images -------------------------------------------------------------------------- */
void talloc_free(void *p) { free(p); }

/* bytes per sample of the images the harness hands to the filter: 1 = NV12 (what the reference negotiates), 2 = the
 * P010 build of the harness (oracle/build_ref.py: the same unmodified filter source, compiled with the two changes a
 * maintainer makes for P010 expressed as a macro — pixelFormat = 1 and the stride in samples instead of bytes) */
#ifndef HR_SIM_BPS
#define HR_SIM_BPS 1
#endif
int hr_sim_bytes_per_sample(void) { return HR_SIM_BPS; }

/* device memory for the IMGFMT_CUDA images of the zero-copy build (HR_SIM_CUDA: linked against libcudart) */
#ifdef HR_SIM_CUDA
extern int cudaMalloc(void **p, size_t n);
extern int cudaFree(void *p);
extern int cudaMemset(void *p, int v, size_t n);
#else
static int cudaMalloc(void **p, size_t n) { (void)n; *p = NULL; return 1; }
static int cudaFree(void *p) { (void)p; return 0; }
static int cudaMemset(void *p, int v, size_t n) { (void)p; (void)v; (void)n; return 1; }
#endif

static int fmt_bps(int fmt) { return (fmt == IMGFMT_P010 || HR_SIM_BPS == 2) ? 2 : 1; }

static struct mp_image *image_alloc(int w, int h, int fmt) {
    struct mp_image *img = calloc(1, sizeof(*img));
    const int stride = (w * fmt_bps(fmt) + 63) & ~63; /* mpv aligns strides to 64 bytes (video/mp_image.h:35) */
    img->w = w;
    img->h = h;
    img->imgfmt = fmt;
    img->storage = malloc((size_t)stride * h * 3 / 2 + 64);
    img->planes[0] = img->storage;
    img->planes[1] = img->storage + (size_t)stride * h;
    img->stride[0] = img->stride[1] = stride;
    img->refcount = malloc(sizeof(int));
    *img->refcount = 1;
    return img;
}
struct mp_image *mp_image_new_ref(struct mp_image *img) {
    struct mp_image *r = malloc(sizeof(*r));
    *r = *img;
    ++*img->refcount;
    return r;
}
void mp_image_unrefp(struct mp_image **p) {
    struct mp_image *img = p ? *p : NULL;
    if (!img) return;
    if (--*img->refcount == 0) {
        free(img->storage);
        if (img->bufferFree) img->bufferFree(img->bufferOpaque, img->buffer);
        if (img->deviceStorage) cudaFree(img->deviceStorage);
        free(img->refcount);
    }
    free(img);
    *p = NULL;
}
void mp_image_copy_attributes(struct mp_image *dst, struct mp_image *src) {
    dst->pts = src->pts;
    dst->nominal_fps = src->nominal_fps;
}
/* NV12 / P010 layout as mp_image_layout() computes it (video/mp_image.c:59-98): strides aligned to stride_align, the
 * chroma plane behind the luma plane, plane offsets aligned likewise */
static int plane_layout(int fmt, int w, int h, int align, int *stride, int *uvOffset) {
    if (w < 1 || h < 1 || align < 1) return -1;
    *stride = (w * fmt_bps(fmt) + align - 1) / align * align;
    *uvOffset = (*stride * h + align - 1) / align * align;
    return *uvOffset + (*stride * ((h + 1) / 2) + align - 1) / align * align;
}
int mp_image_get_alloc_size(int imgfmt, int w, int h, int stride_align) {
    int stride, uvOffset;
    return plane_layout(imgfmt, w, h, stride_align, &stride, &uvOffset);
}
struct mp_image *mp_image_from_buffer(int imgfmt, int w, int h, int stride_align, unsigned char *buffer, int buffer_size, void *free_opaque,
                                      void (*free_fn)(void *opaque, unsigned char *data)) {
    int stride, uvOffset;
    const int size = plane_layout(imgfmt, w, h, stride_align, &stride, &uvOffset);
    const int shift = (int)(((size_t)buffer + stride_align - 1) / stride_align * stride_align - (size_t)buffer);
    if (size < 0 || size > buffer_size || buffer_size - size < shift) return NULL;
    struct mp_image *img = calloc(1, sizeof(*img));
    img->w = w;
    img->h = h;
    img->imgfmt = imgfmt;
    img->planes[0] = buffer + shift;
    img->planes[1] = buffer + shift + uvOffset;
    img->stride[0] = img->stride[1] = stride;
    img->buffer = buffer;
    img->bufferOpaque = free_opaque;
    img->bufferFree = free_fn;
    img->refcount = malloc(sizeof(int));
    *img->refcount = 1;
    return img;
}
struct mp_image *mp_image_alloc(int fmt, int w, int h) { return image_alloc(w, h, fmt); }

/* The pool keeps the images it has handed out and hands an image out again once every other reference to it is gone
 * (video/mp_image_pool.c:117-170, 204-237); new ones come from the allocator callback if one is set (:247-252). */
#define SIM_POOL_MAX 64
struct mp_image_pool {
    struct mp_image *images[SIM_POOL_MAX];
    int n, allocated;
    mp_image_allocator allocator;
    void *allocatorData;
};
static int g_poolImagesAllocated;
struct mp_image_pool *mp_image_pool_new(void *tparent) {
    (void)tparent;
    return calloc(1, sizeof(struct mp_image_pool));
}
void mp_image_pool_set_allocator(struct mp_image_pool *pool, mp_image_allocator cb, void *cb_data) {
    pool->allocator = cb;
    pool->allocatorData = cb_data;
}
struct mp_image *mp_image_pool_get(struct mp_image_pool *pool, int fmt, int w, int h) {
    for (int i = 0; i < pool->n; ++i) {
        struct mp_image *img = pool->images[i];
        if (*img->refcount == 1 && img->imgfmt == fmt && img->w == w && img->h == h) return mp_image_new_ref(img);
    }
    if (pool->n == SIM_POOL_MAX) return NULL;
    struct mp_image *img = pool->allocator ? pool->allocator(pool->allocatorData, fmt, w, h) : image_alloc(w, h, fmt);
    if (!img) return NULL;
    pool->images[pool->n++] = img;
    ++g_poolImagesAllocated;
    return mp_image_new_ref(img);
}
void mp_image_pool_clear(struct mp_image_pool *pool) {
    for (int i = 0; i < pool->n; ++i) mp_image_unrefp(&pool->images[i]); /* images still referenced outside live on */
    pool->n = 0;
}
int hr_sim_pool_images_allocated(void) { return g_poolImagesAllocated; }

/* ---- the slice of libavutil / mpv hardware-frame plumbing the IMGFMT_CUDA patch uses --------------------------- */
static struct AVBufferRef g_deviceRef;          /* "the CUDA device": one per process here                 */
struct SimFramesCtx {
    struct AVBufferRef ref;                      /* ref.data -> ctx                                         */
    AVHWFramesContext ctx;
};
static struct SimFramesCtx g_inputFrames;        /* the frames context the harness's input images claim      */
AVFrame *av_frame_alloc(void) { return calloc(1, sizeof(AVFrame)); }
void av_frame_free(AVFrame **frame) {
    if (frame && *frame) free(*frame);
    if (frame) *frame = NULL;
}
void av_buffer_unref(struct AVBufferRef **buf) {
    if (buf && *buf && *buf != &g_inputFrames.ref) free(*buf); /* a SimFramesCtx starts with its ref */
    if (buf) *buf = NULL;
}
bool mp_update_av_hw_frames_pool(struct AVBufferRef **hw_frames_ctx, struct AVBufferRef *hw_device_ctx, int imgfmt, int sw_imgfmt, int w, int h,
                                 bool disable_multiplane) {
    (void)disable_multiplane;
    if (imgfmt != IMGFMT_CUDA || !hw_device_ctx || w < 1 || h < 1) return false;
    if (*hw_frames_ctx) {
        AVHWFramesContext *c = (void *)(*hw_frames_ctx)->data;
        if (c->device_ref != hw_device_ctx || c->sw_format != sw_imgfmt || c->width != w || c->height != h) av_buffer_unref(hw_frames_ctx);
    }
    if (!*hw_frames_ctx) {
        struct SimFramesCtx *n = calloc(1, sizeof(*n));
        n->ref.data = (unsigned char *)&n->ctx;
        n->ctx.device_ref = hw_device_ctx;
        n->ctx.sw_format = sw_imgfmt;
        n->ctx.width = w;
        n->ctx.height = h;
        *hw_frames_ctx = &n->ref;
    }
    return true;
}
static int g_deviceImagesAllocated;
/* libavutil's CUDA frames: one allocation, the chroma plane behind the luma plane, pitch aligned to 256 bytes */
int av_hwframe_get_buffer(struct AVBufferRef *hwframe_ctx, AVFrame *frame, int flags) {
    (void)flags;
    AVHWFramesContext *c = (void *)hwframe_ctx->data;
    const int bps = c->sw_format == IMGFMT_P010 ? 2 : 1;
    const int pitch = (c->width * bps + 255) & ~255;
    void *d = NULL;
    const size_t n = (size_t)pitch * c->height * 3 / 2;
    if (cudaMalloc(&d, n) != 0 || !d) return -1;
    cudaMemset(d, 0, n);
    frame->data[0] = d;
    frame->data[1] = (unsigned char *)d + (size_t)pitch * c->height;
    frame->linesize[0] = frame->linesize[1] = pitch;
    frame->width = c->width;
    frame->height = c->height;
    frame->hw_frames_ctx = hwframe_ctx;
    ++g_deviceImagesAllocated;
    return 0;
}
struct mp_image *mp_image_from_av_frame(AVFrame *src) {
    AVHWFramesContext *c = (void *)src->hw_frames_ctx->data;
    struct mp_image *img = calloc(1, sizeof(*img));
    img->w = src->width;
    img->h = src->height;
    img->imgfmt = IMGFMT_CUDA;
    img->params.hw_subfmt = c->sw_format;
    img->hwctx = src->hw_frames_ctx;
    img->planes[0] = src->data[0];
    img->planes[1] = src->data[1];
    img->stride[0] = src->linesize[0];
    img->stride[1] = src->linesize[1];
    img->deviceStorage = src->data[0];
    img->refcount = malloc(sizeof(int));
    *img->refcount = 1;
    return img;
}
int hr_sim_device_images_allocated(void) { return g_deviceImagesAllocated; }

bool mp_frame_is_signaling(struct mp_frame frame) { return frame.type == MP_FRAME_EOF; }

/* ---- pins: one slot each -------------------------------------------------------------------------- */
struct mp_pin {
    struct mp_frame slot;
    bool has;
    int role; /* 0 filter input, 1 filter output, 2 autoconvert input, 3 autoconvert output */
    struct mp_filter_sim *sim;
};
#define SIM_MAX_OUT 64
struct mp_filter_sim {
    struct mp_filter *f;
    struct mp_pin in, out, conv_in, conv_out;
    struct mp_pin *fpins[2], *cpins[2];
    struct mp_filter conv_filter;
    struct mp_autoconvert conv;
    struct mp_stream_info info;
    struct mp_image *outputs[SIM_MAX_OUT];
    int n_out;
    bool progress, failed;
    void *opts;
};
static struct mp_filter_sim *g_sim; /* the instance under construction (mp_filter_create has no user pointer) */

struct mp_filter *mp_filter_create(struct mp_filter *parent, const struct mp_filter_info *info) {
    (void)parent;
    struct mp_filter *f = calloc(1, sizeof(*f));
    f->info = info;
    f->priv = calloc(1, info->priv_size);
    f->sim = g_sim;
    g_sim->f = f;
    g_sim->fpins[0] = &g_sim->in;
    g_sim->fpins[1] = &g_sim->out;
    f->pins = f->ppins = g_sim->fpins;
    g_sim->in.role = 0;
    g_sim->out.role = 1;
    g_sim->conv_in.role = 2;
    g_sim->conv_out.role = 3;
    g_sim->in.sim = g_sim->out.sim = g_sim->conv_in.sim = g_sim->conv_out.sim = g_sim;
    return f;
}
struct mp_pin *mp_filter_add_pin(struct mp_filter *f, enum mp_pin_dir dir, const char *name) {
    (void)name;
    f->num_pins++;
    return dir == MP_PIN_IN ? &f->sim->in : &f->sim->out;
}
void mp_filter_internal_mark_progress(struct mp_filter *f) { f->sim->progress = true; }
void mp_filter_internal_mark_failed(struct mp_filter *f) { f->sim->failed = true; }

bool mp_pin_in_needs_data(struct mp_pin *p) {
    if (p->role == 1) return p->sim->n_out < SIM_MAX_OUT; /* the harness drains the output list */
    if (p->role == 2) return !p->sim->conv_out.has;       /* pass-through autoconvert */
    return !p->has;
}
bool mp_pin_can_transfer_data(struct mp_pin *dst, struct mp_pin *src) { return src->has && mp_pin_in_needs_data(dst); }
struct mp_frame mp_pin_out_read(struct mp_pin *p) {
    struct mp_frame fr = p->has ? p->slot : (struct mp_frame){MP_FRAME_NONE, NULL};
    p->has = false;
    return fr;
}
bool mp_pin_in_write(struct mp_pin *p, struct mp_frame frame) {
    if (frame.type == MP_FRAME_NONE) return true;
    struct mp_filter_sim *s = p->sim;
    if (p->role == 1) {
        if (frame.type == MP_FRAME_VIDEO && s->n_out < SIM_MAX_OUT) s->outputs[s->n_out++] = frame.data;
        return true;
    }
    if (p->role == 2) { /* autoconvert: the harness only feeds NV12, so conversion is the identity */
        s->conv_out.slot = frame;
        s->conv_out.has = true;
        s->progress = true;
        return true;
    }
    p->slot = frame;
    p->has = true;
    return true;
}

static double sim_display_fps(struct mp_stream_info *i) { return i->display_fps; }
struct mp_stream_info *mp_filter_find_stream_info(struct mp_filter *f) { return &f->sim->info; }
struct mp_autoconvert *mp_autoconvert_create(struct mp_filter *parent) {
    struct mp_filter_sim *s = parent->sim;
    s->cpins[0] = &s->conv_in;
    s->cpins[1] = &s->conv_out;
    s->conv_filter.pins = s->conv_filter.ppins = s->cpins;
    s->conv_filter.sim = s;
    s->conv.f = &s->conv_filter;
    return &s->conv;
}
void mp_autoconvert_add_imgfmt(struct mp_autoconvert *c, int imgfmt, int subfmt) {
    (void)c;
    (void)imgfmt;
    (void)subfmt;
}

/* ---- harness API ------------------------------------------------------------------------------------ */
struct mp_filter_sim *hr_sim_create(int frameOutput, double displayFps) {
    struct mp_filter_sim *s = calloc(1, sizeof(*s));
    s->info.get_display_fps = sim_display_fps;
    s->info.display_fps = displayFps; /* what --vo=null --vo-null-fps reports, f_output_chain.c:357-364 */
    /* options block as filters/user_filters.c:174-190 would fill it from vf_opts_fields */
    const m_option_t *o = vf_HopperRender.desc.options;
    s->opts = calloc(1, 64);
    *(int *)((char *)s->opts + o[0].offset) = frameOutput < o[0].min || frameOutput > o[0].max ? o[0].defval : frameOutput;
    g_sim = s;
    struct mp_filter *f = vf_HopperRender.create(NULL, s->opts);
    g_sim = NULL;
    if (!f) {
        free(s);
        return NULL;
    }
    return s;
}
const char *hr_sim_filter_name(void) { return vf_HopperRender.desc.name; }

/* feed one NV12 (P010 build: P010) source frame (tightly packed planes of w x h samples) and run the filter until it stalls;
 * returns the number of output frames waiting, or -1 if the filter marked itself failed */
static int sim_run(struct mp_filter_sim *s, struct mp_image *img) {
    s->in.slot = MAKE_FRAME(MP_FRAME_VIDEO, img);
    s->in.has = true;
    do { /* filters/filter.c:211-263: run process() while somebody reports progress */
        s->progress = false;
        s->f->info->process(s->f);
    } while (s->progress && !s->failed);
    return s->failed ? -1 : s->n_out;
}
/* the same for a frame that lives in device memory (IMGFMT_CUDA, as a CUDA decoder delivers it): dY / dUV are device
 * pointers the caller keeps valid while the filter holds the frame (two source frames), pitch in bytes */
int hr_sim_push_device(struct mp_filter_sim *s, void *dY, void *dUV, int w, int h, int pitchBytes, int swFormat, double pts, double nominalFps) {
    struct mp_image *img = calloc(1, sizeof(*img));
    g_inputFrames.ref.data = (unsigned char *)&g_inputFrames.ctx;
    g_inputFrames.ctx.device_ref = &g_deviceRef;
    g_inputFrames.ctx.sw_format = swFormat;
    g_inputFrames.ctx.width = w;
    g_inputFrames.ctx.height = h;
    img->w = w;
    img->h = h;
    img->imgfmt = IMGFMT_CUDA;
    img->params.hw_subfmt = swFormat;
    img->hwctx = &g_inputFrames.ref;
    img->planes[0] = dY;
    img->planes[1] = dUV;
    img->stride[0] = img->stride[1] = pitchBytes;
    img->refcount = malloc(sizeof(int));
    *img->refcount = 1;
    img->pts = pts;
    img->nominal_fps = nominalFps;
    return sim_run(s, img);
}
/* oldest waiting output frame as it is: format, device (or host) plane pointers, pitch. The frame stays alive until the
 * next call of this function or hr_sim_destroy. Returns 0, or 1 when none is waiting. */
int hr_sim_pop_image(struct mp_filter_sim *s, int *imgfmt, int *subfmt, void **p0, void **p1, int *pitch, double *pts) {
    static struct mp_image *held;
    if (held) mp_image_unrefp(&held);
    if (!s || s->n_out == 0) return 1;
    struct mp_image *img = s->outputs[0];
    memmove(s->outputs, s->outputs + 1, sizeof(s->outputs[0]) * (size_t)(--s->n_out));
    if (imgfmt) *imgfmt = img->imgfmt;
    if (subfmt) *subfmt = img->params.hw_subfmt;
    if (p0) *p0 = img->planes[0];
    if (p1) *p1 = img->planes[1];
    if (pitch) *pitch = img->stride[0];
    if (pts) *pts = img->pts;
    held = img;
    return 0;
}
void hr_sim_command_text(struct mp_filter_sim *s, const char *cmd, const char *arg) {
    struct mp_filter_command c = {.type = MP_FILTER_COMMAND_TEXT, .cmd = cmd, .arg = arg, .speed = 1.0};
    s->f->info->command(s->f, &c);
}
/* host frame of an explicit format (the patched filter negotiates NV12 or P010 by the image's format) */
int hr_sim_push_fmt(struct mp_filter_sim *s, const unsigned char *y, const unsigned char *uv, int w, int h, int fmt, double pts, double nominalFps) {
    struct mp_image *img = image_alloc(w, h, fmt);
    const size_t rowBytes = (size_t)w * fmt_bps(fmt);
    for (int r = 0; r < h; ++r) memcpy(img->planes[0] + (size_t)r * img->stride[0], y + (size_t)r * rowBytes, rowBytes);
    for (int r = 0; r < h / 2; ++r) memcpy(img->planes[1] + (size_t)r * img->stride[1], uv + (size_t)r * rowBytes, rowBytes);
    img->pts = pts;
    img->nominal_fps = nominalFps;
    return sim_run(s, img);
}
int hr_sim_push(struct mp_filter_sim *s, const unsigned char *y, const unsigned char *uv, int w, int h, double pts, double nominalFps) {
    struct mp_image *img = image_alloc(w, h, HR_SIM_BPS == 2 ? IMGFMT_P010 : IMGFMT_NV12);
    const size_t rowBytes = (size_t)w * HR_SIM_BPS;
    for (int r = 0; r < h; ++r) memcpy(img->planes[0] + (size_t)r * img->stride[0], y + (size_t)r * rowBytes, rowBytes);
    for (int r = 0; r < h / 2; ++r) memcpy(img->planes[1] + (size_t)r * img->stride[1], uv + (size_t)r * rowBytes, rowBytes);
    img->pts = pts;
    img->nominal_fps = nominalFps;
    return sim_run(s, img);
}
/* oldest waiting output frame -> tightly packed planes; returns 0, or 1 when none is waiting */
int hr_sim_pop(struct mp_filter_sim *s, unsigned char *y, unsigned char *uv, double *pts, int *stride) {
    if (s->n_out == 0) return 1;
    struct mp_image *img = s->outputs[0];
    memmove(s->outputs, s->outputs + 1, sizeof(s->outputs[0]) * (size_t)(--s->n_out));
    const size_t rowBytes = (size_t)img->w * fmt_bps(img->imgfmt);
    for (int r = 0; r < img->h; ++r) memcpy(y + (size_t)r * rowBytes, img->planes[0] + (size_t)r * img->stride[0], rowBytes);
    for (int r = 0; r < img->h / 2; ++r) memcpy(uv + (size_t)r * rowBytes, img->planes[1] + (size_t)r * img->stride[1], rowBytes);
    if (pts) *pts = img->pts;
    if (stride) *stride = img->stride[0];
    mp_image_unrefp(&img);
    return 0;
}
void hr_sim_command_speed(struct mp_filter_sim *s, double speed) {
    struct mp_filter_command c = {.type = MP_FILTER_COMMAND_TEXT, .speed = speed};
    s->f->info->command(s->f, &c);
}
void hr_sim_reset(struct mp_filter_sim *s) { s->f->info->reset(s->f); }
void hr_sim_destroy(struct mp_filter_sim *s) {
    s->f->info->destroy(s->f);
    free(s);
}
