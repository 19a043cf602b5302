/*
 * hr_staging.h — pageable host planes through a pinned staging ring (host code only).
 *
 * The reference moves frames with blocking clEnqueueWriteBuffer / clEnqueueReadBuffer on whatever memory the filter
 * holds (video/filter/HopperRender/opticalFlowCalc.c:98-100,112-114): mpv's image pool and the decoder hand it
 * malloc'd planes. The CUDA driver moves such memory through a staging buffer of its own with ONE copying thread
 * (measured 12.7 GB/s at 1080p against 42 GB/s from pinned planes). Registering the caller's images is not safe (a
 * registered range that the pool frees and malloc hands out again keeps pointing at the old pages), so the library
 * stages them itself: a ring of pinned chunks per context, filled / drained by a few copying threads while the copy
 * engine moves the chunk before / after. Planes that are pinned already never come here (hr_cuda.cu asks
 * cudaPointerGetAttributes), and nothing about the results changes — the bytes only take another road.
 *
 * Knobs (environment, read at hr_create): HR_STAGE_THREADS (copying threads beside the caller's, default 4;
 * 0: leave pageable planes to the driver as before), HR_STAGE_CHUNK_KB (chunk size, default 1024).
 */
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__) && !defined(HR_STAGE_PLAIN_MEMCPY)
#include <emmintrin.h>
#define HR_STAGE_STREAMING 1
#endif

#define HR_STAGE_SLOTS 8

/* The chunks of one transfer: full-size ones, and short ones at the end that nothing overlaps — the first chunk of an
 * upload (the copy engine has nothing to move until the crew has filled it) and the last chunk of a download (the crew
 * has to drain it after the copy engine has gone quiet): chunk / 8, / 4, / 2 before (after) the full-size ones. */
/* One piece of a transfer made of several: `bytes` bytes between host and dev, contiguous (blockBytes == 0) or blocks of
 * blockBytes every `pitch` bytes on both sides. */
struct HrStageSeg {
    uint8_t *host, *dev;
    size_t bytes, blockBytes, pitch;
};

struct HrStagePlan {
    std::vector<size_t> off, len;
    size_t n = 0;
    /* transfers of several segments only (build_segments): the segment of every chunk and where in it the chunk starts */
    std::vector<HrStageSeg> segs;
    std::vector<int> seg;
    std::vector<size_t> segOff;
    void build(size_t bytes, size_t chunk, bool shortFirst) {
        segs.clear();
        len.clear();
        size_t left = bytes, c = chunk / 8 < 65536 ? (chunk < 65536 ? chunk : 65536) : chunk / 8;
        while (left) {
            const size_t l = left < c ? left : c;
            len.push_back(l);
            left -= l;
            if (c < chunk) c = c * 2 < chunk ? c * 2 : chunk;
        }
        if (!shortFirst) std::reverse(len.begin(), len.end());
        n = len.size();
        off.resize(n);
        size_t o = 0;
        for (size_t i = 0; i < n; ++i) {
            off[i] = o;
            o += len[i];
        }
    }
    /* The same for a transfer made of nBlocks blocks of blockBytes each (rows of a pitched picture, hr_cuda.cu: the
     * lattice rows of a frame go first): every chunk is a whole number of blocks, so that the copy engine can move it
     * with one pitched copy. blockBytes <= chunk. */
    void build_blocks(size_t nBlocks, size_t blockBytes, size_t chunk, bool shortFirst) {
        segs.clear();
        len.clear();
        const size_t full = chunk / blockBytes;
        size_t left = nBlocks, c = full / 8 ? full / 8 : 1;
        while (left) {
            const size_t l = left < c ? left : c;
            len.push_back(l * blockBytes);
            left -= l;
            if (c < full) c = c * 2 < full ? c * 2 : full;
        }
        if (!shortFirst) std::reverse(len.begin(), len.end());
        n = len.size();
        off.resize(n);
        size_t o = 0;
        for (size_t i = 0; i < n; ++i) {
            off[i] = o;
            o += len[i];
        }
    }
    /* An upload of several segments in one go (the lattice rows of a frame, then the rows between them): the ring keeps
     * flowing across the seams, only the very first chunks are short. A chunk never spans two segments; in a segment of
     * blocks it is a whole number of them (blockBytes <= chunk). */
    void build_segments(const HrStageSeg *s, int nSegs, size_t chunk) {
        segs.assign(s, s + nSegs);
        len.clear();
        seg.clear();
        segOff.clear();
        size_t ramp = chunk / 8 < 65536 ? (chunk < 65536 ? chunk : 65536) : chunk / 8; /* bytes, as in build() */
        for (int i = 0; i < nSegs; ++i) {
            const size_t bb = s[i].blockBytes;
            for (size_t done = 0; done < s[i].bytes;) {
                size_t l = ramp;
                if (bb) l = l / bb ? l / bb * bb : bb;
                if (l > s[i].bytes - done) l = s[i].bytes - done;
                len.push_back(l);
                seg.push_back(i);
                segOff.push_back(done);
                done += l;
                if (ramp < chunk) ramp = ramp * 2 < chunk ? ramp * 2 : chunk;
            }
        }
        n = len.size();
        off.resize(n);
        size_t o = 0;
        for (size_t i = 0; i < n; ++i) {
            off[i] = o;
            o += len[i];
        }
    }
};

/* One participant's share of a copy. Neither side is read again by this core (the ring is read by the copy engine, the
 * caller's plane by whoever comes next), so the stores go past the cache (MOVNTDQ: no read-for-ownership of the
 * destination lines, a third less memory traffic than memcpy below glibc's own non-temporal threshold). */
static inline void hr_stage_copy(uint8_t *dst, const uint8_t *src, size_t n) {
#ifdef HR_STAGE_STREAMING
    const size_t head = (16 - ((uintptr_t)dst & 15)) & 15;
    if (n < 256 + head) {
        memcpy(dst, src, n);
        return;
    }
    memcpy(dst, src, head);
    dst += head, src += head, n -= head;
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_loadu_si128((const __m128i *)(src + i)), b = _mm_loadu_si128((const __m128i *)(src + i + 16));
        const __m128i c = _mm_loadu_si128((const __m128i *)(src + i + 32)), d = _mm_loadu_si128((const __m128i *)(src + i + 48));
        _mm_stream_si128((__m128i *)(dst + i), a);
        _mm_stream_si128((__m128i *)(dst + i + 16), b);
        _mm_stream_si128((__m128i *)(dst + i + 32), c);
        _mm_stream_si128((__m128i *)(dst + i + 48), d);
    }
    _mm_sfence();
    memcpy(dst + i, src + i, n - i);
#else
    memcpy(dst, src, n);
#endif
}

/* The copying threads. A transfer is a run of chunks between one host range and the slots of the ring; chunk c may be
 * copied once the coordinator (the calling thread, which also drives the copy engine) has released it — its slot is
 * free (host -> device) or its bytes have arrived (device -> host). Every worker copies its share of every chunk, in
 * order, and publishes how far it has come; nobody waits at a barrier between chunks, so the crew keeps copying while
 * the coordinator is inside a CUDA call. Workers spin for a short while after a transfer — the next one follows within
 * microseconds in a running stream — and sleep on a condition variable when the stream pauses. */
class HrCopyCrew {
public:
    explicit HrCopyCrew(int nThreads) : n_(nThreads < 1 ? 1 : nThreads > 16 ? 16 : nThreads) {
        for (int i = 0; i < n_; ++i) progress_[i].v.store(0, std::memory_order_relaxed);
        for (int i = 0; i < n_; ++i) workers_.emplace_back([this, i] { run(i); });
    }
    ~HrCopyCrew() {
        {
            std::lock_guard<std::mutex> lk(m_);
            quit_.store(true, std::memory_order_release);
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    HrCopyCrew(const HrCopyCrew &) = delete;
    HrCopyCrew &operator=(const HrCopyCrew &) = delete;
    int threads() const { return n_; }

    /* toRing: chunk c of the plan, host + off[c] -> ring slot (firstSlot + c) % HR_STAGE_SLOTS; otherwise the other way
     * round. The plan stays the caller's and must not change before finish(). With blockBytes / hostPitch the host side is
     * a pitched picture (blocks of blockBytes every hostPitch bytes) and the ring holds the blocks back to back. */
    void begin(bool toRing, uint8_t *host, uint8_t *ring, size_t chunk, unsigned firstSlot, const HrStagePlan *plan, size_t blockBytes = 0,
               size_t hostPitch = 0) {
        toRing_ = toRing;
        host_ = host;
        blockBytes_ = blockBytes; /* 0: the host range is contiguous; otherwise byte x of the transfer lives at */
        hostPitch_ = hostPitch;   /* host + (x / blockBytes) * hostPitch + x % blockBytes                        */
        ring_ = ring;
        chunk_ = chunk;
        firstSlot_ = firstSlot;
        plan_ = plan;
        nch_ = plan->n;
        released_.store(0, std::memory_order_relaxed);
        for (int i = 0; i < n_; ++i) progress_[i].v.store(0, std::memory_order_relaxed);
        gen_.fetch_add(1, std::memory_order_release);
        if (sleepers_.load(std::memory_order_acquire) > 0) {
            { std::lock_guard<std::mutex> lk(m_); }
            cv_.notify_all();
        }
    }
    void release(size_t upTo) { released_.store(upTo, std::memory_order_release); }
    bool chunk_done(size_t c) const {
        for (int i = 0; i < n_; ++i)
            if (progress_[i].v.load(std::memory_order_acquire) <= c) return false;
        return true;
    }
    void wait_chunk(size_t c) const {
        while (!chunk_done(c)) relax();
    }
    /* every transfer ends here, also a failed one: the workers must be through before the next begin */
    void finish() {
        release(nch_);
        if (nch_) wait_chunk(nch_ - 1);
    }
    static void relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#elif defined(__aarch64__)
        asm volatile("yield");
#endif
    }

private:
    void run(int id) {
        uint64_t seen = 0;
        for (;;) {
            int spins = 0;
            while (gen_.load(std::memory_order_acquire) == seen) {
                if (++spins < 200000) {
                    relax();
                    continue;
                }
                std::unique_lock<std::mutex> lk(m_);
                sleepers_.fetch_add(1, std::memory_order_acq_rel);
                cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
                sleepers_.fetch_sub(1, std::memory_order_acq_rel);
                spins = 0;
            }
            seen = gen_.load(std::memory_order_acquire);
            if (quit_.load(std::memory_order_acquire)) return;
            const size_t nch = nch_;
            for (size_t c = 0; c < nch; ++c) {
                while (released_.load(std::memory_order_acquire) <= c) relax();
                const size_t o = plan_->off[c], len = plan_->len[c];
                const size_t per = (((len + n_ - 1) / n_) + 4095) & ~(size_t)4095, so = (size_t)id * per;
                if (so < len) {
                    uint8_t *r = ring_ + (size_t)((firstSlot_ + c) % HR_STAGE_SLOTS) * chunk_ + so;
                    size_t m = len - so < per ? len - so : per;
                    /* where the chunk's bytes live on the host: one range / one pitched picture for the whole transfer,
                     * or the chunk's segment */
                    uint8_t *hb = host_;
                    size_t x = o + so, bb = blockBytes_, pt = hostPitch_;
                    if (!plan_->segs.empty()) {
                        const HrStageSeg &sg = plan_->segs[plan_->seg[c]];
                        hb = sg.host, x = plan_->segOff[c] + so, bb = sg.blockBytes, pt = sg.pitch;
                    }
                    if (!bb) {
                        uint8_t *h = hb + x;
                        if (toRing_) hr_stage_copy(r, h, m);
                        else hr_stage_copy(h, r, m);
                    } else {
                        while (m) { /* block by block */
                            const size_t w = x % bb, l = bb - w < m ? bb - w : m;
                            uint8_t *h = hb + (x / bb) * pt + w;
                            if (toRing_) hr_stage_copy(r, h, l);
                            else hr_stage_copy(h, r, l);
                            r += l, x += l, m -= l;
                        }
                    }
                }
                progress_[id].v.store(c + 1, std::memory_order_release);
            }
        }
    }

    struct alignas(64) Progress {
        std::atomic<size_t> v;
    };
    const int n_;
    std::vector<std::thread> workers_;
    std::atomic<uint64_t> gen_{0};
    std::atomic<size_t> released_{0};
    std::atomic<int> sleepers_{0};
    std::atomic<bool> quit_{false};
    std::mutex m_;
    std::condition_variable cv_;
    Progress progress_[16];
    bool toRing_ = true;
    uint8_t *host_ = nullptr, *ring_ = nullptr;
    size_t chunk_ = 0, nch_ = 0, blockBytes_ = 0, hostPitch_ = 0;
    const HrStagePlan *plan_ = nullptr;
    unsigned firstSlot_ = 0;
};
