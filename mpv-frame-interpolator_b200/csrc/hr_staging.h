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
 * Knobs (environment, read at hr_create): HR_STAGE_THREADS (copying threads including the caller's, default 4;
 * 0 or 1: leave pageable planes to the driver as before), HR_STAGE_CHUNK_KB (chunk size, default 512).
 */
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#define HR_STAGE_SLOTS 8

/* memcpy of one range by nThreads participants (the caller is one of them). Workers spin for a short while after a job
 * — the next chunk follows within microseconds — and sleep on a condition variable when the stream pauses. */
class HrCopyCrew {
public:
    explicit HrCopyCrew(int nThreads) : n_(nThreads < 1 ? 1 : nThreads) {
        for (int i = 1; i < n_; ++i) workers_.emplace_back([this, i] { run(i); });
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

    void copy(void *dst, const void *src, size_t bytes) {
        if (n_ == 1 || bytes < (size_t)n_ * 16384) {
            memcpy(dst, src, bytes);
            return;
        }
        dst_ = (uint8_t *)dst;
        src_ = (const uint8_t *)src;
        bytes_ = bytes;
        left_.store(n_ - 1, std::memory_order_relaxed);
        gen_.fetch_add(1, std::memory_order_release);
        if (sleepers_.load(std::memory_order_acquire) > 0) {
            { std::lock_guard<std::mutex> lk(m_); }
            cv_.notify_all();
        }
        slice(0);
        while (left_.load(std::memory_order_acquire) > 0) relax();
    }

private:
    static void relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#elif defined(__aarch64__)
        asm volatile("yield");
#endif
    }
    void slice(int id) {
        const size_t per = (((bytes_ + n_ - 1) / n_) + 4095) & ~(size_t)4095;
        const size_t o = (size_t)id * per;
        if (o < bytes_) memcpy(dst_ + o, src_ + o, bytes_ - o < per ? bytes_ - o : per);
    }
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
            slice(id);
            left_.fetch_sub(1, std::memory_order_release);
        }
    }

    const int n_;
    std::vector<std::thread> workers_;
    std::atomic<uint64_t> gen_{0};
    std::atomic<int> left_{0}, sleepers_{0};
    std::atomic<bool> quit_{false};
    std::mutex m_;
    std::condition_variable cv_;
    uint8_t *dst_ = nullptr;
    const uint8_t *src_ = nullptr;
    size_t bytes_ = 0;
};
