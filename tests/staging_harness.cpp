/* CPU harness for csrc/hr_staging.h (tests/test_staging_cpu.py): the copying threads gather the blocks of a pitched picture into
 * the ring (or scatter them back), a memcpy stands in for the copy engine; every block pattern, thread count and direction
 * must reproduce the picture and leave the bytes between the blocks alone. */
#include "hr_staging.h"
#include <cstdio>
#include <cstdlib>
// emulate staged_h2d with blocks: crew gathers pitched host blocks into ring slots, "copy engine" = memcpy scatter
static int run(size_t nBlocks, size_t blockBytes, size_t pitch, size_t chunk, int threads, bool toRing) {
    std::vector<uint8_t> host(nBlocks * pitch + 64), dev(nBlocks * pitch + 64, 0xEE), ring(HR_STAGE_SLOTS * chunk);
    for (size_t i = 0; i < host.size(); ++i) host[i] = (uint8_t)(i * 2654435761u >> 13);
    if (!toRing) std::swap(host, dev), std::fill(host.begin(), host.end(), 0xEE);
    HrStagePlan plan;
    plan.build_blocks(nBlocks, blockBytes, chunk, toRing);
    HrCopyCrew crew(threads);
    // process in windows of HR_STAGE_SLOTS chunks
    crew.begin(toRing, host.data(), ring.data(), chunk, 3, &plan, blockBytes, pitch);
    size_t released = 0;
    for (size_t c = 0; c < plan.n; ++c) {
        const int slot = (3 + c) % HR_STAGE_SLOTS;
        if (toRing) {
            if (released <= c) crew.release(++released);       // slot free: one chunk at a time (ring reuse is safe then)
            crew.wait_chunk(c);
            for (size_t b = 0; b < plan.len[c] / blockBytes; ++b)
                memcpy(dev.data() + (plan.off[c] / blockBytes + b) * pitch, ring.data() + slot * chunk + b * blockBytes, blockBytes);
        } else {
            for (size_t b = 0; b < plan.len[c] / blockBytes; ++b)
                memcpy(ring.data() + slot * chunk + b * blockBytes, dev.data() + (plan.off[c] / blockBytes + b) * pitch, blockBytes);
            crew.release(c + 1);
            crew.wait_chunk(c);
        }
    }
    crew.finish();
    const std::vector<uint8_t> &src = toRing ? host : dev, &dst = toRing ? dev : host;
    for (size_t b = 0; b < nBlocks; ++b)
        for (size_t i = 0; i < pitch; ++i) {
            const uint8_t want = i < blockBytes ? src[b * pitch + i] : 0xEE;
            if (b * pitch + i < dst.size() && dst[b * pitch + i] != want) { printf("mismatch block %zu byte %zu\n", b, i); return 1; }
        }
    return 0;
}
int main() {
    int bad = 0;
    const size_t cases[][4] = {{270, 1920, 7680, 1 << 20}, {270, 5760, 7680, 65536}, {141, 1024, 4096, 65536}, {1080, 15360, 30720, 1 << 19}, {7, 100, 333, 4096}, {1, 5000, 9000, 8192}, {540, 7680, 15360, 65536}};
    for (auto &c : cases)
        for (int th : {1, 3, 4})
            for (int dir = 0; dir < 2; ++dir) bad += run(c[0], c[1], c[2], c[3], th, dir == 0);
    printf(bad ? "FAILED %d\n" : "ok\n", bad);
    return bad != 0;
}
