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
// a whole frame as hr_update_frame sends it with HR_STAGE_SPLIT=1: lattice rows of both planes, then the rows between
// them and the partial last groups, as ONE transfer of several segments (HrStagePlan::build_segments)
static int run_frame(int H, size_t rowBytes, int s, size_t chunk, int threads) {
    const size_t ylen = (size_t)H * rowBytes, uvlen = (size_t)(H / 2) * rowBytes;
    std::vector<uint8_t> host(ylen + uvlen), dev(ylen + uvlen, 0xEE), ring(HR_STAGE_SLOTS * chunk);
    for (size_t i = 0; i < host.size(); ++i) host[i] = (uint8_t)(i * 2246822519u >> 11);
    uint8_t *src[2] = {host.data(), host.data() + ylen}, *dpl[2] = {dev.data(), dev.data() + ylen};
    const int rows[2] = {H, H / 2}, stride[2] = {1 << s, (1 << s) / 2 > 1 ? (1 << s) / 2 : 1};
    HrStageSeg segs[6];
    int n = 0;
    for (int pl = 0; pl < 2; ++pl) {
        const size_t groups = (size_t)(rows[pl] + stride[pl] - 1) / stride[pl];
        if (stride[pl] > 1) segs[n++] = HrStageSeg{src[pl], dpl[pl], groups * rowBytes, rowBytes, stride[pl] * rowBytes};
        else segs[n++] = HrStageSeg{src[pl], dpl[pl], groups * rowBytes, 0, 0};
    }
    for (int pl = 0; pl < 2; ++pl) {
        const size_t full = (size_t)rows[pl] / stride[pl], tail = rows[pl] - full * stride[pl];
        if (stride[pl] > 1 && full > 0) segs[n++] = HrStageSeg{src[pl] + rowBytes, dpl[pl] + rowBytes, full * (stride[pl] - 1) * rowBytes, (stride[pl] - 1) * rowBytes, stride[pl] * rowBytes};
        if (tail > 1) {
            const size_t o = (full * stride[pl] + 1) * rowBytes;
            segs[n++] = HrStageSeg{src[pl] + o, dpl[pl] + o, (tail - 1) * rowBytes, 0, 0};
        }
    }
    HrStagePlan plan;
    plan.build_segments(segs, n, chunk);
    size_t total = 0;
    for (size_t c = 0; c < plan.n; ++c) {
        const HrStageSeg &sg = segs[plan.seg[c]];
        if (plan.len[c] == 0 || plan.len[c] > chunk || (sg.blockBytes && plan.len[c] % sg.blockBytes) || plan.segOff[c] + plan.len[c] > sg.bytes) { printf("bad chunk %zu\n", c); return 1; }
        total += plan.len[c];
    }
    if (total != ylen + uvlen) { printf("plan covers %zu of %zu bytes\n", total, ylen + uvlen); return 1; }
    HrCopyCrew crew(threads);
    crew.begin(true, NULL, ring.data(), chunk, 5, &plan);
    for (size_t c = 0; c < plan.n; ++c) {
        crew.release(c + 3 < plan.n ? c + 3 : plan.n); // the crew may run up to three slots ahead (< HR_STAGE_SLOTS)
        crew.wait_chunk(c);
        const HrStageSeg &sg = segs[plan.seg[c]];
        const uint8_t *slot = ring.data() + ((5 + c) % HR_STAGE_SLOTS) * chunk;
        if (!sg.blockBytes) memcpy(sg.dev + plan.segOff[c], slot, plan.len[c]);
        else
            for (size_t b = 0; b < plan.len[c] / sg.blockBytes; ++b) memcpy(sg.dev + (plan.segOff[c] / sg.blockBytes + b) * sg.pitch, slot + b * sg.blockBytes, sg.blockBytes);
    }
    crew.finish();
    if (memcmp(host.data(), dev.data(), host.size())) { printf("frame %dx%zu s=%d chunk %zu differs\n", H, rowBytes, s, chunk); return 1; }
    return 0;
}
int main() {
    int bad = 0;
    const size_t cases[][4] = {{270, 1920, 7680, 1 << 20}, {270, 5760, 7680, 65536}, {141, 1024, 4096, 65536}, {1080, 15360, 30720, 1 << 19}, {7, 100, 333, 4096}, {1, 5000, 9000, 8192}, {540, 7680, 15360, 65536}};
    for (auto &c : cases)
        for (int th : {1, 3, 4})
            for (int dir = 0; dir < 2; ++dir) bad += run(c[0], c[1], c[2], c[3], th, dir == 0);
    const size_t frames[][4] = {{1080, 1920, 2, 1 << 20}, {1080, 1920, 2, 65536}, {562, 1024, 2, 65536}, {480, 896, 1, 1 << 17}, {2160, 7680, 3, 1 << 19}, {270, 512, 1, 4096}, {1082, 1920, 3, 65536}, {4320, 15360, 4, 1 << 20}};
    for (auto &f : frames)
        for (int th : {1, 4}) bad += run_frame((int)f[0], f[1], (int)f[2], f[3], th);
    printf(bad ? "FAILED %d\n" : "ok\n", bad);
    return bad != 0;
}
