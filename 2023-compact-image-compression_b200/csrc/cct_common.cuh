// cct_common.cuh — constants and small host/device helpers shared by every stage.
//
// Format constants restate the reference's config and tag table:
//   block size 16, EOF byte 59            reference src/config.json:9,15
//   DELTA 0xxxxxxx / JUMP 10xxxxxx / FULL 1110xxxx   reference src/codec/core.py:40-50
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CCT_HD __host__ __device__ __forceinline__
#define CCT_HDN __host__ __device__
#else
#define CCT_HD inline
#define CCT_HDN
#endif

namespace cct {

constexpr int kBlock = 16;        // pixels per segmentation block
constexpr int kEOF = 59;          // end_of_file byte
constexpr int kHeaderBytes = 13;  // core.py:193-210
constexpr int kLookAhead = 63;    // cluster.py:122  blocks[i+1 : i+64]

// zlib level-9 parameters
constexpr int kWSize = 32768;
constexpr int kMaxDist = kWSize - 262;  // 32506
constexpr int kMinMatch = 3;
constexpr int kMaxMatch = 258;
constexpr int kTooFar = 4096;
constexpr int kSymEnd = 16383;  // symbols per block before a flush
constexpr int kChainFull = 4096, kChainGood = 1024, kGoodLen = 32;

CCT_HD int ilog2(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}

// payload capacity of one image: every pixel FULL (2N) + one JUMP per two blocks + EOF, padded to 16
CCT_HD int64_t payload_bound(int64_t n_px) { return ((2 * n_px + n_px / (2 * kBlock) + 1) + 15) / 16 * 16; }
// zlib stream bound for `len` payload bytes: stored blocks cost 5 B per 16383 symbols (every block holds
// at least 16383 bytes unless it is the last), + 2 header + 4 trailer, + slack for the bit packer's word stores
CCT_HD int64_t deflate_bound(int64_t len) { return len + 5 * (len / kSymEnd + 2) + 6 + 16; }

// Classic Hilbert d2xy for a power-of-two square of side n: position d along the curve -> raster index.
// equals the reference's generalized curve (curve.py:83-138) for n in {2,4,...}.
CCT_HD uint32_t hilbert_d2raster(uint32_t n, uint32_t d) {
    uint32_t x = 0, y = 0, t = d;
    for (uint32_t s = 1; s < n; s <<= 1) {
        uint32_t rx = 1u & (t >> 1), ry = 1u & (t ^ rx);
        if (ry == 0) {
            if (rx == 1) { x = s - 1 - x; y = s - 1 - y; }
            uint32_t tmp = x; x = y; y = tmp;
        }
        x += s * rx; y += s * ry;
        t >>= 2;
    }
    return y * n + x;
}

#if defined(__CUDACC__)
// Adler-32 building blocks: sa += sum of d_i, sb += sum of d_i * (n - i) over the calling thread's share of w[0..n)
// (A = 1 + sa, B = n + sb, both mod 65521).  Sixteen bytes per load where the row is 16-byte aligned, and two
// instructions per word: the byte sum (vsadu4) and the byte sum weighted 0,1,2,3 (dp4a).
__device__ __forceinline__ void adler_sums(const uint8_t *w, int n, int tid, int nthreads, unsigned long long &sa,
                                           unsigned long long &sb) {
    const uint4 *v = reinterpret_cast<const uint4 *>(w);
    const int nv = (reinterpret_cast<uintptr_t>(w) & 15u) ? 0 : n >> 4;  // (a row that is not 16-byte aligned goes byte by byte)
    for (int i = tid; i < nv; i += nthreads) {
        const uint4 q = v[i];
        const uint32_t x[4] = {q.x, q.y, q.z, q.w};
        uint32_t s = 0, t = 0;  // s = sum of the 16 bytes, t = sum of j * byte_j, j = 0..15
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t sk = __vsadu4(x[k], 0u);
            s += sk;
            t += __dp4a(x[k], 0x03020100u, 0u) + 4u * (uint32_t)k * sk;
        }
        sa += s;
        sb += (unsigned long long)s * (unsigned long long)(n - 16 * i) - t;
    }
    for (int i = (nv << 4) + tid; i < n; i += nthreads) {
        const unsigned long long d = w[i];
        sa += d; sb += d * (unsigned long long)(n - i);
    }
}
#endif

}  // namespace cct
