// cct_encode.cu — encoder front end: traversal table, Hilbert gather + block statistics, block meshing
// (segmentation) and op emission.  Everything up to (not including) the entropy stage.
//
// Reference path restated (reference src/codec/):
//   curve.py:61-138      traversal order           -> k_curve_pow2 (closed form) / host gilbert (general)
//   core.py:254-255      gather in traversal order  -> k_gather_stats
//   cluster.py:30-59     big-delta prefix sum, difficult test -> k_gather_stats (per-block counts, no prefix array)
//   cluster.py:79-190    look-ahead + first-fit meshing        -> k_mesh (one warp per image, 63 candidates in parallel)
//   core.py:274-330      JUMP/DELTA/FULL emission + EOF        -> k_emit (per-image scan of unit sizes)
#include <cuda_runtime.h>

#include <algorithm>

#include "cct_kernels.h"

namespace cct {

// ------------------------------------------------------------------ zero-fill
__global__ void __launch_bounds__(256) k_zero(uint4 *__restrict__ p16, size_t n16, uint8_t *__restrict__ tail, int ntail) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) p16[i] = make_uint4(0u, 0u, 0u, 0u);
    if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] = 0;
}
void launch_zero(void *p, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return;
    uint8_t *b = (uint8_t *)p;
    if (reinterpret_cast<uintptr_t>(p) & 15u) { cudaMemsetAsync(p, 0, bytes, st); return; }  // (workspace pieces are 256-byte aligned)
    const size_t n16 = bytes >> 4;
    const int ntail = (int)(bytes & 15u);
    const unsigned grid = (unsigned)std::min<size_t>((n16 + 1023) / 1024 + 1, 148 * 16);
    k_zero<<<grid, 256, 0, st>>>(reinterpret_cast<uint4 *>(b), n16, b + (n16 << 4), ntail);
}

// ------------------------------------------------------------------ traversal tables
__global__ void k_curve_pow2(int32_t *order, uint32_t n) {
    uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n * n) order[d] = (int32_t)hilbert_d2raster(n, d);
}
__global__ void k_iota(int32_t *order, uint32_t n) {
    uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) order[d] = (int32_t)d;
}

void launch_curve_pow2(int32_t *d_order, int n, cudaStream_t st) {
    uint32_t total = (uint32_t)n * (uint32_t)n;
    { CCT_KERNEL(k_curve_pow2, st);
    k_curve_pow2<<<(total + 255) / 256, 256, 0, st>>>(d_order, (uint32_t)n);
    }
}
void launch_iota(int32_t *d_order, int64_t n, cudaStream_t st) {
    { CCT_KERNEL(k_iota, st);
    k_iota<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_order, (uint32_t)n);
    }
}

// ------------------------------------------------------------------ gather + per-block statistics
// One thread per 16-pixel block.  info[b] = intra | lead << 4 where intra = number of |d|>64 transitions
// inside the block (15 of them, cluster.py:54-59) and lead = the transition into the block
// (cluster.py:109-110 reads prefix[16b-1]).
__global__ void __launch_bounds__(256) k_gather_stats(const uint16_t *__restrict__ px, const int32_t *__restrict__ order,
                                                      uint16_t *__restrict__ hv, uint8_t *__restrict__ info, int64_t n_px,
                                                      int nb, int64_t total_blocks) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_blocks) return;
    int64_t img = t / nb;
    int b = (int)(t - img * nb);
    const uint16_t *p = px + img * n_px;
    const int4 *o4 = reinterpret_cast<const int4 *>(order + (int64_t)b * kBlock);
    int idx[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int4 o = __ldg(o4 + i);
        idx[4 * i] = o.x; idx[4 * i + 1] = o.y; idx[4 * i + 2] = o.z; idx[4 * i + 3] = o.w;
    }
    int v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __ldg(p + idx[i]);
    int intra = 0;
#pragma unroll
    for (int i = 1; i < 16; i++) intra += (abs(v[i] - v[i - 1]) > 64);
    int lead = 0;
    if (b > 0) {
        int pv = __ldg(p + __ldg(order + (int64_t)b * kBlock - 1));
        lead = abs(v[0] - pv) > 64;
    }
    uint4 lo, hi;
    lo.x = v[0] | (v[1] << 16); lo.y = v[2] | (v[3] << 16); lo.z = v[4] | (v[5] << 16); lo.w = v[6] | (v[7] << 16);
    hi.x = v[8] | (v[9] << 16); hi.y = v[10] | (v[11] << 16); hi.z = v[12] | (v[13] << 16); hi.w = v[14] | (v[15] << 16);
    uint4 *dst = reinterpret_cast<uint4 *>(hv + img * n_px + (int64_t)b * kBlock);
    dst[0] = lo; dst[1] = hi;
    info[t] = (uint8_t)(intra | (lead << 4));
}

void launch_gather_stats(const uint16_t *px, const int32_t *order, uint16_t *hv, uint8_t *info, int B, int64_t n_px,
                         cudaStream_t st) {
    int nb = (int)(n_px / kBlock);
    int64_t total = (int64_t)B * nb;
    if (total == 0) return;
    { CCT_KERNEL(k_gather_stats, st);
    k_gather_stats<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(px, order, hv, info, n_px, nb, total);
    }
}

// ------------------------------------------------------------------ meshing
__device__ __forceinline__ void load_block16(const uint16_t *p, int (&v)[16]) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1];
    v[0] = a.x & 0xFFFF; v[1] = a.x >> 16; v[2] = a.y & 0xFFFF; v[3] = a.y >> 16;
    v[4] = a.z & 0xFFFF; v[5] = a.z >> 16; v[6] = a.w & 0xFFFF; v[7] = a.w >> 16;
    v[8] = b.x & 0xFFFF; v[9] = b.x >> 16; v[10] = b.y & 0xFFFF; v[11] = b.y >> 16;
    v[12] = b.z & 0xFFFF; v[13] = b.z >> 16; v[14] = b.w & 0xFFFF; v[15] = b.w >> 16;
}

// cluster.py:138-153: interleave A0 B0 A1 B1 ... and count diffs >= 65 (one-sided), + 1
__device__ __forceinline__ uint32_t mesh_cost(const int (&A)[16], const uint16_t *pB) {
    int Bv[16];
    load_block16(pB, Bv);
    uint32_t num = 1;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        num += (Bv[k] - A[k] >= 65);
        if (k < 15) num += (A[k + 1] - Bv[k] >= 65);
    }
    return num;
}

// One warp per image.  Walks the difficult blocks in order (cluster.py:79); for each one not yet taken as a
// partner the 63 look-ahead candidates are costed by the 32 lanes (two each) and the first fit wins
// (cluster.py:122-181).  `win` is the warp-uniform bitmap of taken blocks in [wbase, wbase + 64).
constexpr int kMeshWarps = 4;
__global__ void __launch_bounds__(32 * kMeshWarps) k_mesh(const uint16_t *__restrict__ hv, const uint8_t *__restrict__ info,
                                                          uint8_t *__restrict__ jump, uint8_t *__restrict__ taken,
                                                          int32_t *__restrict__ stats, int64_t n_px, int nb, int B) {
    int img = blockIdx.x * kMeshWarps + (threadIdx.x >> 5);
    if (img >= B) return;
    const int lane = threadIdx.x & 31;
    const uint16_t *h = hv + (int64_t)img * n_px;
    const uint8_t *inf = info + (int64_t)img * nb;
    uint8_t *jp = jump + (int64_t)img * nb;
    uint8_t *tk = taken + (int64_t)img * nb;
    uint64_t win = 0;
    int wbase = 0, n_difficult = 0, n_jump = 0;
    for (int base = 0; base < nb; base += 32) {
        int b = base + lane;
        int my = b < nb ? inf[b] : 0;
        unsigned dmask = __ballot_sync(0xffffffffu, (my & 15) >= kBlock / 2);  // cluster.py:58
        n_difficult += __popc(dmask);
        while (dmask) {
            int l = __ffs(dmask) - 1;
            dmask &= dmask - 1;
            int A = base + l;
            int shift = A - wbase;
            win = shift >= 64 ? 0ull : (win >> shift);
            wbase = A;
            if (win & 1ull) continue;  // completed[i] (cluster.py:98-99)
            int infA = __shfl_sync(0xffffffffu, my, l);
            uint32_t thr;
            if (A > 0) {
                thr = (uint32_t)((infA & 15) + ((infA >> 4) & 1)) - 2u;  // current_delta - 2 (cluster.py:110,158)
            } else {
                // block 0: prefix[15] - prefix[N-1] in uint32: T = all changes of the image
                uint32_t T = 0;
                for (int i = lane; i < nb; i += 32) { int f = inf[i]; T += (uint32_t)((f & 15) + ((f >> 4) & 1)); }
                for (int o = 16; o; o >>= 1) T += __shfl_xor_sync(0xffffffffu, T, o);
                thr = ((uint32_t)(infA & 15) - T) - 2u;
            }
            int Av[16];
            load_block16(h + (int64_t)A * kBlock, Av);
            int j1 = lane + 1, j2 = lane + 33;
            bool ok1 = false, ok2 = false;
            if (A + j1 < nb && !((win >> j1) & 1ull)) ok1 = mesh_cost(Av, h + (int64_t)(A + j1) * kBlock) < thr;
            if (j2 <= kLookAhead && A + j2 < nb && !((win >> j2) & 1ull)) ok2 = mesh_cost(Av, h + (int64_t)(A + j2) * kBlock) < thr;
            unsigned m1 = __ballot_sync(0xffffffffu, ok1), m2 = __ballot_sync(0xffffffffu, ok2);
            int j = m1 ? __ffs(m1) : (m2 ? 32 + __ffs(m2) : 0);
            if (j) {
                if (lane == 0) { jp[A] = (uint8_t)j; tk[A + j] = 1; }
                win |= 1ull << j;
                n_jump++;
            }
        }
    }
    if (lane == 0 && stats) { stats[4 * img + 2] = n_jump; stats[4 * img + 3] = n_difficult; }
}

void launch_mesh(const uint16_t *hv, const uint8_t *info, uint8_t *jump, uint8_t *taken, int32_t *stats, int B,
                 int64_t n_px, cudaStream_t st) {
    if (B == 0) return;
    int nb = (int)(n_px / kBlock);
    { CCT_KERNEL(k_mesh, st);
    k_mesh<<<(B + kMeshWarps - 1) / kMeshWarps, 32 * kMeshWarps, 0, st>>>(hv, info, jump, taken, stats, n_px, nb, B);
    }
}

// ------------------------------------------------------------------ emission
// Op bytes of one pixel step (core.py:313-323).  Returns the number of bytes (1 or 2).
__device__ __forceinline__ int op_bytes(int delta, uint8_t *out) {
    if (delta > -64 && delta < 65) {
        out[0] = (uint8_t)(delta & 127);  // (delta + 128) % 128
        return 1;
    }
    uint32_t u = (uint32_t)delta & 4095u;  // (delta + 4096) % 4096
    out[0] = (uint8_t)(0xE0u | (u >> 8));
    out[1] = (uint8_t)(u & 0xFFu);
    return 2;
}

constexpr int kEmitThreads = 256;
constexpr int kEmitUnitMax = 1 + 4 * kBlock;  // JUMP + 32 FULL ops
// One CTA per image; tiles of kEmitThreads blocks; each thread sizes its unit (16 px, or 32 px + JUMP for a
// meshed pair, nothing for a block taken as a partner), a block-wide exclusive scan places it, the thread writes
// its bytes into a shared-memory image of the tile, and the CTA copies the tile out in aligned 16-byte stores (the
// bytes past the last 16-byte boundary stay in shared memory and lead the next tile).
__global__ void __launch_bounds__(kEmitThreads) k_emit(const uint16_t *__restrict__ hv, const uint8_t *__restrict__ jump,
                                                       const uint8_t *__restrict__ taken, uint8_t *__restrict__ payload,
                                                       int64_t pstride, int32_t *__restrict__ plen, uint32_t *__restrict__ status,
                                                       int32_t *__restrict__ stats, int64_t n_px, int nb) {
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint16_t *h = hv + (int64_t)img * n_px;
    const uint8_t *jp = jump + (int64_t)img * nb;
    const uint8_t *tk = taken + (int64_t)img * nb;
    uint8_t *out = payload + (int64_t)img * pstride;  // 16-byte aligned (pstride is a multiple of 16)
    asm volatile("" : "+l"(h), "+l"(jp), "+l"(tk), "+l"(out));  // (kept in registers: the loops below would re-derive them from the kernel parameters)
    __shared__ int warp_sum[kEmitThreads / 32];
    __shared__ int red_full[kEmitThreads / 32], red_dom[kEmitThreads / 32];
    __shared__ __align__(16) uint8_t tile[kEmitThreads * kEmitUnitMax + 32];
    int n_full = 0, dom = 0;
    int done = 0;  // payload bytes before this tile (every thread keeps its own copy); done & 15 of them still sit in tile[]
    for (int base = 0; base < nb; base += kEmitThreads) {
        int b = base + tid;
        int size = 0, j = 0, prevv = 0;
        int Av[16], Bv[16];
        bool active = b < nb && !tk[b];
        if (active) {
            j = jp[b];
            load_block16(h + (int64_t)b * kBlock, Av);
            if (j) load_block16(h + (int64_t)(b + j) * kBlock, Bv);
            if (b > 0) {  // last pixel of the previous unit in output order
                int p = b - 1;
                while (tk[p]) p--;
                int pj = jp[p];
                prevv = h[(int64_t)(p + pj) * kBlock + 15];
            }
            int pv = prevv;
            size = j ? 1 : 0;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                int d = Av[k] - pv; pv = Av[k];
                int two = !(d > -64 && d < 65);
                size += 1 + two; n_full += two; dom |= (d < -2047 || d > 2048);
                if (j) {
                    d = Bv[k] - pv; pv = Bv[k];
                    two = !(d > -64 && d < 65);
                    size += 1 + two; n_full += two; dom |= (d < -2047 || d > 2048);
                }
            }
        }
        // block-wide exclusive scan of size
        int incl = size;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sum[wid] = incl;
        __syncthreads();
        int woff = 0, tile_total = 0;
        for (int w = 0; w < kEmitThreads / 32; w++) { if (w < wid) woff += warp_sum[w]; tile_total += warp_sum[w]; }
        const int held = done & 15;
        if (active) {
            uint8_t *o = tile + held + woff + incl - size;
            if (j) *o++ = (uint8_t)(0x80 | j);  // JUMP (core.py:290-294)
            int pv = prevv;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                o += op_bytes(Av[k] - pv, o); pv = Av[k];
                if (j) { o += op_bytes(Bv[k] - pv, o); pv = Bv[k]; }
            }
        }
        __syncthreads();
        const int end = held + tile_total, full = end & ~15;
        uint4 *dst = reinterpret_cast<uint4 *>(out + (done - held));
        for (int c = tid; c < (full >> 4); c += kEmitThreads) dst[c] = reinterpret_cast<const uint4 *>(tile)[c];
        uint8_t keep = 0;
        if (tid < end - full) keep = tile[full + tid];
        __syncthreads();
        if (tid < end - full) tile[tid] = keep;
        done += tile_total;
        // (the next tile's writes come after its first barrier)
    }
    // reductions for Encoder.info / status
    for (int o = 16; o; o >>= 1) { n_full += __shfl_xor_sync(0xffffffffu, n_full, o); dom |= __shfl_xor_sync(0xffffffffu, dom, o); }
    if (lane == 0) { red_full[wid] = n_full; red_dom[wid] = dom; }
    __syncthreads();
    if (tid < (done & 15)) out[(done & ~15) + tid] = tile[tid];  // what the last tile left behind
    if (tid == 0) {
        out[done] = (uint8_t)kEOF;  // core.py:329-330
        plen[img] = done + 1;
        int nf = 0, dm = 0;
        for (int w = 0; w < kEmitThreads / 32; w++) { nf += red_full[w]; dm |= red_dom[w]; }
        if (stats) { stats[4 * img + 0] = (int32_t)(n_px - nf); stats[4 * img + 1] = nf; }
        if (status) status[img] = dm ? 1u : 0u;
    }
}

void launch_emit(const uint16_t *hv, const uint8_t *jump, const uint8_t *taken, uint8_t *payload, int64_t pstride,
                 int32_t *plen, uint32_t *status, int32_t *stats, int B, int64_t n_px, cudaStream_t st) {
    if (B == 0) return;
    { CCT_KERNEL(k_emit, st);
    k_emit<<<B, kEmitThreads, 0, st>>>(hv, jump, taken, payload, pstride, plen, status, stats, n_px, (int)(n_px / kBlock));
    }
}

}  // namespace cct
