// cct_deflate.cu — bit-exact zlib-1.3 level-9 `compress2` on the GPU, one zlib stream per image payload.
//
// Replaces `zlib.compress(data, level = 9)` (reference src/codec/core.py:340).  Stages (see
// deflate_core.cuh for the decomposition and SURVEY.md Appendix B for the algorithm being matched):
//   k_run_edges / k_run_link / k_run_counts   equal-byte run lengths around every position
//   k_hs_hist / k_hs_scan / k_hs_split / k_hs_bins   the positions of every stream sorted by (3-byte hash, position):
//                  a hash chain is then a contiguous piece of the sorted array
//   k_bprobe       cold-start longest_match for every position, one thread per sorted entry scanning the entries
//                  below it (coalesced), results scattered by position
//   k_next / k_walk_spec / k_walk_stitch / k_symscan   deflate_slow's lazy state machine as a windowed speculative walk
//   k_adler        Adler-32 of the payload (CTA per stream)
//   k_block_build  per 16383-symbol block: histograms, build_tree x3, block-type choice, header bits
//   k_layout       bit offsets of the blocks of each stream, stream length, zlib header/trailer
//   k_pack         Huffman/extra bits of every symbol, placed by a prefix sum of code lengths
//   k_offsets / k_copy   .cct header + stream -> caller's arena at scanned offsets
#include <cuda_runtime.h>

#include <cstddef>

#include <algorithm>
#include <mutex>

#include "cct_kernels.h"
#include "deflate_core.cuh"

namespace cct {

#ifndef CCT_WALK_WIN
#define CCT_WALK_WIN 1024
#endif
constexpr int kWalkWin = CCT_WALK_WIN;  // positions per speculative parse window
constexpr int kHsBins = 128;   // bins of the first sort pass (sort_digit1)
constexpr int kHsTile = 2048;  // positions per tile of the first sort pass (4096: 6.6 ms in k_hs_split, 1024: 6.4 + longer scans; 2048: 5.4)

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

size_t deflate_ws_bytes(int B, int64_t max_len) {
    DeflateWs ws;
    return deflate_ws_carve(&ws, nullptr, B, max_len);
}

size_t deflate_ws_carve(DeflateWs *ws, void *mem, int B, int64_t max_len) {
    char *p = (char *)mem;
    size_t used = 0;
    auto take = [&](size_t bytes) { void *r = p ? p + used : nullptr; used += (size_t)align_up((int64_t)bytes, 256); return r; };
    ws->stride = align_up(max_len + 16, 16);
    ws->max_blocks = (int)(max_len / kSymEnd + 2);
    ws->ent = (SortEnt *)take((size_t)B * ws->stride * sizeof(SortEnt));
    ws->runs = (uint32_t *)take((size_t)B * ws->stride * sizeof(uint32_t));
    ws->mtab = (uint32_t *)take((size_t)B * ws->stride * 2 * sizeof(uint32_t));  // both probe tables, one block
    ws->mhi = ws->mtab ? ws->mtab + (size_t)B * ws->stride : nullptr;
    ws->p1 = (SortEnt *)ws->mtab;  // pass-1 output of the sort (8 bytes per position): dead before k_bprobe fills the tables
    ws->syms = ws->mtab;  // the dense symbol array reuses the probe table once the walk is over
    ws->rec = (uint32_t *)take((size_t)B * ws->stride * sizeof(uint32_t));
    ws->counters = (uint32_t *)take(64 * sizeof(uint32_t));
    ws->max_tiles = (int)((max_len + kHsTile - 1) / kHsTile);
    ws->thist = (uint32_t *)take((size_t)B * ws->max_tiles * kHsBins * sizeof(uint32_t));
    ws->binbase = (uint32_t *)take((size_t)B * (kHsBins + 1) * sizeof(uint32_t));
    ws->mask_stride = align_up(ws->stride / 32 + 1, 4);
    ws->mask = (uint32_t *)take((size_t)B * ws->mask_stride * sizeof(uint32_t));
    ws->max_win = (int)(max_len / kWalkWin + 1);
    ws->exits = (Walk *)take((size_t)B * ws->max_win * sizeof(Walk));
    ws->blk_end = (int32_t *)take((size_t)B * ws->max_blocks * sizeof(int32_t));
    ws->nsym = (int32_t *)take((size_t)B * sizeof(int32_t));
    ws->nflush = (int32_t *)take((size_t)B * sizeof(int32_t));
    ws->adler = (uint32_t *)take((size_t)B * sizeof(uint32_t));
    ws->desc = (BlockDesc *)take((size_t)B * ws->max_blocks * sizeof(BlockDesc));
    return used;
}

// ------------------------------------------------------------------ k_run_*
// Run lengths around every position: rb[p] / rl[p] = number of bytes equal to w[p] immediately before / after p,
// packed into runs[p] = rl | rb << 16 (each capped at 65535) for the group step of the scan.  Everything follows from
// one flag per position, E[p] = (w[p] == w[p - 1]): rb counts the flags set up to p, rl those set from p + 1 on.
// One warp owns a chunk of kRunChunk positions, 128 per step: a lane takes one aligned 32-bit word of the payload, its
// four flags are one SIMD byte compare, and "where is the nearest cleared flag" is a ballot of the lanes whose four
// flags are not all set, a count-leading-zeros and one shuffle -- no shared memory, 4-byte loads, 16-byte stores.
// Chunks are independent once the count entering each is known, so the work is three launches:
//   k_run_edges  per chunk: (does one run cover the whole chunk?, count leaving it if nothing came in), both ways
//   k_run_link   per stream: warp scan of those summaries -> the count entering every chunk from either side
//   k_run_counts per chunk: counts of every position, written out
// The summaries combine as (whole, val) o (whole', val') = whole' ? (whole, val + val') : (false, val').
// (The payload rows are 16-byte aligned: include/compact_b200.h.)
constexpr int kRunChunk = 1024;
constexpr int kRunSteps = kRunChunk / 128;
constexpr int kRunWarps = 8;
constexpr unsigned kAll = 0xffffffffu;

struct RunCarry { int whole, val; };
__device__ __forceinline__ RunCarry run_combine(RunCarry a, RunCarry b) {  // a = earlier part, b = later part
    return b.whole ? RunCarry{a.whole, a.val + b.val} : RunCarry{0, b.val};
}

// The flags of a chunk: nibble j of `e` = E of the lane's four positions in step j (position base + 128 j + 4 lane + k
// in bit k); `after` = E[base + kRunChunk] (the same in every lane).
struct RunFlags { uint32_t e, after; };
__device__ __forceinline__ RunFlags run_flags(const uint32_t *words, int n, int base, int lane) {
    const int end = n;
    RunFlags r{0u, 0u};
    uint32_t last = base > 0 ? __ldg(words + (base >> 2) - 1) : 0u;  // the word before the step's (lane 0 needs its top byte)
#pragma unroll
    for (int j = 0; j < kRunSteps; j++) {
        const int p0 = base + 128 * j + 4 * lane;
        const uint32_t w = p0 < end ? __ldg(words + (p0 >> 2)) : 0u;
        uint32_t pb = __shfl_up_sync(kAll, w, 1);
        if (lane == 0) pb = last;
        last = __shfl_sync(kAll, w, 31);
        const uint32_t cmp = __vcmpeq4(w, (w << 8) | (pb >> 24));  // byte k: 0xff where w[p0 + k] == w[p0 + k - 1]
        // E is defined for 1 <= p < n
        const int lo = 1 - p0, hi = end - p0;
        const uint32_t vm = (lo <= 0 ? 0xffffffffu : lo >= 4 ? 0u : 0xffffffffu << (8 * lo)) &
                            (hi >= 4 ? 0xffffffffu : hi <= 0 ? 0u : 0xffffffffu >> (8 * (4 - hi)));
        r.e |= (((cmp & vm & 0x01010101u) * 0x01020408u) >> 24) << (4 * j);  // bit 0 of each byte -> a nibble
    }
    const int q = base + kRunChunk;
    if (q < end) r.after = (__ldg(words + (q >> 2)) & 0xFFu) == (last >> 24) ? 1u : 0u;
    return r;
}
__device__ __forceinline__ int top_ones(uint32_t nib) { return __clz((int)~(nib << 28)); }    // set flags at the top of a nibble (0..4)
__device__ __forceinline__ int low_ones(uint32_t nib) { return __ffs((int)(~nib & 0x1Fu | 0x10u)) - 1; }  // ... at the bottom
// the nibble of "next" flags F[p] = E[p + 1] of the lane's positions in step j
__device__ __forceinline__ uint32_t run_next_nibble(const RunFlags &r, int j, int lane) {
    const uint32_t e4 = (r.e >> (4 * j)) & 0xFu;
    uint32_t nb = __shfl_down_sync(kAll, e4, 1) & 1u;
    const uint32_t wrap = j + 1 < kRunSteps ? __shfl_sync(kAll, (r.e >> (4 * (j + 1) & 31)) & 1u, 0) : r.after;
    if (lane == 31) nb = wrap;
    return (e4 >> 1) | (nb << 3);
}

__global__ void __launch_bounds__(32 * kRunWarps) k_run_edges(const uint8_t *__restrict__ in, int64_t in_stride,
                                                             const int32_t *__restrict__ in_len, int4 *__restrict__ edges, int max_chunks) {
    const int stream = blockIdx.y, chunk = blockIdx.x * kRunWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int n = in_len[stream], base = chunk * kRunChunk;
    if (base >= n) return;
    const RunFlags r = run_flags(reinterpret_cast<const uint32_t *>(in + (int64_t)stream * in_stride), n, base, lane);
    int fw = 1, fv = 0, bw = 1, bv = 0;
#pragma unroll
    for (int j = 0; j < kRunSteps; j++) {
        const uint32_t e4 = (r.e >> (4 * j)) & 0xFu;
        const unsigned open = __ballot_sync(kAll, e4 != 0xFu);  // lanes with a cleared flag
        if (!open) fv += 128;
        else { const int l = 31 - __clz((int)open); fw = 0; fv = __shfl_sync(kAll, top_ones(e4), l) + 4 * (31 - l); }
    }
#pragma unroll
    for (int j = kRunSteps - 1; j >= 0; j--) {
        const uint32_t f4 = run_next_nibble(r, j, lane);
        const unsigned open = __ballot_sync(kAll, f4 != 0xFu);
        if (!open) bv += 128;
        else { const int l = __ffs((int)open) - 1; bw = 0; bv = __shfl_sync(kAll, low_ones(f4), l) + 4 * l; }
    }
    if (lane == 0) edges[(int64_t)stream * max_chunks + chunk] = make_int4(fw, fv, bw, bv);
}

// enters[chunk] = (count entering the chunk from the left, count entering it from the right); one warp per stream
__global__ void __launch_bounds__(32 * kRunWarps) k_run_link(const int32_t *__restrict__ in_len, const int4 *__restrict__ edges,
                                                            int2 *__restrict__ enters, int max_chunks, int B) {
    const int stream = blockIdx.x * kRunWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (stream >= B) return;
    const int nt = (in_len[stream] + kRunChunk - 1) / kRunChunk;
    const int4 *e = edges + (int64_t)stream * max_chunks;
    int2 *o = enters + (int64_t)stream * max_chunks;
    for (int dir = 0; dir < 2; dir++) {
        int carry = 0;
        for (int t0 = 0; t0 < nt; t0 += 32) {
            const int k = t0 + lane;                      // k-th chunk in scan order
            const int t = dir ? nt - 1 - k : k;
            RunCarry incl{1, 0};
            if (k < nt) { const int4 v = e[t]; incl = dir ? RunCarry{v.z, v.w} : RunCarry{v.x, v.y}; }
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                RunCarry a{__shfl_up_sync(kAll, incl.whole, d), __shfl_up_sync(kAll, incl.val, d)};
                if (lane >= d) incl = run_combine(a, incl);
            }
            RunCarry excl{__shfl_up_sync(kAll, incl.whole, 1), __shfl_up_sync(kAll, incl.val, 1)};
            if (lane == 0) excl = RunCarry{1, 0};
            const int in_val = run_combine(RunCarry{1, carry}, excl).val;
            if (k < nt) { if (dir) o[t].y = in_val; else o[t].x = in_val; }
            const RunCarry tot{__shfl_sync(kAll, incl.whole, 31), __shfl_sync(kAll, incl.val, 31)};
            carry = run_combine(RunCarry{1, carry}, tot).val;
        }
    }
}

__global__ void __launch_bounds__(32 * kRunWarps) k_run_counts(const uint8_t *__restrict__ in, int64_t in_stride,
                                                              const int32_t *__restrict__ in_len, const int2 *__restrict__ enters,
                                                              int max_chunks, uint32_t *__restrict__ runs_all, int64_t stride) {
    const int stream = blockIdx.y, chunk = blockIdx.x * kRunWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int n = in_len[stream], base = chunk * kRunChunk;
    if (base >= n) return;
    uint32_t *runs = runs_all + (int64_t)stream * stride;
    asm volatile("" : "+l"(runs));  // (kept in registers instead of re-derived from the parameters)
    const RunFlags r = run_flags(reinterpret_cast<const uint32_t *>(in + (int64_t)stream * in_stride), n, base, lane);
    const int2 enter = enters[(int64_t)stream * max_chunks + chunk];
    uint32_t rb[kRunSteps];  // the four counts of step j, a byte each would not do: two 16-bit halves per word, two words
    uint32_t rb2[kRunSteps];
    const unsigned below = (1u << lane) - 1u;
    int carry = enter.x;  // count of the position just before the step
#pragma unroll
    for (int j = 0; j < kRunSteps; j++) {
        const uint32_t e4 = (r.e >> (4 * j)) & 0xFu;
        const unsigned open = __ballot_sync(kAll, e4 != 0xFu) & below;  // lanes below with a cleared flag
        const int l = open ? 31 - __clz((int)open) : 0;
        const int t = __shfl_sync(kAll, top_ones(e4), l);
        const int ent = open ? t + 4 * (lane - l - 1) : carry + 4 * lane;  // count of the position just before the lane's four
        const int b0 = (e4 & 1u) ? ent + 1 : 0, b1 = (e4 & 2u) ? b0 + 1 : 0, b2 = (e4 & 4u) ? b1 + 1 : 0, b3 = (e4 & 8u) ? b2 + 1 : 0;
        carry = __shfl_sync(kAll, b3, 31);
        rb[j] = (uint32_t)min(b0, 65535) | ((uint32_t)min(b1, 65535) << 16);
        rb2[j] = (uint32_t)min(b2, 65535) | ((uint32_t)min(b3, 65535) << 16);
    }
    const unsigned above = ~below & ~(1u << lane);
    carry = enter.y;  // count of the position just after the step
#pragma unroll
    for (int j = kRunSteps - 1; j >= 0; j--) {
        const uint32_t f4 = run_next_nibble(r, j, lane);
        const unsigned open = __ballot_sync(kAll, f4 != 0xFu) & above;  // lanes above with a cleared flag
        const int l = open ? __ffs((int)open) - 1 : 0;
        const int t = __shfl_sync(kAll, low_ones(f4), l);
        const int ext = open ? t + 4 * (l - lane - 1) : carry + 4 * (31 - lane);  // count of the position just after the lane's four
        const int f3 = (f4 & 8u) ? ext + 1 : 0, f2 = (f4 & 4u) ? f3 + 1 : 0, f1 = (f4 & 2u) ? f2 + 1 : 0, f0 = (f4 & 1u) ? f1 + 1 : 0;
        carry = __shfl_sync(kAll, f0, 0);
        const int p0 = base + 128 * j + 4 * lane;
        const uint4 v = make_uint4((uint32_t)min(f0, 65535) | (rb[j] << 16), (uint32_t)min(f1, 65535) | (rb[j] & 0xFFFF0000u),
                                   (uint32_t)min(f2, 65535) | (rb2[j] << 16), (uint32_t)min(f3, 65535) | (rb2[j] & 0xFFFF0000u));
        if (p0 + 4 <= n) *reinterpret_cast<uint4 *>(runs + p0) = v;
        else {
            if (p0 < n) runs[p0] = v.x;
            if (p0 + 1 < n) runs[p0 + 1] = v.y;
            if (p0 + 2 < n) runs[p0 + 2] = v.z;
        }
    }
}

// ------------------------------------------------------------------ k_hs_*: positions sorted by (hash, position)
// The chain members of a stream (positions 1 .. n-3) are sorted by their 15-bit hash, stably, so that every hash
// bucket lists its positions in increasing order -- zlib's chain, read backwards (deflate_core.cuh, "longest_match
// over hash-sorted positions").  The order of the buckets among themselves is irrelevant, so the sort key is the hash
// re-expressed as two digits that spread a CT payload evenly (sort_digit1 / sort_digit2 in deflate_core.cuh: the top
// seven bits alone put two thirds of the positions into bins too large for one chunk of pass 2).  Two radix passes:
//   pass 1  by the 7-bit digit (128 bins), the usual three launches over tiles of kHsTile positions:
//           k_hs_hist   per tile: how many positions fall into each bin
//           k_hs_scan   per stream: bin starts, and for every tile the offset of its share of each bin
//           k_hs_split  per tile: (position, first four bytes) written to p1[] at bin start + offset + rank inside the tile
//   pass 2  by the 8-bit digit, one CTA per (stream, bin): digit histogram, scan, ranked scatter of the finished
//           8-byte entries (position, flags, first four bytes) into ent[] -- a bin is a contiguous piece of p1[] and of
//           ent[], and is small enough (n / 128 positions on average) that its second read comes from L1/L2.  The
//           payload itself is read once, coalesced, by pass 1.
// Ranks inside a group of 32 keys: the lanes holding the same digit (key_peers), and of those the ones below.
constexpr int kHsThreads = 256;
constexpr int kHsWarps = kHsThreads / 32;
constexpr int kHsRows = kHsTile / kHsThreads;  // 32-position rows per warp

// Lanes of the warp holding the same key (BITS significant bits; idle lanes pass valid = false and match nobody).
// One ballot per bit: MATCH.ANY costs about two cycles per DISTINCT key among the lanes (measured 62 cycles per warp for
// random 8-bit keys, profiles/r02_match_bench.txt), which is the usual case here; a row inside a run -- all keys
// equal -- is settled by one vote.
template <int BITS>
__device__ __forceinline__ unsigned key_peers(uint32_t key, bool valid) {
    const uint32_t k0 = __shfl_sync(kAll, key, 0);
    if (__all_sync(kAll, valid && key == k0)) return kAll;
    unsigned peers = __ballot_sync(kAll, valid);
#pragma unroll
    for (int b = 0; b < BITS; b++) {
        // peers &= ballot(bit b of key) ^ (bit b ? 0 : ~0), in four instructions (mask test straight into a predicate,
        // vote, select, one LOP3); the compiler's own sequence for the C++ statement takes six
        asm("{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b32 t, bal, m;\n\t"
            "and.b32 t, %1, %2;\n\t"
            "setp.ne.u32 p, t, 0;\n\t"
            "vote.sync.ballot.b32 bal, p, 0xffffffff;\n\t"
            "selp.b32 m, 0, 0xffffffff, p;\n\t"
            "xor.b32 bal, bal, m;\n\t"
            "and.b32 %0, %0, bal;\n\t"
            "}"
            : "+r"(peers) : "r"(key), "r"(1u << b));
    }
    return valid ? peers : 0u;
}

// Tile bytes [base, base + kHsTile + 3) -> shared memory (zero past the end), then the hash digit of each position.
struct HsTileSmem { uint32_t words[kHsTile / 4 + 4]; };
__device__ __forceinline__ void hs_stage(HsTileSmem &sm, const uint8_t *__restrict__ w, int n, int base) {
    const uint4 *src = reinterpret_cast<const uint4 *>(w + base);  // base is a multiple of 16, the row 16-byte aligned
    uint4 *dst = reinterpret_cast<uint4 *>(sm.words);
    for (int i = threadIdx.x; i < kHsTile / 16 + 1; i += kHsThreads) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (base + 16 * i < n) v = __ldg(src + i);  // (the 16 bytes may reach past n: those never enter a valid hash)
        dst[i] = v;
    }
}
__device__ __forceinline__ uint32_t hs_hash_at(const HsTileSmem &sm, int off) {  // hash3 of the bytes at tile offset off
    const uint32_t a = sm.words[off >> 2], b = sm.words[(off >> 2) + 1];
    return hash_quad(__funnelshift_r(a, b, (off & 3) * 8));
}

__global__ void __launch_bounds__(kHsThreads) k_hs_hist(const uint8_t *__restrict__ in, int64_t in_stride,
                                                        const int32_t *__restrict__ in_len, uint32_t *__restrict__ thist,
                                                        int max_tiles) {
    __shared__ HsTileSmem sm;
    __shared__ uint32_t hist[kHsBins];
    const int stream = blockIdx.y, tile = blockIdx.x, n = in_len[stream], base = tile * kHsTile;
    if (base >= n) return;
    if (threadIdx.x < kHsBins) hist[threadIdx.x] = 0;
    hs_stage(sm, in + (int64_t)stream * in_stride, n, base);
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < kHsRows; r++) {
        const int off = r * kHsThreads + threadIdx.x, p = base + off;
        const bool valid = p >= 1 && p + 2 < n;
        const uint32_t d = sort_digit1(hs_hash_at(sm, off));
        const uint32_t d0 = __shfl_sync(kAll, d, 0);
        if (__all_sync(kAll, valid && d == d0)) {  // a run: one add for the warp
            if ((threadIdx.x & 31) == 0) atomicAdd(&hist[d0], 32u);
        } else if (valid) atomicAdd(&hist[d], 1u);
    }
    __syncthreads();
    if (threadIdx.x < kHsBins) thist[((int64_t)stream * max_tiles + tile) * kHsBins + threadIdx.x] = hist[threadIdx.x];
}

// counts -> absolute offsets inside the stream: thist[tile][bin] = start of the tile's share of the bin;
// binbase[stream][bin] = start of the bin (binbase[stream][128] = number of chain members)
__global__ void __launch_bounds__(kHsBins) k_hs_scan(const int32_t *__restrict__ in_len, uint32_t *__restrict__ thist, int max_tiles,
                                                     uint32_t *__restrict__ binbase, uint32_t *__restrict__ tails, int64_t stride) {
    __shared__ uint32_t tot[kHsBins];
    const int stream = blockIdx.x, d = threadIdx.x, n = in_len[stream];
    if (d == 0) tails[(int64_t)stream * stride] = 0;  // the stream's list of run tails (k_bprobe appends, k_run_tails works it off)
    const int nt = (n + kHsTile - 1) / kHsTile;
    uint32_t *h = thist + (int64_t)stream * max_tiles * kHsBins + d;
    uint32_t run = 0;
    for (int t = 0; t < nt; t++) { const uint32_t c = h[(int64_t)t * kHsBins]; h[(int64_t)t * kHsBins] = run; run += c; }
    tot[d] = run;
    __syncthreads();
    uint32_t start = 0;  // exclusive scan over the bins (128 values: a serial sum per thread is as fast as anything)
    for (int k = 0; k < d; k++) start += tot[k];
    for (int t = 0; t < nt; t++) h[(int64_t)t * kHsBins] += start;
    binbase[(int64_t)stream * (kHsBins + 1) + d] = start;
    if (d == kHsBins - 1) binbase[(int64_t)stream * (kHsBins + 1) + kHsBins] = start + run;
}

__global__ void __launch_bounds__(kHsThreads, 6) k_hs_split(const uint8_t *__restrict__ in, int64_t in_stride,
                                                         const int32_t *__restrict__ in_len, const uint32_t *__restrict__ thist,
                                                         int max_tiles, SortEnt *__restrict__ p1_all, int64_t stride) {
    __shared__ HsTileSmem sm;
    __shared__ uint32_t wc[kHsWarps][kHsBins];  // per warp: count of each bin, then the offset of the warp's share
    const int stream = blockIdx.y, tile = blockIdx.x, n = in_len[stream], base = tile * kHsTile;
    if (base >= n) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kHsWarps * kHsBins; i += kHsThreads) (&wc[0][0])[i] = 0;
    hs_stage(sm, in + (int64_t)stream * in_stride, n, base);
    __syncthreads();
    // warp `wid` owns the 32 * kHsRows consecutive positions from base + wid * 32 * kHsRows: (warp, row, lane) is
    // position order, so ranks taken in that order keep the sort stable
    uint32_t key[kHsRows];  // bin | rank inside the warp's share << 8   (0xFFFFFFFF: not a chain member)
    const unsigned below = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < kHsRows; r++) {
        const int off = (wid * kHsRows + r) * 32 + lane, p = base + off;
        const bool valid = p >= 1 && p + 2 < n;
        const uint32_t d = sort_digit1(hs_hash_at(sm, off));
        const unsigned peers = key_peers<7>(d, valid);
        uint32_t old = 0;
        if (valid) old = wc[wid][d];
        __syncwarp();
        if (valid && (peers >> lane) == 1u) wc[wid][d] = old + (uint32_t)__popc(peers);  // the highest peer books the group
        __syncwarp();
        key[r] = valid ? (d | ((old + (uint32_t)__popc(peers & below)) << 8)) : 0xFFFFFFFFu;
    }
    __syncthreads();
    if (threadIdx.x < kHsBins) {
        uint32_t run = thist[((int64_t)stream * max_tiles + tile) * kHsBins + threadIdx.x];
        for (int k = 0; k < kHsWarps; k++) { const uint32_t c = wc[k][threadIdx.x]; wc[k][threadIdx.x] = run; run += c; }
    }
    __syncthreads();
    SortEnt *p1 = p1_all + (int64_t)stream * stride;
    asm volatile("" : "+l"(p1));  // (kept in registers: the stores below would re-derive it from the parameters)
#pragma unroll
    for (int r = 0; r < kHsRows; r++) {
        if (key[r] == 0xFFFFFFFFu) continue;
        const int off = (wid * kHsRows + r) * 32 + lane, p = base + off;
        uint32_t q = __funnelshift_r(sm.words[off >> 2], sm.words[(off >> 2) + 1], (off & 3) * 8);
        if (p + 3 >= n) q &= 0x00FFFFFFu;  // (the staged bytes may reach past the end of the payload)
        *reinterpret_cast<uint2 *>(p1 + wc[wid][key[r] & 0xFFu] + (key[r] >> 8)) = make_uint2((uint32_t)p, q);
    }
}

// pass 2: one CTA per (stream, bin), bins handed out through a counter (they differ in size by orders of magnitude:
// the bin of the zero run holds a quarter of a CT payload).  A bin is worked off in chunks of kBinChunk entries held in
// registers, warp w owning rows 8w .. 8w+7 of the chunk so that (chunk, warp, row, lane) is position order:
//   digit totals   a bin of one chunk (the usual case) gets them from the ranking itself; a longer one from a first
//                  sweep over its entries (shared-memory histogram)
//   per chunk      rank inside the warp (key_peers + per-warp counters), offsets of the warps' shares, scatter of the
//                  finished entries -- the bin's piece of ent[] is a few KB, so the 8-byte stores meet in L2.
constexpr int kBinThreads = 256;
constexpr int kBinWarps = kBinThreads / 32;
#ifndef CCT_BIN_ROWS
#define CCT_BIN_ROWS 8
#endif
constexpr int kBinRows = CCT_BIN_ROWS;                       // rows (of 32 entries) per warp and chunk
constexpr int kBinChunk = kBinThreads * kBinRows;            // 2048 entries
struct BinSmem {
    uint32_t wc[kBinWarps][256];  // per warp: count of each digit in the chunk, then the offset of the warp's share
    uint32_t run[256];            // next free slot of each digit
    uint32_t first[256];          // first slot of each digit (its lowest position gets kEntFirst)
    uint32_t warp_sum[kBinWarps];
    int unit;
};

__device__ __forceinline__ void bin_rank_chunk(BinSmem &sm, const uint2 *__restrict__ src, int S, int c0, uint2 (&pq)[kBinRows],
                                               uint32_t (&key)[kBinRows]) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned below = (1u << lane) - 1u;
    const int w0 = c0 + wid * kBinRows * 32;  // the warp's first entry
#pragma unroll
    for (int r = 0; r < kBinRows; r++) {
        const int k = w0 + r * 32 + lane;
        pq[r] = k < S ? __ldg(src + k) : make_uint2(0xFFFFFFFFu, 0u);
        key[r] = 0xFFFFFFFFu;
    }
#pragma unroll
    for (int r = 0; r < kBinRows; r++) {
        if (w0 + r * 32 >= S) break;  // (warp-uniform: the rows past the end of the bin cost nothing)
        const bool valid = pq[r].x != 0xFFFFFFFFu;
        const uint32_t d = sort_digit2(hash_quad(pq[r].y));
        const unsigned peers = key_peers<8>(d, valid);
        uint32_t old = 0;
        if (valid) old = sm.wc[wid][d];
        __syncwarp();
        if (valid && (peers >> lane) == 1u) sm.wc[wid][d] = old + (uint32_t)__popc(peers);  // the highest peer books the group
        __syncwarp();
        if (valid) key[r] = d | ((old + (uint32_t)__popc(peers & below)) << 8);
    }
}

__global__ void __launch_bounds__(kBinThreads, 5) k_hs_bins(const uint32_t *__restrict__ binbase, const SortEnt *__restrict__ p1_all,
                                                         SortEnt *__restrict__ ent_all, int64_t stride, int total_bins,
                                                         uint32_t *__restrict__ next_unit) {
    __shared__ BinSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (;;) {
        __syncthreads();  // (the previous bin is done with sm)
        if (tid == 0) sm.unit = (int)atomicAdd(next_unit, 1u);
        for (int i = tid; i < kBinWarps * 256; i += kBinThreads) (&sm.wc[0][0])[i] = 0;
        sm.run[tid] = 0;
        __syncthreads();
        const int unit = sm.unit;
        if (unit >= total_bins) return;
        const int stream = unit / kHsBins, bin = unit - stream * kHsBins;
        const uint32_t lo = binbase[(int64_t)stream * (kHsBins + 1) + bin], hi = binbase[(int64_t)stream * (kHsBins + 1) + bin + 1];
        const int S = (int)(hi - lo);
        if (S == 0) continue;
        const uint2 *src = reinterpret_cast<const uint2 *>(p1_all + (int64_t)stream * stride + lo);
        SortEnt *ent = ent_all + (int64_t)stream * stride + lo;
        asm volatile("" : "+l"(src), "+l"(ent));  // (kept in registers instead of re-derived from the parameters)
        const bool single = S <= kBinChunk;
        uint2 pq[kBinRows];
        uint32_t key[kBinRows];
        // digit totals -> sm.run[d]
        if (single) {
            bin_rank_chunk(sm, src, S, 0, pq, key);
            __syncthreads();
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < kBinWarps; w++) t += sm.wc[w][tid];
            sm.run[tid] = t;
        } else {
            for (int k0 = 0; k0 < S; k0 += kBinThreads * 4) {
                uint2 q[4];
#pragma unroll
                for (int r = 0; r < 4; r++) { const int k = k0 + r * kBinThreads + tid; q[r] = k < S ? __ldg(src + k) : make_uint2(0u, 0u); }
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const bool valid = k0 + r * kBinThreads + tid < S;
                    const uint32_t d = sort_digit2(hash_quad(q[r].y));
                    const uint32_t d0 = __shfl_sync(kAll, d, 0);
                    if (__all_sync(kAll, valid && d == d0)) { if (lane == 0) atomicAdd(&sm.run[d0], 32u); }  // a run: one add for the warp
                    else if (valid) atomicAdd(&sm.run[d], 1u);
                }
            }
        }
        __syncthreads();
        {   // exclusive scan of the 256 totals (one per thread)
            const uint32_t t = sm.run[tid];
            uint32_t incl = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(kAll, incl, o); if (lane >= o) incl += u; }
            if (lane == 31) sm.warp_sum[wid] = incl;
            __syncthreads();
            uint32_t off = 0;
#pragma unroll
            for (int w = 0; w < kBinWarps; w++) if (w < wid) off += sm.warp_sum[w];
            const uint32_t start = off + incl - t;
            sm.run[tid] = start;
            sm.first[tid] = start;
        }
        for (int c0 = 0; c0 < S; c0 += kBinChunk) {
            if (!single) {
                __syncthreads();  // (the previous chunk's scatter has read wc)
                for (int i = tid; i < kBinWarps * 256; i += kBinThreads) (&sm.wc[0][0])[i] = 0;
                __syncthreads();
                bin_rank_chunk(sm, src, S, c0, pq, key);
            }
            __syncthreads();
            {   // the warps' shares of digit `tid` follow one another from the digit's next free slot
                uint32_t run = sm.run[tid];
#pragma unroll
                for (int w = 0; w < kBinWarps; w++) { const uint32_t c = sm.wc[w][tid]; sm.wc[w][tid] = run; run += c; }
                sm.run[tid] = run;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < kBinRows; r++) {
                if (key[r] == 0xFFFFFFFFu) continue;
                const uint32_t d = key[r] & 0xFFu, idx = sm.wc[wid][d] + (key[r] >> 8), q = pq[r].y;
                uint32_t a = pq[r].x;
                if (quad_is_run(q)) a |= kEntRun;
                if (idx == sm.first[d]) a |= kEntFirst;
                *reinterpret_cast<uint2 *>(ent + idx) = make_uint2(a, q);
            }
        }
    }
}

// ------------------------------------------------------------------ k_bprobe
// longest_match from a cold start for every chain member, one thread per entry of the sorted array (bucket_probe in
// deflate_core.cuh): the lanes of a warp hold neighbouring targets of one bucket and step down the entries below them
// together, so the candidate loads are coalesced 8-byte reads of one contiguous piece of ent[].  Results are scattered
// by position: mlo[p] = the entry for chain budget 4096, with kProbeHiDiff set where the budget-1024 entry differs (then
// it is written to mhi[p]; otherwise mhi[p] is left alone).
#ifndef CCT_BPROBE_THREADS
#define CCT_BPROBE_THREADS 128
#endif
#ifndef CCT_BPROBE_CTAS
#define CCT_BPROBE_CTAS 64
#endif
constexpr int kBprobeThreads = CCT_BPROBE_THREADS;
constexpr int kBprobeCtas = CCT_BPROBE_CTAS;  // CTAs per stream (each strides over the stream's sorted entries)
__global__ void __launch_bounds__(kBprobeThreads, 1536 / kBprobeThreads) k_bprobe(const uint8_t *__restrict__ in, int64_t in_stride,
                                                           const int32_t *__restrict__ in_len, const SortEnt *__restrict__ ent_all,
                                                           const uint32_t *__restrict__ runs_all, uint32_t *__restrict__ mlo_all,
                                                           uint32_t *__restrict__ mhi_all, uint32_t *__restrict__ tails_all, int64_t stride) {
    const int stream = blockIdx.y, n = in_len[stream];
    const int members = n >= 4 ? n - 3 : 0;
    const int64_t o = (int64_t)stream * stride;
    uint32_t *mlo = mlo_all + o, *mhi = mhi_all + o, *tails = tails_all + o;
    if (blockIdx.x == 0 && threadIdx.x < 3) {  // positions that are in no chain: 0 and the last two
        const int p = threadIdx.x == 0 ? 0 : n - 3 + (int)threadIdx.x;
        if (p >= 0 && p < n && (threadIdx.x == 0 || p >= 1)) mlo[p] = 0;
    }
    const uint8_t *w_s = in + (int64_t)stream * in_stride;
    const SortEnt *ent_s = ent_all + o;
    const uint32_t *runs_s = runs_all + o;
    // opaque to the compiler: it would otherwise re-derive the three bases from the kernel parameters (six instructions)
    // at every use inside the walks instead of keeping them
    asm volatile("" : "+l"(w_s), "+l"(ent_s), "+l"(runs_s), "+l"(mlo), "+l"(mhi), "+l"(tails));
    BucketView v{w_s, ent_s, runs_s, n};
    // (the walks are warp-wide: all 32 lanes go round together, lanes past the end idle)
    const int lane = threadIdx.x & 31;
    for (int i0 = blockIdx.x * kBprobeThreads + (threadIdx.x & ~31); i0 < members; i0 += gridDim.x * kBprobeThreads) {
        const int i = i0 + lane;
        const bool valid = i < members;
        const SortEnt me = valid ? load_ent(v.ent + i) : SortEnt{kEntFirst, 0u};
        uint32_t lo, hi;
        const bool settled = bucket_probe_quick(v, i, me, valid, true, &lo);
        hi = lo;
        if (__any_sync(0xffffffffu, !settled)) {
            uint32_t glo, ghi;
            bucket_probe(v, i, me, !settled, &glo, &ghi);
            if (!settled) { lo = glo; hi = ghi; }
        }
        if (valid) {
            const int p = ent_pos(me);
            if (lo >= kProbeTail) {  // a member of a byte run: k_run_tails writes the entries; the start announces the run
                if (lo == kProbeTailLeader) tails[1u + atomicAdd(tails, 1u)] = (uint32_t)i;
            } else {
                if (lo != hi) { mhi[p] = hi; lo |= kProbeHiDiff; }
                mlo[p] = lo;
            }
        }
    }
}

// ------------------------------------------------------------------ k_run_tails
// The chain members of byte runs (deflate_core.cuh, "byte runs, settled once per run"): the starts announced by k_bprobe,
// a lane per run, 32 neighbouring runs per warp.  The lanes walk the run groups below their runs together (run_collect:
// the start's table entry, the few groups that can give the run's tail a longer match than the run itself); then, run by
// run, the tail targets -- consecutive entries of the sorted array, consecutive positions -- read their entries off
// those records, 32 at a time.  A run the walk cannot summarise goes through the general walk.
#ifndef CCT_TAIL_CTAS
#define CCT_TAIL_CTAS 9
#endif
constexpr int kTailWarps = 4, kTailCtas = 8;
__global__ void __launch_bounds__(32 * kTailWarps, CCT_TAIL_CTAS) k_run_tails(const uint8_t *__restrict__ in, int64_t in_stride,
                                                              const int32_t *__restrict__ in_len, const SortEnt *__restrict__ ent_all,
                                                              const uint32_t *__restrict__ runs_all, uint32_t *__restrict__ mlo_all,
                                                              uint32_t *__restrict__ mhi_all, const uint32_t *__restrict__ tails_all,
                                                              int64_t stride) {
    __shared__ TailGroup recs[kTailWarps][kTailRecords][32];
    const int stream = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t o = (int64_t)stream * stride;
    const uint32_t *tails = tails_all + o;
    const int cnt = (int)tails[0];
    if ((int)blockIdx.x * kTailWarps * 32 >= cnt) return;
    uint32_t *mlo = mlo_all + o, *mhi = mhi_all + o;
    BucketView v{in + (int64_t)stream * in_stride, ent_all + o, runs_all + o, in_len[stream]};
    auto put = [&](int p, uint32_t lo, uint32_t hi) {
        if (lo != hi) { mhi[p] = hi; lo |= kProbeHiDiff; }
        mlo[p] = lo;
    };
    for (int base = (blockIdx.x * kTailWarps + wid) * 32; base < cnt; base += gridDim.x * kTailWarps * 32) {
        const bool have = base + lane < cnt;
        const int i0 = have ? (int)tails[1 + base + lane] : 0;
        uint32_t slo, shi;
        const int nrec = run_collect(v, i0, have, &recs[wid][0][lane], 32, &slo, &shi);
        int s0 = 0, len = 0;
        if (have) {
            s0 = ent_pos(load_ent(v.ent + i0));
            len = (int)(load_u32(v.runs + s0) & 0xFFFFu) + 1;
            if (nrec >= 0) put(s0, slo, shi);
        }
        __syncwarp();
        const int nruns = min(32, cnt - base);
        for (int l = 0; l < nruns; l++) {  // (everything below is the same for all lanes: the loops are warp-wide)
            const int li0 = __shfl_sync(kAll, i0, l), ls0 = __shfl_sync(kAll, s0, l), llen = __shfl_sync(kAll, len, l);
            const int lnrec = __shfl_sync(kAll, nrec, l);
            const int jmin = max(1, llen - (kMaxMatch - 1)), jmax = llen - kMinMatch;  // the tail: R = len - j = 257 .. 3
            uint32_t lo, hi;
            if (lnrec >= 0) {
                for (int j = jmin + lane; j <= jmax; j += 32) {
                    tail_entry(&recs[wid][0][l], 32, lnrec, llen - j, &lo, &hi);
                    put(ls0 + j, lo, hi);
                }
            } else {
                for (int jb = jmin - 32; jb <= jmax; jb += 32) {  // (the first round is the start's, in lane 0)
                    const int j = jb < jmin ? 0 : jb + lane;
                    const bool valid = jb < jmin ? lane == 0 : j <= jmax;
                    const SortEnt me = valid ? load_ent(v.ent + li0 + j) : SortEnt{kEntFirst, 0u};
                    bucket_probe(v, li0 + j, me, valid, &lo, &hi);
                    if (valid) put(ls0 + j, lo, hi);
                }
            }
        }
        __syncwarp();  // (the records are overwritten by the next 32 runs)
    }
}

// ------------------------------------------------------------------ parse: k_next / k_walk_spec / k_walk_stitch / k_symscan
// deflate_slow's lazy state machine, re-indexed by symbol start (deflate_core.cuh):
//   k_next         one thread per position: the symbol a walk standing there emits, from the probe table alone
//                  (coalesced; kSymSlow where the step depends on the walk's history)
//   k_walk_spec    one thread per kWalkWin-position window walks it speculatively from a cold start: one 4-byte
//                  entry per symbol, symbol starts collected in a bit mask (32 positions per word)
//   k_walk_stitch  one thread per stream joins the windows: it re-walks only until the true path meets the
//                  speculative one (typically a few symbols), then adopts the window's recorded exit state
//   k_symscan      CTA per stream: prefix sum over the mask -> dense symbol array, block boundaries
__global__ void __launch_bounds__(256) k_next(const uint8_t *__restrict__ in, int64_t in_stride, const int32_t *__restrict__ in_len,
                                              const uint32_t *__restrict__ mlo, const uint32_t *__restrict__ mhi,
                                              uint32_t *__restrict__ rec, int64_t stride) {
    // four positions per thread: 16-byte loads of the table entries and stores of the symbols (strides are multiples
    // of 16 elements, so a quad never straddles a stream); the entry after the quad is the next thread's first
    const int stream = blockIdx.y;
    const int n = in_len[stream];
    const int64_t o = (int64_t)stream * stride;
    const uint8_t *w = in + (int64_t)stream * in_stride;
    const uint32_t *lo = mlo + o, *hi = mhi + o;
    for (int p0 = (blockIdx.x * 256 + threadIdx.x) * 4; p0 < n; p0 += gridDim.x * 1024) {
        const uint4 e4 = *reinterpret_cast<const uint4 *>(lo + p0);
        const uint32_t f[5] = {e4.x, e4.y, e4.z, e4.w, p0 + 4 < n ? lo[p0 + 4] : 0u};  // entries with the kProbeHiDiff flag
        uint32_t e[5];
#pragma unroll
        for (int i = 0; i < 5; i++) e[i] = f[i] & ~kProbeHiDiff;
        const uint32_t w4 = *reinterpret_cast<const uint32_t *>(w + p0);  // payload rows are padded to 16 bytes
        uint32_t out[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int p = p0 + i, look = n - p;
            const uint32_t lit = sym_pack((w4 >> (8 * i)) & 0xFFu, 0);
            uint32_t r = lit;  // walk_precompute (deflate_core.cuh) with the entries held in registers
            if (look >= kMinMatch) {
                const uint32_t e0 = e[i];
                const int l0 = mtab_len(e0), d0 = mtab_dist(e0);
                if (e0 != 0 && !(l0 == kMinMatch && d0 > kTooFar)) {
                    r = sym_pack(l0 - kMinMatch, d0);
                    const int look1 = look - 1;
                    if (look1 >= kMinMatch) {
                        const uint32_t e1 = (l0 >= kGoodLen && (f[i + 1] & kProbeHiDiff)) ? hi[p + 1] : e[i + 1];
                        if (!(e1 == 0 || l0 >= kMaxMatch || l0 >= look1)) {
                            if (mtab_len(e1) > l0) r = (l0 >= kGoodLen && e[i + 1] != e1) ? kSymSlow : lit;
                        }
                    }
                }
            }
            out[i] = r;
        }
        *reinterpret_cast<uint4 *>(rec + o + p0) = make_uint4(out[0], out[1], out[2], out[3]);
    }
}

__global__ void __launch_bounds__(128) k_walk_spec(const uint8_t *__restrict__ in, int64_t in_stride,
                                                   const int32_t *__restrict__ in_len,
                                                   const uint32_t *__restrict__ mlo, const uint32_t *__restrict__ mhi, uint32_t *__restrict__ mask,
                                                   int64_t mask_stride, uint32_t *__restrict__ rec, Walk *__restrict__ exits, int64_t stride,
                                                   int max_win) {
    const int stream = blockIdx.y;
    const int k = blockIdx.x * 128 + threadIdx.x;
    const int n = in_len[stream];
    const int ws = k * kWalkWin;
    if (ws >= n) return;
    const int64_t o = (int64_t)stream * stride;
    const uint8_t *w_s = in + (int64_t)stream * in_stride;
    const uint32_t *lo_s = mlo + o, *hi_s = mhi + o;
    uint32_t *rec_s = rec + o, *mask_s = mask + (int64_t)stream * mask_stride;
    // opaque to the compiler, which would otherwise re-derive the bases from the kernel parameters inside the walk
    asm volatile("" : "+l"(w_s), "+l"(lo_s), "+l"(hi_s), "+l"(rec_s), "+l"(mask_s));
    LzView v{w_s, n};
    exits[(int64_t)stream * max_win + k] = walk_window(v, ProbeTab{lo_s, hi_s}, ws, min(ws + kWalkWin, n), mask_s, rec_s);
}

__global__ void __launch_bounds__(128) k_walk_stitch(const uint8_t *__restrict__ in, int64_t in_stride,
                                                    const int32_t *__restrict__ in_len,
                                                    const uint32_t *__restrict__ mlo, const uint32_t *__restrict__ mhi, uint32_t *__restrict__ mask,
                                                    int64_t mask_stride, uint32_t *__restrict__ rec, const Walk *__restrict__ exits, int64_t stride,
                                                    int max_win, int B) {
    // one warp per stream, lane 0 walks: streams diverge completely, so sharing a warp would serialise them
    const int stream = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (stream >= B || (threadIdx.x & 31)) return;
    const int64_t o = (int64_t)stream * stride;
    LzView v{in + (int64_t)stream * in_stride, in_len[stream]};
    walk_stitch(v, ProbeTab{mlo + o, mhi + o}, kWalkWin, mask + (int64_t)stream * mask_stride, rec + o, exits + (int64_t)stream * max_win);
}

constexpr int kScanThreads = 256;
// Symbol starts per tile of kScanThreads mask words (8192 positions): a tile of k_symscan then knows where its symbols
// go without waiting for the tiles before it.
__global__ void __launch_bounds__(kScanThreads) k_symcount(const int32_t *__restrict__ in_len, const uint32_t *__restrict__ mask,
                                                           int64_t mask_stride, int32_t *__restrict__ tile_cnt, int64_t cnt_stride) {
    const int stream = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const int n_words = (in_len[stream] + 31) >> 5;
    if (tile * kScanThreads >= n_words) return;
    const int wi = tile * kScanThreads + tid;
    int c = wi < n_words ? __popc(mask[(int64_t)stream * mask_stride + wi]) : 0;
    __shared__ int warp_sum[kScanThreads / 32];
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((tid & 31) == 0) warp_sum[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < kScanThreads / 32; w++) t += warp_sum[w];
        tile_cnt[(int64_t)stream * cnt_stride + tile] = t;
    }
}

// One CTA per tile: prefix sum over the mask -> the tile's piece of the dense symbol array and the block boundaries in
// it; the CTA of the stream's last tile also closes the stream (symbol count, flushed blocks).
__global__ void __launch_bounds__(kScanThreads) k_symscan(const int32_t *__restrict__ in_len, const uint32_t *__restrict__ mask,
                                                          int64_t mask_stride, const uint32_t *__restrict__ rec, uint32_t *__restrict__ syms,
                                                          int64_t stride, int32_t *__restrict__ blk_end, int max_blocks,
                                                          int32_t *__restrict__ nsym, int32_t *__restrict__ nflush,
                                                          const int32_t *__restrict__ tile_cnt, int64_t cnt_stride) {
    const int stream = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = in_len[stream];
    const int n_words = (n + 31) >> 5;
    const int base = tile * kScanThreads;
    if (base >= n_words) {
        if (tile == 0 && tid == 0) {  // an empty stream
            nsym[stream] = 0;
            nflush[stream] = finish_blocks(0, 0u, n, blk_end + (int64_t)stream * max_blocks);
        }
        return;
    }
    const int64_t o = (int64_t)stream * stride;
    const uint32_t *mk = mask + (int64_t)stream * mask_stride;
    const uint32_t *rc = rec + o;
    uint32_t *out = syms + o;
    int32_t *be = blk_end + (int64_t)stream * max_blocks;
    asm volatile("" : "+l"(mk), "+l"(rc), "+l"(out), "+l"(be));  // (kept in registers instead of re-derived from the parameters)
    __shared__ int warp_sum[kScanThreads / 32], warp_carry[kScanThreads / 32];
    // symbols before this tile
    int before = 0;
    for (int t = tid; t < tile; t += kScanThreads) before += tile_cnt[(int64_t)stream * cnt_stride + t];
    for (int d = 16; d; d >>= 1) before += __shfl_xor_sync(0xffffffffu, before, d);
    if (lane == 0) warp_carry[wid] = before;
    const int wi = base + tid;  // one mask word (32 positions) per thread
    const uint32_t bits = wi < n_words ? mk[wi] : 0u;
    const int cnt = __popc(bits);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sum[wid] = incl;
    __syncthreads();
    int carry = 0, woff = 0, total = 0;
    for (int w = 0; w < kScanThreads / 32; w++) { carry += warp_carry[w]; if (w < wid) woff += warp_sum[w]; total += warp_sum[w]; }
    const int idx = carry + woff + incl - cnt;
    // compaction, a word at a time by the whole warp: lane l looks at position 32 j + l of word j, so the symbol
    // loads and the dense stores are coalesced
    const int wbase = base + (wid << 5);
#pragma unroll 8
    for (int j = 0; j < 32; j++) {
        const uint32_t bj = __shfl_sync(0xffffffffu, bits, j);
        const int ij = __shfl_sync(0xffffffffu, idx, j);
        if ((bj >> lane) & 1u) {
            const int p = ((wbase + j) << 5) + lane;
            const int at = ij + __popc(bj & ((1u << lane) - 1u));
            if (at > 0 && at % kSymEnd == 0) be[at / kSymEnd - 1] = p;
            out[at] = rc[p] & ~kSymSlow;
        }
    }
    if (tid == 0 && base + kScanThreads >= n_words) {  // the last tile: close the stream
        const int all = carry + total;
        uint32_t last = 0;
        if (all) {  // the last symbol of the stream: at the highest position marked (read from rec, not from another CTA's output)
            int w = n_words - 1;
            while (mk[w] == 0u) w--;
            last = rc[(w << 5) + 31 - __clz((int)mk[w])] & ~kSymSlow;
        }
        nsym[stream] = all;
        nflush[stream] = finish_blocks(all, last, n, be);
    }
}

// ------------------------------------------------------------------ k_adler
__global__ void __launch_bounds__(256) k_adler(const uint8_t *__restrict__ in, int64_t in_stride,
                                               const int32_t *__restrict__ in_len, uint32_t *__restrict__ adler) {
    const int stream = blockIdx.x, tid = threadIdx.x;
    const int n = in_len[stream];
    const uint8_t *w = in + (int64_t)stream * in_stride;
    // A = 1 + sum d_i ; B = n + sum d_i * (n - i)   (mod 65521)
    unsigned long long sa = 0, sb = 0;
    adler_sums(w, n, tid, 256, sa, sb);
    __shared__ unsigned long long ra[8], rb[8];
    for (int o = 16; o; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
    if ((tid & 31) == 0) { ra[tid >> 5] = sa; rb[tid >> 5] = sb % 65521ull; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long A = 1, Bv = (unsigned long long)n;
        for (int k = 0; k < 8; k++) { A += ra[k]; Bv += rb[k]; }
        adler[stream] = (uint32_t)(((Bv % 65521ull) << 16) | (A % 65521ull));
    }
}

// ------------------------------------------------------------------ k_block_build
// One warp per 16383-symbol block.  All lanes histogram the block's symbols (16-byte loads, shared-memory atomics);
// lane 0 then runs zlib's tree construction verbatim (its tie-breaking lives in the heap order, so it cannot be
// replaced by a parallel Huffman construction).  That part is a chain of dependent shared-memory accesses, so the
// kernel's speed is the number of blocks resident per SM: the histogram counters share storage with the priority
// queue and the order array, and the block descriptor is assembled in their place once the trees are built.
#ifndef CCT_BB_PER_WARP
#define CCT_BB_PER_WARP 3
#endif
constexpr int kBBPerWarp = CCT_BB_PER_WARP;  // blocks of one stream per warp: their trees are built side by side, a lane each
struct BlockBuildSmem {
    BlockWork wk[kBBPerWarp];
};
static_assert(offsetof(TreeScratch, pq) == 0 && offsetof(TreeScratch, heap) == sizeof(TreeScratch::pq) &&
              sizeof(BlockDesc) <= sizeof(TreeScratch::pq) + sizeof(TreeScratch::heap), "descriptor must fit in sc.pq + sc.heap");
static_assert((kLCodes + 2) * sizeof(uint32_t) <= sizeof(TreeScratch::pq) && (kDCodes + 2) * sizeof(uint32_t) <= sizeof(TreeScratch::heap),
              "histograms must fit in the queue and the order array");

__global__ void __launch_bounds__(32) k_block_build(const uint32_t *__restrict__ syms, int64_t stride,
                                                    const int32_t *__restrict__ blk_end, int max_blocks,
                                                    const int32_t *__restrict__ nsym, const int32_t *__restrict__ nflush,
                                                    const int32_t *__restrict__ in_len, BlockDesc *__restrict__ desc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BlockBuildSmem &sm = *reinterpret_cast<BlockBuildSmem *>(smem_raw);
    const int stream = blockIdx.y, blk0 = blockIdx.x * kBBPerWarp, lane = threadIdx.x;
    const int nf = nflush[stream];
    if (blk0 > nf) return;
    const int nblk = min(kBBPerWarp, nf + 1 - blk0);  // blocks of this warp
    uint32_t *z = reinterpret_cast<uint32_t *>(&sm);
    for (int i = lane; i < (int)(sizeof(BlockBuildSmem) / 4); i += 32) z[i] = 0;
    __syncwarp();
    const int32_t *be = blk_end + (int64_t)stream * max_blocks;
    const int n_in = in_len[stream], n_sym = nsym[stream];
    // the frequencies of every block, by all lanes
    for (int k = 0; k < nblk; k++) {
        BlockWork &wk = sm.wk[k];
        const int blk = blk0 + k;
        uint32_t *lh = wk.sc.pq, *dh = reinterpret_cast<uint32_t *>(wk.sc.heap);
        const uint32_t sym_begin = (uint32_t)blk * kSymEnd;
        const uint32_t ns = blk < nf ? (uint32_t)kSymEnd : (uint32_t)(n_sym - nf * kSymEnd);
        const uint32_t *sy = syms + (int64_t)stream * stride + sym_begin;  // stream base is 64-byte aligned
        auto tally = [&](uint32_t s) {
            if (sym_dist(s) == 0) atomicAdd(&lh[sym_lc(s)], 1u);
            else { atomicAdd(&lh[257 + len_code(sym_lc(s))], 1u); atomicAdd(&dh[dist_code(sym_dist(s) - 1)], 1u); }
        };
        const uint32_t head = min(ns, (4u - (sym_begin & 3u)) & 3u);  // symbols before the first 16-byte boundary
        if ((uint32_t)lane < head) tally(sy[lane]);
        const uint4 *v4 = reinterpret_cast<const uint4 *>(sy + head);
        const uint32_t nv = (ns - head) >> 2;
#pragma unroll 4
        for (uint32_t i = lane; i < nv; i += 32) {
            const uint4 q = __ldg(v4 + i);
            tally(q.x); tally(q.y); tally(q.z); tally(q.w);
        }
        for (uint32_t i = head + 4 * nv + lane; i < ns; i += 32) tally(sy[i]);
        __syncwarp();
        for (int i = lane; i < kLCodes; i += 32) wk.lt.freq[i] = (uint16_t)lh[i];
        if (lane < kDCodes) wk.dt.freq[lane] = (uint16_t)dh[lane];
        __syncwarp();
    }
    // the trees: lane k builds block k's
    if (lane < nblk) {
        BlockWork &wk = sm.wk[lane];
        const int blk = blk0 + lane;
        const uint32_t sym_begin = (uint32_t)blk * kSymEnd;
        const uint32_t ns = blk < nf ? (uint32_t)kSymEnd : (uint32_t)(n_sym - nf * kSymEnd);
        const uint32_t pos_begin = blk == 0 ? 0u : (uint32_t)be[blk - 1];
        const uint32_t pos_end = blk < nf ? (uint32_t)be[blk] : (uint32_t)n_in;
        const uint32_t *sy = syms + (int64_t)stream * stride + sym_begin;
        BlockDesc *d = reinterpret_cast<BlockDesc *>(wk.sc.pq);
        wk.lt.freq[256] = 1;
        // loop top of the iteration that flushes the block: the start of its last symbol + 1, or the end of the input
        const int64_t top = blk < nf ? (int64_t)pos_end - sym_advance(sy[ns - 1]) + 1 : (int64_t)n_in;
        build_block(wk, d, pos_end - pos_begin, blk == nf, (int64_t)pos_begin >= window_base_at(top, n_in));
        d->sym_begin = sym_begin; d->nsym = ns; d->pos_begin = pos_begin;
    }
    __syncwarp();
    for (int k = 0; k < nblk; k++) {
        BlockWork &wk = sm.wk[k];
        BlockDesc *d = reinterpret_cast<BlockDesc *>(wk.sc.pq);
        fill_block_tables(wk, d, lane, 32);
        __syncwarp();
        uint32_t *dst = reinterpret_cast<uint32_t *>(desc + (int64_t)stream * max_blocks + blk0 + k);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(d);
        for (int i = lane; i < (int)(sizeof(BlockDesc) / 4); i += 32) dst[i] = src[i];
    }
}

// ------------------------------------------------------------------ k_layout
__global__ void k_layout(BlockDesc *__restrict__ desc, int max_blocks, const int32_t *__restrict__ nflush,
                         const uint32_t *__restrict__ adler, uint8_t *__restrict__ out, int64_t out_stride,
                         int32_t *__restrict__ out_len, int B) {
    int stream = blockIdx.x * blockDim.x + threadIdx.x;
    if (stream >= B) return;
    BlockDesc *d = desc + (int64_t)stream * max_blocks;
    uint64_t bit = 16;
    int nblk = nflush[stream] + 1;
    for (int b = 0; b < nblk; b++) {
        d[b].bit_start = bit;
        if (d[b].type == 0) {
            bit += 3;
            bit = (bit + 7) & ~7ull;
            bit += 32 + 8ull * d[b].stored_len;
        } else {
            bit += d[b].hdr_bits + d[b].body_bits;
        }
    }
    bit = (bit + 7) & ~7ull;
    uint8_t *o = out + (int64_t)stream * out_stride;
    int64_t len = (int64_t)(bit >> 3);
    o[0] = 0x78; o[1] = 0xDA;  // CMF/FLG for wbits 15, level 9
    uint32_t ad = adler[stream];
    o[len] = (uint8_t)(ad >> 24); o[len + 1] = (uint8_t)(ad >> 16); o[len + 2] = (uint8_t)(ad >> 8); o[len + 3] = (uint8_t)ad;
    out_len[stream] = (int32_t)(len + 4);
}

// ------------------------------------------------------------------ k_pack
__device__ __forceinline__ void or_bits(uint32_t *out32, uint64_t bitpos, uint32_t value, int nb) {
    uint64_t word = bitpos >> 5;
    int off = (int)(bitpos & 31);
    atomicOr(out32 + word, value << off);
    if (off + nb > 32) atomicOr(out32 + word + 1, value >> (32 - off));
}

struct BitSink {  // per-thread streaming writer: first and last word are OR-ed, interior words stored
    uint32_t *wp; unsigned long long acc; int fill; bool first;
    __device__ __forceinline__ void put(uint32_t v, int nb) {
        acc |= (unsigned long long)v << fill;
        fill += nb;
        if (fill >= 32) {
            uint32_t word = (uint32_t)acc;
            if (first) { atomicOr(wp, word); first = false; } else *wp = word;
            wp++; acc >>= 32; fill -= 32;
        }
    }
    __device__ __forceinline__ void finish() { if (fill > 0) atomicOr(wp, (uint32_t)acc); }
};

constexpr int kPackThreads = 256;
__global__ void __launch_bounds__(kPackThreads) k_pack(const uint8_t *__restrict__ in, int64_t in_stride,
                                                       const BlockDesc *__restrict__ desc, int max_blocks,
                                                       const int32_t *__restrict__ nflush, const uint32_t *__restrict__ syms,
                                                       int64_t stride, uint8_t *__restrict__ out, int64_t out_stride) {
    const int stream = blockIdx.y, blk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (blk > nflush[stream]) return;
    const BlockDesc &d = desc[(int64_t)stream * max_blocks + blk];
    uint32_t *out32 = reinterpret_cast<uint32_t *>(out + (int64_t)stream * out_stride);
    const uint64_t bit_start = d.bit_start;
    const int hdr_bits = (int)d.hdr_bits;
    if (wid == 0) {
        for (int i = lane; i * 32 < hdr_bits; i += 32) {
            int nb = hdr_bits - i * 32; if (nb > 32) nb = 32;
            or_bits(out32, bit_start + 32ull * i, d.hdr[i], nb);
        }
    }
    if (d.type == 0) {  // stored block: LEN, ~LEN, raw bytes at the next byte boundary
        uint8_t *o = out + (int64_t)stream * out_stride + (((bit_start + 3) + 7) >> 3);
        const uint32_t len = d.stored_len;
        if (tid == 0) { o[0] = (uint8_t)len; o[1] = (uint8_t)(len >> 8); o[2] = (uint8_t)~len; o[3] = (uint8_t)(~len >> 8); }
        const uint8_t *src = in + (int64_t)stream * in_stride + d.pos_begin;
        for (uint32_t i = tid; i < len; i += kPackThreads) o[4 + i] = src[i];
        return;
    }
    __shared__ uint16_t lcode[kLCodes + 2], dcode[kDCodes + 2];
    __shared__ uint8_t llen[kLCodes + 2], dlen[kDCodes + 2];
    __shared__ uint32_t warp_sum[kPackThreads / 32];
    // A match costs two more look-ups than a literal, not a chain of code arithmetic (the lanes of a warp hold both
    // kinds and run both paths): per length lc the bits of its length code with the extra bits appended and their
    // number; per distance code the number of bits of code and extra bits together.
    __shared__ uint32_t mbits[256];
    __shared__ uint8_t mnb[256], dnb[kDCodes + 2];
    for (int i = tid; i < kLCodes + 2; i += kPackThreads) { lcode[i] = d.lcode[i]; llen[i] = d.llen[i]; }
    if (tid < kDCodes + 2) { dcode[tid] = d.dcode[tid]; dlen[tid] = d.dlen[tid]; dnb[tid] = (uint8_t)(d.dlen[tid] + (tid < kDCodes ? dist_extra_bits(tid) : 0)); }
    static_assert(kPackThreads == 256, "one length per thread");
    {
        const int code = len_code(tid), extra = len_extra_bits(code), nb = d.llen[code + 257];
        mbits[tid] = (uint32_t)d.lcode[code + 257] | (((uint32_t)tid & ((1u << extra) - 1u)) << nb);
        mnb[tid] = (uint8_t)(nb + extra);
    }
    __syncthreads();
    const int ns = (int)d.nsym;
    // A thread's symbols are consecutive (its bits are, too), so its loads stride the warp.  The runs are cut on
    // 16-byte boundaries of the symbol array -- `mis` symbols before the block's first one is such a boundary -- so a
    // thread fetches its run in 16-byte loads, two at a time (a quarter of the requests of 4-byte loads; the first and
    // last quad of the block may reach outside it: those symbols are skipped, the array is padded).
    const uint32_t *sy = syms + (int64_t)stream * stride + d.sym_begin;
    const int mis = (int)(d.sym_begin & 3u);
    const uint4 *quads = reinterpret_cast<const uint4 *>(sy - mis);
    asm volatile("" : "+l"(quads), "+l"(out32));  // (kept in registers: the loops below would re-derive them from the parameters)
    const int nq = (mis + ns + 3) >> 2;                                  // quads covering the block
    const int per = (nq + kPackThreads - 1) / kPackThreads;              // quads per thread
    const int q0 = min(nq, tid * per), q1 = min(nq, q0 + per);
    const int lo = mis, hi = mis + ns;                                   // the block's symbols in quad coordinates
    auto for_each = [&](auto &&fn) {
        int q = q0;
        for (; q + 2 <= q1; q += 2) {
            const uint4 u = __ldg(quads + q), v = __ldg(quads + q + 1);
            const uint32_t sym[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
            if (4 * q >= lo && 4 * q + 8 <= hi) {
#pragma unroll
                for (int j = 0; j < 8; j++) fn(sym[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) if (4 * q + j >= lo && 4 * q + j < hi) fn(sym[j]);
            }
        }
        for (; q < q1; q++) {
            const uint4 u = __ldg(quads + q);
            const uint32_t sym[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; j++) if (4 * q + j >= lo && 4 * q + j < hi) fn(sym[j]);
        }
    };
    uint32_t bits = 0;
    for_each([&](uint32_t sym) {
        const int dist = sym_dist(sym), lc = sym_lc(sym);
        bits += dist ? (uint32_t)mnb[lc] + dnb[dist_code(dist - 1)] : (uint32_t)llen[lc];
    });
    uint32_t incl = bits;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sum[wid] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < wid; w++) woff += warp_sum[w];
    const uint64_t start = bit_start + (uint64_t)hdr_bits + woff + incl - bits;
    if (bits) {
        BitSink sink{out32 + (start >> 5), 0ull, (int)(start & 31), true};
        for_each([&](uint32_t sym) {
            const int dist = sym_dist(sym), lc = sym_lc(sym);
            if (dist == 0) sink.put(lcode[lc], llen[lc]);
            else {  // length code + extra (at most 20 bits), then distance code + extra (at most 28)
                sink.put(mbits[lc], mnb[lc]);
                const int dd = dist - 1, code = dist_code(dd), cl = dlen[code];
                sink.put((uint32_t)dcode[code] | (((uint32_t)dd & ((1u << dist_extra_bits(code)) - 1u)) << cl), dnb[code]);
            }
        });
        sink.finish();
    }
    if (tid == 0)  // END_BLOCK closes the block
        or_bits(out32, bit_start + (uint64_t)hdr_bits + d.body_bits - llen[256], lcode[256], llen[256]);
}

// Per-device launch state: opt-in shared-memory sizes are a property of (function, device), so they are set once per
// device (several devices may be driven from one process through cct_*_host).
static void dev_setup() {
    static std::mutex mu;
    static bool done[64];
    int dev = 0;
    cudaGetDevice(&dev);
    dev = std::min(std::max(dev, 0), 63);
    std::lock_guard<std::mutex> lk(mu);
    if (done[dev]) return;
    cudaFuncSetAttribute(k_block_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BlockBuildSmem));
    done[dev] = true;
}

// run counts, hash sort and probe tables of B streams (the first half of launch_deflate; also the stage tap)
static void launch_match_tables(const uint8_t *in, int64_t in_stride, const int32_t *in_len, int B, int64_t max_len,
                                const DeflateWs &ws, cudaStream_t st) {
    {   // the chunk summaries live in arrays that are filled later
        const int max_chunks = (int)((max_len + kRunChunk - 1) / kRunChunk);
        int4 *edges = reinterpret_cast<int4 *>(ws.rec);
        int2 *enters = reinterpret_cast<int2 *>(ws.mhi);
        dim3 gt((unsigned)((max_chunks + kRunWarps - 1) / kRunWarps), (unsigned)B);
        { CCT_KERNEL(k_run_edges, st);
        k_run_edges<<<gt, 32 * kRunWarps, 0, st>>>(in, in_stride, in_len, edges, max_chunks);
        }
        { CCT_KERNEL(k_run_link, st);
        k_run_link<<<(B + kRunWarps - 1) / kRunWarps, 32 * kRunWarps, 0, st>>>(in_len, edges, enters, max_chunks, B);
        }
        { CCT_KERNEL(k_run_counts, st);
        k_run_counts<<<gt, 32 * kRunWarps, 0, st>>>(in, in_stride, in_len, enters, max_chunks, ws.runs, ws.stride);
        }
    }
    dim3 gtile((unsigned)ws.max_tiles, (unsigned)B);
    { CCT_KERNEL(k_hs_hist, st);
    k_hs_hist<<<gtile, kHsThreads, 0, st>>>(in, in_stride, in_len, ws.thist, ws.max_tiles);
    }
    { CCT_KERNEL(k_hs_scan, st);
    k_hs_scan<<<B, kHsBins, 0, st>>>(in_len, ws.thist, ws.max_tiles, ws.binbase, ws.rec, ws.stride);
    }
    { CCT_KERNEL(k_hs_split, st);
    k_hs_split<<<gtile, kHsThreads, 0, st>>>(in, in_stride, in_len, ws.thist, ws.max_tiles, ws.p1, ws.stride);
    }
    {
        const int64_t total_bins = (int64_t)B * kHsBins;
        int dev = 0, n_sm = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        const int64_t ctas = std::min<int64_t>(total_bins, (int64_t)n_sm * 8);
        { CCT_KERNEL(m_zero_counters, st); launch_zero(ws.counters, 64 * sizeof(uint32_t), st); }
        CCT_KERNEL(k_hs_bins, st);
        k_hs_bins<<<(unsigned)ctas, kBinThreads, 0, st>>>(ws.binbase, ws.p1, ws.ent, ws.stride, (int)total_bins, ws.counters);
    }
    {
        dim3 gp((unsigned)std::min<int64_t>((max_len + kBprobeThreads - 1) / kBprobeThreads, kBprobeCtas), (unsigned)B);
        CCT_KERNEL(k_bprobe, st);
        k_bprobe<<<gp, kBprobeThreads, 0, st>>>(in, in_stride, in_len, ws.ent, ws.runs, ws.mtab, ws.mhi, ws.rec, ws.stride);
    }
    {
        CCT_KERNEL(k_run_tails, st);
        k_run_tails<<<dim3(kTailCtas, (unsigned)B), 32 * kTailWarps, 0, st>>>(in, in_stride, in_len, ws.ent, ws.runs, ws.mtab, ws.mhi,
                                                                              ws.rec, ws.stride);
    }
}

void launch_probe_tables(const uint8_t *in, int64_t in_stride, const int32_t *in_len, int B, int64_t max_len,
                         const DeflateWs &ws, cudaStream_t st) {
    if (B == 0) return;
    launch_match_tables(in, in_stride, in_len, B, max_len, ws, st);
}

void launch_deflate(const uint8_t *in, int64_t in_stride, const int32_t *in_len, int B, int64_t max_len, uint8_t *out,
                    int64_t out_stride, int32_t *out_len, const DeflateWs &ws, cudaStream_t st) {
    if (B == 0) return;
    dev_setup();
    launch_match_tables(in, in_stride, in_len, B, max_len, ws, st);
    dim3 gn((unsigned)std::min<int64_t>((max_len + 1023) / 1024, 512), (unsigned)B);
    { CCT_KERNEL(k_next, st);
    k_next<<<gn, 256, 0, st>>>(in, in_stride, in_len, ws.mtab, ws.mhi, ws.rec, ws.stride);
    }
    dim3 gw((unsigned)((ws.max_win + 127) / 128), (unsigned)B);
    { CCT_KERNEL(k_walk_spec, st);
    k_walk_spec<<<gw, 128, 0, st>>>(in, in_stride, in_len, ws.mtab, ws.mhi, ws.mask, ws.mask_stride, ws.rec, ws.exits,
                                    ws.stride, ws.max_win);
    }
    { CCT_KERNEL(k_walk_stitch, st);
    k_walk_stitch<<<(B + 3) / 4, 128, 0, st>>>(in, in_stride, in_len, ws.mtab, ws.mhi, ws.mask, ws.mask_stride, ws.rec,
                                                ws.exits, ws.stride, ws.max_win, B);
    }
    {   // (the per-tile counts live where the window exits were: the walks are over)
        int32_t *tile_cnt = reinterpret_cast<int32_t *>(ws.exits);
        const int64_t cnt_stride = (int64_t)ws.max_win * (int64_t)(sizeof(Walk) / sizeof(int32_t));
        const dim3 gt((unsigned)((ws.mask_stride + kScanThreads - 1) / kScanThreads), (unsigned)B);
        { CCT_KERNEL(k_symcount, st);
        k_symcount<<<gt, kScanThreads, 0, st>>>(in_len, ws.mask, ws.mask_stride, tile_cnt, cnt_stride);
        }
        CCT_KERNEL(k_symscan, st);
        k_symscan<<<gt, kScanThreads, 0, st>>>(in_len, ws.mask, ws.mask_stride, ws.rec, ws.syms, ws.stride, ws.blk_end, ws.max_blocks,
                                               ws.nsym, ws.nflush, tile_cnt, cnt_stride);
    }
    { CCT_KERNEL(k_adler, st);
    k_adler<<<B, 256, 0, st>>>(in, in_stride, in_len, ws.adler);
    }
    dim3 gb((unsigned)ws.max_blocks, (unsigned)B);
    { CCT_KERNEL(k_block_build, st);
    k_block_build<<<dim3((unsigned)((ws.max_blocks + kBBPerWarp - 1) / kBBPerWarp), (unsigned)B), 32, sizeof(BlockBuildSmem), st>>>(ws.syms, ws.stride, ws.blk_end, ws.max_blocks, ws.nsym, ws.nflush,
                                                          in_len, ws.desc);
    }
    { CCT_KERNEL(k_layout, st);
    k_layout<<<(B + 127) / 128, 128, 0, st>>>(ws.desc, ws.max_blocks, ws.nflush, ws.adler, out, out_stride, out_len, B);
    }
    { CCT_KERNEL(k_pack, st);
    k_pack<<<gb, kPackThreads, 0, st>>>(in, in_stride, ws.desc, ws.max_blocks, ws.nflush, ws.syms, ws.stride, out, out_stride);
    }
}

// ------------------------------------------------------------------ assembly into the caller's arena
__global__ void __launch_bounds__(1024) k_offsets(const int32_t *__restrict__ body_len, int B, int64_t *__restrict__ offsets,
                                                  int first) {
    __shared__ long long warp_sum[32];
    __shared__ long long carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = offsets[first];
    __syncthreads();
    for (int base = 0; base < B; base += 1024) {
        int i = base + tid;
        long long v = i < B ? (long long)kHeaderBytes + body_len[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sum[wid] = incl;
        __syncthreads();
        long long woff = 0, total = 0;
        for (int w = 0; w < 32; w++) { if (w < wid) woff += warp_sum[w]; total += warp_sum[w]; }
        if (i < B) offsets[first + 1 + i] = carry + woff + incl;
        __syncthreads();
        if (tid == 0) carry += total;
        __syncthreads();
    }
}

constexpr int kCopyBytes = 4096;  // bytes per CTA
__global__ void __launch_bounds__(256) k_copy(const uint8_t *__restrict__ body, int64_t body_stride,
                                              const int32_t *__restrict__ body_len, int rows, int cols, uint32_t flags,
                                              uint8_t *__restrict__ arena, int64_t arena_cap,
                                              const int64_t *__restrict__ offsets, int first, uint32_t *__restrict__ status) {
    const int img = blockIdx.y;
    const int len = body_len[img];
    const int64_t lo = offsets[first + img], hi = offsets[first + img + 1];
    if (hi > arena_cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0 && status) status[img] |= 2u;  // CCT_ST_CAPACITY
        return;
    }
    const int64_t start = (int64_t)blockIdx.x * kCopyBytes;
    if (start >= len) return;
    uint8_t *dst = arena + lo;
    if (blockIdx.x == 0 && threadIdx.x < kHeaderBytes) {  // core.py:193-210
        const int t = threadIdx.x;
        uint8_t hb;
        switch (t) {
            case 0: hb = 'p'; break; case 1: hb = 'a'; break; case 2: hb = 'c'; break; case 3: hb = 't'; break;
            case 4: hb = (uint8_t)(rows >> 8); break; case 5: hb = (uint8_t)rows; break;
            case 6: hb = (uint8_t)(cols >> 8); break; case 7: hb = (uint8_t)cols; break;
            case 8: hb = 1; break; case 9: hb = 2; break;
            case 10: hb = (flags & 1u) != 0; break; case 11: hb = (flags & 2u) != 0; break;
            default: hb = (flags & 4u) != 0; break;
        }
        dst[t] = hb;
    }
    const uint8_t *src = body + (int64_t)img * body_stride;
    const int64_t end = min((int64_t)len, start + kCopyBytes);
    for (int64_t i = start + threadIdx.x; i < end; i += 256) dst[kHeaderBytes + i] = src[i];
}

void launch_assemble(const uint8_t *body, int64_t body_stride, const int32_t *body_len, int B, int rows, int cols,
                     uint32_t flags, uint8_t *arena, int64_t arena_cap, int64_t *offsets, int first, uint32_t *status,
                     cudaStream_t st) {
    if (B == 0) return;
    { CCT_KERNEL(k_offsets, st);
    k_offsets<<<1, 1024, 0, st>>>(body_len, B, offsets, first);
    }
    dim3 g((unsigned)((body_stride + kCopyBytes - 1) / kCopyBytes), (unsigned)B);
    { CCT_KERNEL(k_copy, st);
    k_copy<<<g, 256, 0, st>>>(body, body_stride, body_len, rows, cols, flags, arena, arena_cap, offsets, first,
                              status ? status + first : nullptr);
    }
}

}  // namespace cct
