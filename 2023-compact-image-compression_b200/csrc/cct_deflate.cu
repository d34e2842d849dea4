// cct_deflate.cu — bit-exact zlib-1.3 level-9 `compress2` on the GPU, one zlib stream per image payload.
//
// Replaces `zlib.compress(data, level = 9)` (reference src/codec/core.py:340).  Stages (see
// deflate_core.cuh for the decomposition and SURVEY.md Appendix B for the algorithm being matched):
//   k_chains       hash-chain links for every position (CTA per stream, warp per hash class, head table in shared memory)
//   k_run_edges / k_run_link / k_run_recs   equal-byte run lengths around every position; packs the per-position ChainRec
//   k_probe        cold-start longest_match over the first kProbeK chain entries, one thread per position (persistent,
//                  work placed by SM)
//   k_next / k_walk_spec / k_walk_stitch / k_symscan   deflate_slow's lazy state machine as a windowed speculative walk
//   k_adler        Adler-32 of the payload (CTA per stream)
//   k_block_build  per 16383-symbol block: histograms, build_tree x3, block-type choice, header bits
//   k_layout       bit offsets of the blocks of each stream, stream length, zlib header/trailer
//   k_pack         Huffman/extra bits of every symbol, placed by a prefix sum of code lengths
//   k_offsets / k_copy   .cct header + stream -> caller's arena at scanned offsets
#include <cuda_runtime.h>

#include <algorithm>
#include <mutex>

#include "cct_kernels.h"
#include "deflate_core.cuh"

namespace cct {

#ifndef CCT_WALK_WIN
#define CCT_WALK_WIN 1024
#endif
constexpr int kWalkWin = CCT_WALK_WIN;  // positions per speculative parse window
constexpr int kProbeK = 64;  // loop steps of the parallel probe (run groups count as one step)

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
    ws->dist16 = (uint16_t *)take((size_t)B * ws->stride * sizeof(uint16_t));
    ws->recs = (ChainRec *)take((size_t)B * ws->stride * sizeof(ChainRec));
    ws->mtab = (uint32_t *)take((size_t)B * ws->stride * sizeof(uint32_t));
    ws->mhi = (uint32_t *)take((size_t)B * ws->stride * sizeof(uint32_t));
    ws->syms = ws->mtab;  // the dense symbol array reuses the probe table once the walk is over
    ws->rec = (uint32_t *)take((size_t)B * ws->stride * sizeof(uint32_t));
    ws->probe_next = (uint32_t *)take(kMaxSm * sizeof(uint32_t));
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

// ------------------------------------------------------------------ k_chains / k_runs
// k_chains: distance from every position to the previous position with the same 3-byte hash (zlib's prev[]),
// one CTA per stream.  The 32768-entry head table (low 16 bits of the latest position per hash) lives in shared
// memory, split by hash class between the CTA's warps; entries are kept unambiguous by a sweep every 8192
// positions that re-parks everything older than 32767 at "49152 behind".  (A global-memory table with every stream resident was
// measured 3.4x slower: 268 MB of randomly accessed tables do not fit the 126 MB L2.)
// k_runs: per position, the count of equal bytes before and after it (forward and reverse sweep with a carry),
// assembled with the link and the first two bytes into the 8-byte ChainRec the chain walks load.
constexpr int kHeadEntries = 32768;
#ifndef CCT_CHAIN_WARPS_LOG2
#define CCT_CHAIN_WARPS_LOG2 2
#endif
constexpr int kChainWarpsLog2 = CCT_CHAIN_WARPS_LOG2;
constexpr int kChainWarps = 1 << kChainWarpsLog2;  // hash classes = warps per stream
constexpr int kRing = 256;      // pending positions per warp: up to 31 waiting + the 128 a scan step can add

struct ChainSmem {
    uint32_t ring_s[kChainWarps][kRing];
    uint16_t ring_h[kChainWarps][kRing];
};

// One table step for up to 32 pending positions of one hash class (cnt lanes active, positions increasing):
// previous position with the same hash = the nearest lower lane holding it, else the head table; the latest
// position of every hash ends up in the table.  Same-hash lanes inside a batch are found without a full match:
// either every lane has one hash (a run), or all hashes are distinct -- detected by letting every lane store its
// position and read the slot back -- or a few hashes repeat, and the lanes that lost their slot name them.
__device__ __forceinline__ void chain_batch(volatile uint16_t *tab, uint16_t *d16, int lane, int cnt, uint32_t s, uint32_t h) {
    // `tab` is volatile on purpose: the read-back below must see what the OTHER lanes stored to the same slot
    const bool valid = lane < cnt;
    const unsigned vmask = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
    if (!valid) h = 0x8000u | (uint32_t)lane;  // unique key for idle lanes
    int dist = 0;
    uint32_t old = 0;
    if (valid) old = tab[h];
    const uint32_t h0 = __shfl_sync(0xffffffffu, h, 0);
    const bool uniform = __ballot_sync(0xffffffffu, valid && h == h0) == vmask;
    unsigned peers;
    __syncwarp();
    if (uniform) {
        peers = vmask;
        if (valid && (vmask >> lane) == 1u) tab[h] = (uint16_t)s;
    } else {
        if (valid) tab[h] = (uint16_t)s;
        __syncwarp();
        const bool lost = valid && tab[h] != (uint16_t)s;
        unsigned lm = __ballot_sync(0xffffffffu, lost);
        peers = 1u << lane;  // all distinct: every lane owns its slot
        if (lm) {
            // a few hashes occur more than once (a third of the batches, about three hashes each): one ballot per such
            // hash, found through the lanes that lost the slot (__match_any_sync is a software loop over all 32 lanes here)
            do {
                const uint32_t hl = __shfl_sync(0xffffffffu, h, __ffs((int)lm) - 1);
                const unsigned eq = __ballot_sync(0xffffffffu, h == hl);
                if (h == hl) peers = eq;
                lm &= ~eq;
            } while (lm);
            __syncwarp();
            if (valid && (peers >> lane) == 1u) tab[h] = (uint16_t)s;  // the latest position wins
        }
    }
    const unsigned lower = peers & ((1u << lane) - 1u);
    const uint32_t s_low = __shfl_sync(0xffffffffu, s, lower ? 31 - __clz((int)lower) : 0);
    if (valid) {
        uint32_t d = lower ? s - s_low : ((s - old) & 0xFFFFu);
        if (lower ? d >= 32768u : (d < 1u || d >= 32768u)) d = 0;
        dist = (int)d;
        d16[s] = (uint16_t)dist;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(32 * kChainWarps) k_chains(const uint8_t *__restrict__ in, int64_t in_stride,
                                                             const int32_t *__restrict__ in_len, uint16_t *__restrict__ dist16,
                                                             int64_t stride) {
    // One CTA per stream; warp k owns the hashes with (h & 3) == k and a quarter of the head table.  Every warp
    // scans the whole payload, 128 positions per step with four consecutive positions per lane (the lane's word and
    // the next lane's: a funnel shift gives the three bytes of each position), keeps the positions of its own class
    // in a ring, and runs the table step on 32 of them at a time.  Classes never interact, so the four warps need
    // no synchronisation -- four times the warps per SM of a single-warp scan with the same 64 KB of tables.
    // (Sharing the scan instead -- each warp hashes a quarter of every 128 positions and files them into all four
    // rings, two barriers per step for the offsets -- has 40 % fewer instructions and measured 26 % slower: the table
    // steps are latency-bound and the barriers make every warp wait for the slowest one's.)
    extern __shared__ __align__(16) unsigned char chain_raw[];
    uint16_t *head = reinterpret_cast<uint16_t *>(chain_raw);
    ChainSmem &sm = *reinterpret_cast<ChainSmem *>(chain_raw + kHeadEntries * sizeof(uint16_t));
    const int stream = blockIdx.x, lane = threadIdx.x & 31, k = threadIdx.x >> 5;
    const int n = in_len[stream];
    const uint8_t *w = in + (int64_t)stream * in_stride;
    uint16_t *d16 = dist16 + (int64_t)stream * stride;
    uint16_t *tab = head + k * (kHeadEntries / kChainWarps);
    uint32_t *tab32 = reinterpret_cast<uint32_t *>(tab);
    constexpr int kTabWords = kHeadEntries / kChainWarps / 2;
    for (int i = lane; i < kTabWords; i += 32) tab32[i] = 0x40004000u;  // (0 - 49152) & 0xFFFF
    uint32_t *ring_s = sm.ring_s[k];
    uint16_t *ring_h = sm.ring_h[k];
    __syncwarp();
    // positions that are in no chain (position 0 doubles as NIL in zlib; the last two have no hash triple)
    if (k == 0 && lane < 3) { const int s = lane == 0 ? 0 : n - 3 + lane; if (s >= 0 && s < n) d16[s] = 0; }
    const uint32_t *w32 = reinterpret_cast<const uint32_t *>(w);  // stream base is 16-byte aligned
    const int n_words = (n + 3) >> 2;
    auto fetch = [&](int cbase) -> uint32_t {  // the lane's word of the 128-byte chunk starting at cbase
        const int wi = (cbase >> 2) + lane;
        return wi < n_words ? __ldg(w32 + wi) : 0u;
    };
    auto flush = [&](int &rd, int wr, int at_least) {  // table steps while `at_least` positions are pending
        while (wr - rd >= at_least && wr != rd) {
            const int cnt = min(32, wr - rd);
            chain_batch(tab, d16, lane, cnt, ring_s[(rd + lane) & (kRing - 1)], ring_h[(rd + lane) & (kRing - 1)]);
            rd += cnt;
        }
    };
    uint32_t cur = fetch(0), nxt = fetch(128);  // chunks are fetched two ahead
    int wr = 0, rd = 0;  // ring write / read cursors (warp-uniform)
    const unsigned below = (1u << lane) - 1u;
    for (int cbase = 0; cbase < n; cbase += 128) {
        const uint32_t a = cur, nn = nxt;
        cur = nn;
        nxt = fetch(cbase + 256);
        if (cbase && (cbase & 8191) == 0) {
            // sweep: flush what is pending (so the table reflects every position below cbase), then re-park the
            // entries older than 32767 at "49152 behind" to keep the 16-bit positions unambiguous
            flush(rd, wr, 1);
            const uint32_t far = (uint32_t)(cbase - 49152) & 0xFFFFu;
            for (int i = lane; i < kTabWords; i += 32) {
                uint32_t v = tab32[i];
                uint32_t lo = v & 0xFFFFu, hi = v >> 16;
                if (((uint32_t)(cbase - lo) & 0xFFFFu) >= 32768u) lo = far;
                if (((uint32_t)(cbase - hi) & 0xFFFFu) >= 32768u) hi = far;
                tab32[i] = lo | (hi << 16);
            }
            __syncwarp();
        }
        // bytes 4..7 after the lane's word: the next lane's word, or the next chunk's first
        const uint32_t up = __shfl_down_sync(0xffffffffu, a, 1), wrap = __shfl_sync(0xffffffffu, nn, 0);
        const uint32_t b = lane == 31 ? wrap : up;
        const int s0 = cbase + 4 * lane;
        // positions with a hash triple that may enter a chain: all four, except in the first and the last chunks
        unsigned ok = 0xFu;
        if (cbase == 0 || cbase + 130 >= n) {  // (warp-uniform)
            ok = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) if (s0 + j >= 1 && s0 + j + 2 < n) ok |= 1u << j;
        }
        uint32_t hh[4];
        unsigned mine = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t t = __funnelshift_r(a, b, 8 * j);  // bytes of position s0 + j, low byte first
            hh[j] = (((t & 0xFFu) << 10) ^ ((t >> 3) & 0x1FE0u) ^ ((t >> 16) & 0xFFu)) & 0x7FFFu;  // hash3
            if ((hh[j] & (kChainWarps - 1)) == (uint32_t)k) mine |= 1u << j;
        }
        mine &= ok;
        // ring slots in position order: the lanes below first, then this lane's own earlier positions
        int before = 0, total = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const unsigned bal = __ballot_sync(0xffffffffu, (mine >> j) & 1u);
            before += __popc(bal & below);
            total += __popc(bal);
        }
        int slot = wr + before;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if ((mine >> j) & 1u) {
                ring_s[slot & (kRing - 1)] = (uint32_t)(s0 + j);
                ring_h[slot & (kRing - 1)] = (uint16_t)(hh[j] >> kChainWarpsLog2);
                slot++;
            }
        }
        wr += total;
        __syncwarp();
        flush(rd, wr, 32);
    }
    flush(rd, wr, 1);
}

// Run lengths around every position: rb[p] / rl[p] = number of bytes equal to w[p] immediately before / after p,
// packed with the chain link and the first bytes into the ChainRec the walks load.  One warp owns a chunk of
// kRunChunk positions, 32 per step, one position per lane: "where did the current run start" is a ballot and a
// count-leading-zeros, the value entering the next step one shuffle -- no shared memory, coalesced loads and stores.
// Chunks are independent once the count entering each is known, so the work is three launches:
//   k_run_edges  per chunk: (does one run cover the whole chunk?, count leaving it if nothing came in), both ways
//   k_run_link   per stream: warp scan of those summaries -> the count entering every chunk from either side
//   k_run_recs   per chunk: counts of every position, records written out
// The summaries combine as (whole, val) o (whole', val') = whole' ? (whole, val + val') : (false, val').
constexpr int kRunChunk = 256;
constexpr int kRunSteps = kRunChunk / 32;
constexpr int kRunWarps = 8;
constexpr unsigned kAll = 0xffffffffu;

struct RunCarry { int whole, val; };
__device__ __forceinline__ RunCarry run_combine(RunCarry a, RunCarry b) {  // a = earlier part, b = later part
    return b.whole ? RunCarry{a.whole, a.val + b.val} : RunCarry{0, b.val};
}

// The chunk's bytes, one per lane per step (c[j] = w[base + 32 j + lane], 0 past the end), and the bytes around it.
struct RunBytes { uint32_t c[kRunSteps]; uint32_t before, after; };  // before = w[base - 1]; after: lanes 0..2 hold w[base + kRunChunk + lane]
__device__ __forceinline__ void run_bytes(RunBytes &r, const uint8_t *w, int n, int base, int lane) {
#pragma unroll
    for (int j = 0; j < kRunSteps; j++) { const int p = base + 32 * j + lane; r.c[j] = p < n ? __ldg(w + p) : 0u; }
    r.before = base > 0 ? __ldg(w + base - 1) : 0u;
    const int q = base + kRunChunk + lane;
    r.after = (lane < 3 && q < n) ? __ldg(w + q) : 0u;
}
// step j: mask of lanes whose byte equals the byte before it / after it (false outside the payload)
__device__ __forceinline__ unsigned run_eq_prev(const RunBytes &r, int j, int n, int base, int lane) {
    const uint32_t prev_group = j ? r.c[j - 1] : 0u;
    uint32_t pc = __shfl_up_sync(kAll, r.c[j], 1);
    const uint32_t last_of_prev = __shfl_sync(kAll, prev_group, 31);
    if (lane == 0) pc = j ? last_of_prev : r.before;
    const int p = base + 32 * j + lane;
    return __ballot_sync(kAll, p >= 1 && p < n && r.c[j] == pc);
}
__device__ __forceinline__ unsigned run_eq_next(const RunBytes &r, int j, int n, int base, int lane) {
    const uint32_t next_group = j + 1 < kRunSteps ? r.c[j + 1] : r.after;
    uint32_t nc = __shfl_down_sync(kAll, r.c[j], 1);
    const uint32_t first_of_next = __shfl_sync(kAll, next_group, 0);
    if (lane == 31) nc = first_of_next;
    const int p = base + 32 * j + lane;
    return __ballot_sync(kAll, p + 1 < n && r.c[j] == nc);
}

__global__ void __launch_bounds__(32 * kRunWarps) k_run_edges(const uint8_t *__restrict__ in, int64_t in_stride,
                                                             const int32_t *__restrict__ in_len, int4 *__restrict__ edges, int max_chunks) {
    const int stream = blockIdx.y, chunk = blockIdx.x * kRunWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int base = chunk * kRunChunk, n = in_len[stream];
    if (base >= n) return;
    RunBytes r;
    run_bytes(r, in + (int64_t)stream * in_stride, n, base, lane);
    int fw = 1, fv = 0, bw = 1, bv = 0;
#pragma unroll
    for (int j = 0; j < kRunSteps; j++) {
        const unsigned m = run_eq_prev(r, j, n, base, lane);
        if (m == kAll) fv += 32; else { fw = 0; fv = __clz((int)~m); }  // equal flags at the top of the group
    }
#pragma unroll
    for (int j = kRunSteps - 1; j >= 0; j--) {
        const unsigned m = run_eq_next(r, j, n, base, lane);
        if (m == kAll) bv += 32; else { bw = 0; bv = __ffs((int)~m) - 1; }  // equal flags at the bottom of the group
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

__global__ void __launch_bounds__(32 * kRunWarps) k_run_recs(const uint8_t *__restrict__ in, int64_t in_stride,
                                                            const int32_t *__restrict__ in_len, const uint16_t *__restrict__ dist16,
                                                            const int2 *__restrict__ enters, int max_chunks,
                                                            ChainRec *__restrict__ recs_all, int64_t stride) {
    const int stream = blockIdx.y, chunk = blockIdx.x * kRunWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int base = chunk * kRunChunk, n = in_len[stream];
    if (base >= n) return;
    const uint16_t *d16 = dist16 + (int64_t)stream * stride;
    ChainRec *recs = recs_all + (int64_t)stream * stride;
    const uint8_t *w = in + (int64_t)stream * in_stride;
    RunBytes r;
    run_bytes(r, w, n, base, lane);
    const int2 enter = enters[(int64_t)stream * max_chunks + chunk];
    uint32_t link[kRunSteps];  // chain links of the chunk, fetched up front with the bytes (all loads in flight at once)
#pragma unroll
    for (int j = 0; j < kRunSteps; j++) { const int p = base + 32 * j + lane; link[j] = p < n ? __ldg(d16 + p) : 0u; }
    uint32_t rb[kRunSteps];
    int carry = enter.x;  // count of the position just before the group
#pragma unroll
    for (int j = 0; j < kRunSteps; j++) {
        const unsigned m = run_eq_prev(r, j, n, base, lane);
        const unsigned zeros_below = ~m & ((2u << lane) - 1u);  // positions <= lane whose flag is 0
        const int back = zeros_below ? lane - (31 - __clz((int)zeros_below)) : lane + 1 + carry;
        carry = __shfl_sync(kAll, back, 31);
        rb[j] = (uint32_t)min(back, 65535);
    }
    carry = enter.y;  // count of the position just after the group
#pragma unroll
    for (int j = kRunSteps - 1; j >= 0; j--) {
        const unsigned m = run_eq_next(r, j, n, base, lane);
        const unsigned zeros_above = ~m & ~((1u << lane) - 1u);  // positions >= lane whose flag is 0
        const int fwd = zeros_above ? (__ffs((int)zeros_above) - 1) - lane : (32 - lane) + carry;
        carry = __shfl_sync(kAll, fwd, 0);
        // w[p + 1 .. p + 3]: the next lanes' bytes, or the first lanes of the next group
        const uint32_t next_group = j + 1 < kRunSteps ? r.c[j + 1] : r.after;
        uint32_t quad = r.c[j];
#pragma unroll
        for (int i = 1; i <= 3; i++) {
            const uint32_t a = __shfl_sync(kAll, r.c[j], (lane + i) & 31), nx = __shfl_sync(kAll, next_group, (lane + i) & 31);
            quad |= (lane + i < 32 ? a : nx) << (8 * i);
        }
        const int p = base + 32 * j + lane;
        if (p < n) {
            const int rl = min(fwd, 65535);
            ChainRec rec = pack_rec_plain((int)link[j], rl, quad);
            if (rl >= 2) {
                // a run position: the link of the bottom of its run group and the byte after the run (one address per
                // run, so the lanes of a run share the loads)
                const int bottom = max(p - (int)rb[j], 1), q = p + rl + 1;
                rec = pack_rec_run((int)__ldg(d16 + bottom), rl, (int)rb[j], r.c[j], (rl < 65535 && q < n) ? (uint32_t)__ldg(w + q) : 0u);
            }
            store_rec(recs + p, rec);
        }
    }
}

// ------------------------------------------------------------------ k_probe
// Cold-start longest_match for every position (match_probe in deflate_core.cuh), one thread per position, 32
// registers so that 64 warps per SM hide the dependent chain loads.  Chain lengths vary wildly (over half the
// positions have no chain, some need the full kProbeK steps) and only ~11 of 32 lanes are active on average; two
// work-queue variants that refill finished lanes from a shared counter (per lane, and in rounds of >= 8 idle
// lanes) were measured 12 % slower -- the extra state and the refill branch cost more than the idle lanes.  Running
// the walk in phases with growing step caps (2/4/10/48) and compacting the open walks in between halves the warp
// steps on paper, yet measured 9 % slower with the phases inside one CTA (shared-memory queues, barriers) and 20 %
// slower as separate kernels over per-stream queues: the late phases are chains of dependent L2 loads with nothing
// else to issue, while here the short walks of other warps fill those gaps.
constexpr int kProbeThreads = 128;
constexpr int kProbeUnit = 128;  // consecutive positions a warp takes at a time
__device__ __forceinline__ uint32_t sm_id() { uint32_t r; asm("mov.u32 %0, %%smid;" : "=r"(r)); return r; }

// The walks of neighbouring positions read the same 256 KB of records (the 32 K positions before them), far too much
// to share when the warps of an SM sit on tiles scattered over the batch (L1 hit rate 36 %, and the kernel is bound
// by the latency of those loads: six CTAs per SM instead of eight cost 20 %).  So the kernel is persistent and
// places work by SM: the units (kProbeUnit positions) of the whole batch are cut into one contiguous range per SM,
// and a warp takes the next unit of the range of the SM it runs on -- the 64 warps of an SM walk 8 K consecutive
// positions of one stream and share their window in L1.  A warp whose range is used up moves on to the next SM's.
// (Halving the window with a 4-byte "hot" word per position -- link, run flag and the 11 bits of the first two bytes
// the hash does not determine, the full record opened only on a key match or a run group -- measured 20 % SLOWER:
// run groups and true matches then cost two dependent loads, and lanes on the two record formats diverge.  On top of
// this placement, giving every position 2-10 steps first and finishing the open walks 32 at a time from a per-warp
// shared-memory queue -- 27 % fewer warp steps on paper -- changed nothing: 30.7 ms against 30.2.)
__global__ void __launch_bounds__(kProbeThreads, 16) k_probe(const uint8_t *__restrict__ in, uint32_t in_stride,
                                                             const int32_t *__restrict__ in_len, const ChainRec *__restrict__ recs,
                                                             uint32_t *__restrict__ mlo, uint32_t *__restrict__ mhi, uint32_t stride,
                                                             uint32_t units_per_stream, uint32_t total_units,
                                                             uint32_t *__restrict__ sm_next, uint32_t n_sm) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t per = (total_units + n_sm - 1) / n_sm;
    const uint32_t home = sm_id() % n_sm;
    for (uint32_t hop = 0; hop < n_sm; hop++) {
        const uint32_t sm = home + hop < n_sm ? home + hop : home + hop - n_sm;
        const uint32_t lo_u = sm * per, hi_u = min(total_units, lo_u + per);
        for (;;) {
            uint32_t u = 0;
            if (lane == 0) u = lo_u + atomicAdd(sm_next + sm, 1u);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (u >= hi_u) break;
            const uint32_t stream = u / units_per_stream, first = (u - stream * units_per_stream) * kProbeUnit;
            const int n = in_len[stream];
            if ((int)first >= n) continue;
            // strides are 32-bit so that re-deriving a per-stream pointer is one wide multiply-add
            const uint64_t o = (uint64_t)stream * stride;
            const uint8_t *w_s = in + (uint64_t)stream * in_stride;
            const ChainRec *recs_s = recs + o;
            // opaque to the compiler: at 32 registers it would otherwise re-derive both bases from the kernel
            // parameters (nine instructions) at every use inside the walk instead of keeping them
            asm volatile("" : "+l"(w_s), "+l"(recs_s));
            LzView v{w_s, recs_s, n};
#pragma unroll 1
            for (int s = (int)(first + lane); s < (int)first + kProbeUnit && s < n; s += 32) {
                uint32_t lo, hi;
                match_probe(v, s, kProbeK, &lo, &hi);
                mlo[o + s] = lo;
                mhi[o + s] = hi;
            }
        }
    }
}

// ------------------------------------------------------------------ parse: k_next / k_walk_spec / k_walk_stitch / k_symscan
// deflate_slow's lazy state machine, re-indexed by symbol start (deflate_core.cuh):
//   k_next         one thread per position: the symbol a walk standing there emits, from the probe table alone
//                  (coalesced; kSymSlow where the step needs match_full or depends on the walk's history)
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
        const uint32_t e[5] = {e4.x, e4.y, e4.z, e4.w, p0 + 4 < n ? lo[p0 + 4] : 0u};
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
                if (e0 & kProbeIncomplete) r = kSymSlow;
                else if (e0 != 0 && !(l0 == kMinMatch && d0 > kTooFar)) {
                    r = sym_pack(l0 - kMinMatch, d0);
                    const int look1 = look - 1;
                    if (look1 >= kMinMatch) {
                        const uint32_t e1 = l0 >= kGoodLen ? hi[p + 1] : e[i + 1];
                        if (!(e1 == 0 || l0 >= kMaxMatch || l0 >= look1)) {
                            if (e1 & kProbeIncomplete) r = kSymSlow;
                            else if (mtab_len(e1) > l0) r = (l0 >= kGoodLen && e[i + 1] != e1) ? kSymSlow : lit;
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
                                                   const int32_t *__restrict__ in_len, const ChainRec *__restrict__ recs,
                                                   const uint32_t *__restrict__ mlo, const uint32_t *__restrict__ mhi, uint32_t *__restrict__ mask,
                                                   int64_t mask_stride, uint32_t *__restrict__ rec, Walk *__restrict__ exits, int64_t stride,
                                                   int max_win) {
    const int stream = blockIdx.y;
    const int k = blockIdx.x * 128 + threadIdx.x;
    const int n = in_len[stream];
    const int ws = k * kWalkWin;
    if (ws >= n) return;
    const int64_t o = (int64_t)stream * stride;
    LzView v{in + (int64_t)stream * in_stride, recs + o, n};
    exits[(int64_t)stream * max_win + k] = walk_window(v, ProbeTab{mlo + o, mhi + o}, ws, min(ws + kWalkWin, n),
                                                       mask + (int64_t)stream * mask_stride, rec + o);
}

__global__ void __launch_bounds__(128) k_walk_stitch(const uint8_t *__restrict__ in, int64_t in_stride,
                                                    const int32_t *__restrict__ in_len, const ChainRec *__restrict__ recs,
                                                    const uint32_t *__restrict__ mlo, const uint32_t *__restrict__ mhi, uint32_t *__restrict__ mask,
                                                    int64_t mask_stride, uint32_t *__restrict__ rec, const Walk *__restrict__ exits, int64_t stride,
                                                    int max_win, int B) {
    // one warp per stream, lane 0 walks: streams diverge completely, so sharing a warp would serialise them
    const int stream = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (stream >= B || (threadIdx.x & 31)) return;
    const int64_t o = (int64_t)stream * stride;
    LzView v{in + (int64_t)stream * in_stride, recs + o, in_len[stream]};
    walk_stitch(v, ProbeTab{mlo + o, mhi + o}, kWalkWin, mask + (int64_t)stream * mask_stride, rec + o, exits + (int64_t)stream * max_win);
}

constexpr int kScanThreads = 256;
__global__ void __launch_bounds__(kScanThreads) k_symscan(const int32_t *__restrict__ in_len, const uint32_t *__restrict__ mask,
                                                          int64_t mask_stride, const uint32_t *__restrict__ rec, uint32_t *__restrict__ syms,
                                                          int64_t stride, int32_t *__restrict__ blk_end, int max_blocks,
                                                          int32_t *__restrict__ nsym, int32_t *__restrict__ nflush) {
    const int stream = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = in_len[stream];
    const int64_t o = (int64_t)stream * stride;
    const uint32_t *mk = mask + (int64_t)stream * mask_stride;
    const uint32_t *rc = rec + o;
    uint32_t *out = syms + o;
    int32_t *be = blk_end + (int64_t)stream * max_blocks;
    __shared__ int warp_sum[kScanThreads / 32];
    __shared__ int carry_s;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    const int n_words = (n + 31) >> 5;
    for (int base = 0; base < n_words; base += kScanThreads) {  // one mask word (32 positions) per thread per tile
        const int wi = base + tid;
        uint32_t bits = wi < n_words ? mk[wi] : 0u;
        const int cnt = __popc(bits);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) warp_sum[wid] = incl;
        __syncthreads();
        int woff = 0, total = 0;
        for (int w = 0; w < kScanThreads / 32; w++) { if (w < wid) woff += warp_sum[w]; total += warp_sum[w]; }
        int idx = carry_s + woff + incl - cnt;
        __syncthreads();
        if (tid == 0) carry_s += total;
        // compaction, a word at a time by the whole warp: lane l looks at position 32 j + l of word j, so the symbol
        // loads and the dense stores are coalesced
        const int wbase = base + (wid << 5);
#pragma unroll 4
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
        __syncthreads();
    }
    if (tid == 0) {
        const int total = carry_s;
        nsym[stream] = total;
        nflush[stream] = finish_blocks(total, total ? out[total - 1] : 0u, n, be);
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
    for (int i = tid; i < n; i += 256) {
        unsigned long long d = w[i];
        sa += d; sb += d * (unsigned long long)(n - i);
    }
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
// queue and the block descriptor is assembled where the literal tree's Freq/Dad arrays were (both dead by then).
struct BlockBuildSmem {
    BlockWork wk;
};
static_assert(sizeof(BlockDesc) <= 2 * kHeapSize * sizeof(uint16_t), "descriptor must fit in lt.freq + lt.dad");
static_assert((kLCodes + 2 + kDCodes + 2) * sizeof(uint32_t) <= sizeof(TreeScratch::pq), "histogram must fit in the queue");

__global__ void __launch_bounds__(32) k_block_build(const uint32_t *__restrict__ syms, int64_t stride,
                                                    const int32_t *__restrict__ blk_end, int max_blocks,
                                                    const int32_t *__restrict__ nsym, const int32_t *__restrict__ nflush,
                                                    const int32_t *__restrict__ in_len, BlockDesc *__restrict__ desc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BlockBuildSmem &sm = *reinterpret_cast<BlockBuildSmem *>(smem_raw);
    const int stream = blockIdx.y, blk = blockIdx.x, lane = threadIdx.x;
    const int nf = nflush[stream];
    if (blk > nf) return;
    uint32_t *z = reinterpret_cast<uint32_t *>(&sm);
    for (int i = lane; i < (int)(sizeof(BlockBuildSmem) / 4); i += 32) z[i] = 0;
    __syncwarp();
    uint32_t *lh = sm.wk.sc.pq, *dh = sm.wk.sc.pq + kLCodes + 2;
    const int32_t *be = blk_end + (int64_t)stream * max_blocks;
    const uint32_t sym_begin = (uint32_t)blk * kSymEnd;
    const uint32_t ns = blk < nf ? (uint32_t)kSymEnd : (uint32_t)(nsym[stream] - nf * kSymEnd);
    const uint32_t pos_begin = blk == 0 ? 0u : (uint32_t)be[blk - 1];
    const uint32_t pos_end = blk < nf ? (uint32_t)be[blk] : (uint32_t)in_len[stream];
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
    for (int i = lane; i < kLCodes; i += 32) sm.wk.lt.freq[i] = (uint16_t)lh[i];
    if (lane < kDCodes) sm.wk.dt.freq[lane] = (uint16_t)dh[lane];
    __syncwarp();
    BlockDesc *d = reinterpret_cast<BlockDesc *>(sm.wk.lt.freq);
    if (lane == 0) {
        sm.wk.lt.freq[256] = 1;
        build_block(sm.wk, d, pos_end - pos_begin, blk == nf);
        d->sym_begin = sym_begin; d->nsym = ns; d->pos_begin = pos_begin;
    }
    __syncwarp();
    fill_block_tables(sm.wk, d, lane, 32);
    __syncwarp();
    uint32_t *dst = reinterpret_cast<uint32_t *>(desc + (int64_t)stream * max_blocks + blk);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(d);
    for (int i = lane; i < (int)(sizeof(BlockDesc) / 4); i += 32) dst[i] = src[i];
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
    for (int i = tid; i < kLCodes + 2; i += kPackThreads) { lcode[i] = d.lcode[i]; llen[i] = d.llen[i]; }
    if (tid < kDCodes + 2) { dcode[tid] = d.dcode[tid]; dlen[tid] = d.dlen[tid]; }
    __syncthreads();
    const int ns = (int)d.nsym;
    // A thread's symbols are consecutive (its bits are, too), so its loads stride the warp.  The runs are cut on
    // 16-byte boundaries of the symbol array -- `mis` symbols before the block's first one is such a boundary -- so a
    // thread fetches its run in 16-byte loads, two at a time (a quarter of the requests of 4-byte loads; the first and
    // last quad of the block may reach outside it: those symbols are skipped, the array is padded).
    const uint32_t *sy = syms + (int64_t)stream * stride + d.sym_begin;
    const int mis = (int)(d.sym_begin & 3u);
    const uint4 *quads = reinterpret_cast<const uint4 *>(sy - mis);
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
    for_each([&](uint32_t sym) { bits += (uint32_t)sym_nbits(sym, llen, dlen); });
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
            int nb;
            unsigned long long b = sym_bits(sym, lcode, llen, dcode, dlen, &nb);
            if (nb > 32) { sink.put((uint32_t)b, 32); sink.put((uint32_t)(b >> 32), nb - 32); }
            else sink.put((uint32_t)b, nb);
        });
        sink.finish();
    }
    if (tid == 0)  // END_BLOCK closes the block
        or_bits(out32, bit_start + (uint64_t)hdr_bits + d.body_bits - llen[256], lcode[256], llen[256]);
}

// Per-device launch state: opt-in shared-memory sizes are a property of (function, device) and the SM count of the
// device, so both are kept per device (several devices may be driven from one process through cct_*_host).
struct DevState { int n_sm = 0; };
static const DevState &dev_state() {
    static std::mutex mu;
    static DevState per_dev[64];
    int dev = 0;
    cudaGetDevice(&dev);
    dev = std::min(std::max(dev, 0), 63);
    std::lock_guard<std::mutex> lk(mu);
    DevState &d = per_dev[dev];
    if (!d.n_sm) {
        const int chain_smem = kHeadEntries * (int)sizeof(uint16_t) + (int)sizeof(ChainSmem);
        cudaFuncSetAttribute(k_chains, cudaFuncAttributeMaxDynamicSharedMemorySize, chain_smem);
        cudaFuncSetAttribute(k_block_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BlockBuildSmem));
        cudaDeviceGetAttribute(&d.n_sm, cudaDevAttrMultiProcessorCount, dev);
    }
    return d;
}

void launch_chains(const uint8_t *in, int64_t in_stride, const int32_t *in_len, int B, uint16_t *dist16, int64_t stride,
                   cudaStream_t st) {
    if (B == 0) return;
    const int chain_smem = kHeadEntries * (int)sizeof(uint16_t) + (int)sizeof(ChainSmem);
    dev_state();
    { CCT_KERNEL(k_chains, st);
    k_chains<<<B, 32 * kChainWarps, chain_smem, st>>>(in, in_stride, in_len, dist16, stride);
    }
}

void launch_deflate(const uint8_t *in, int64_t in_stride, const int32_t *in_len, int B, int64_t max_len, uint8_t *out,
                    int64_t out_stride, int32_t *out_len, const DeflateWs &ws, cudaStream_t st) {
    if (B == 0) return;
    const int chain_smem = kHeadEntries * (int)sizeof(uint16_t) + (int)sizeof(ChainSmem);
    const DevState &dev = dev_state();
    { CCT_KERNEL(k_chains, st);
    k_chains<<<B, 32 * kChainWarps, chain_smem, st>>>(in, in_stride, in_len, ws.dist16, ws.stride);
    }
    {   // the chunk summaries live in arrays the parse fills later
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
        { CCT_KERNEL(k_run_recs, st);
        k_run_recs<<<gt, 32 * kRunWarps, 0, st>>>(in, in_stride, in_len, ws.dist16, enters, max_chunks, ws.recs, ws.stride);
        }
    }
    {
        // (asking for an all-L1 carve-out made the kernel 38 % slower: fewer CTAs fit once no shared memory is configured)
        const int n_sm = dev.n_sm;
        const uint32_t units_per_stream = (uint32_t)((max_len + kProbeUnit - 1) / kProbeUnit);
        { CCT_KERNEL(m_zero_probe, st); cudaMemsetAsync(ws.probe_next, 0, kMaxSm * sizeof(uint32_t), st); }
        CCT_KERNEL(k_probe, st);
        k_probe<<<n_sm * 16, kProbeThreads, 0, st>>>(in, (uint32_t)in_stride, in_len, ws.recs, ws.mtab, ws.mhi, (uint32_t)ws.stride,
                                                     units_per_stream, units_per_stream * (uint32_t)B, ws.probe_next,
                                                     (uint32_t)std::min(n_sm, kMaxSm));
    }
    dim3 gn((unsigned)std::min<int64_t>((max_len + 1023) / 1024, 512), (unsigned)B);
    { CCT_KERNEL(k_next, st);
    k_next<<<gn, 256, 0, st>>>(in, in_stride, in_len, ws.mtab, ws.mhi, ws.rec, ws.stride);
    }
    dim3 gw((unsigned)((ws.max_win + 127) / 128), (unsigned)B);
    { CCT_KERNEL(k_walk_spec, st);
    k_walk_spec<<<gw, 128, 0, st>>>(in, in_stride, in_len, ws.recs, ws.mtab, ws.mhi, ws.mask, ws.mask_stride, ws.rec, ws.exits,
                                    ws.stride, ws.max_win);
    }
    { CCT_KERNEL(k_walk_stitch, st);
    k_walk_stitch<<<(B + 3) / 4, 128, 0, st>>>(in, in_stride, in_len, ws.recs, ws.mtab, ws.mhi, ws.mask, ws.mask_stride, ws.rec,
                                                ws.exits, ws.stride, ws.max_win, B);
    }
    { CCT_KERNEL(k_symscan, st);
    k_symscan<<<B, kScanThreads, 0, st>>>(in_len, ws.mask, ws.mask_stride, ws.rec, ws.syms, ws.stride, ws.blk_end, ws.max_blocks,
                                          ws.nsym, ws.nflush);
    }
    { CCT_KERNEL(k_adler, st);
    k_adler<<<B, 256, 0, st>>>(in, in_stride, in_len, ws.adler);
    }
    dim3 gb((unsigned)ws.max_blocks, (unsigned)B);
    { CCT_KERNEL(k_block_build, st);
    k_block_build<<<gb, 32, sizeof(BlockBuildSmem), st>>>(ws.syms, ws.stride, ws.blk_end, ws.max_blocks, ws.nsym, ws.nflush,
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
