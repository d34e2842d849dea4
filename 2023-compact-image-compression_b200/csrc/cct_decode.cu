// cct_decode.cu — decoder: header check, inflate, op stream -> values, JUMP units, scatter to raster.
//
// Reference path restated (reference src/codec/core.py):
//   core.py:385-402   read_header (magic, dims, flags)              -> k_dec_header
//   core.py:420-421   zlib.decompress(file_bytes[13:])              -> k_inflate_sym (thread per stream) + k_lz_expand
//                                                                      (warp per stream) + k_adler_check
//   core.py:453-520   padded-order walk with JUMP splicing          -> k_dec_ops_scan/link/emit (chunk-parallel ops ->
//                                                                      running values), k_dec_units (JUMP unit table)
//                                                                      k_dec_scatter (value k -> raster index)
#include <cuda_runtime.h>

#include "cct_kernels.h"
#include "inflate_core.cuh"

namespace cct {

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

size_t decode_ws_bytes(int B, int64_t n_px) {
    DecodeWs ws;
    return decode_ws_carve(&ws, nullptr, B, n_px);
}

size_t decode_ws_carve(DecodeWs *ws, void *mem, int B, int64_t n_px) {
    char *p = (char *)mem;
    size_t used = 0;
    auto take = [&](size_t bytes) { void *r = p ? p + used : nullptr; used += (size_t)align_up((int64_t)bytes, 256); return r; };
    int64_t nb = n_px / kBlock;
    ws->pstride = payload_bound(n_px) + 16;
    ws->payload = (uint8_t *)take((size_t)B * ws->pstride);
    ws->plen = (int32_t *)take((size_t)B * sizeof(int32_t));
    ws->vals = (uint16_t *)take((size_t)B * n_px * sizeof(uint16_t));
    ws->jflag = (uint8_t *)take((size_t)B * nb);
    ws->slotA = (int32_t *)take((size_t)B * nb * sizeof(int32_t));
    ws->slotB = (int32_t *)take((size_t)B * nb * sizeof(int32_t));
    ws->fl = (uint8_t *)take((size_t)B);
    used += inflate_ws_carve(&ws->iw, p ? p + used : nullptr, B, ws->pstride);
    ws->max_chunks = (int)(ws->pstride / kOpsChunk + 1);
    ws->chunks = (OpsChunkSum *)take((size_t)B * ws->max_chunks * sizeof(OpsChunkSum));
    return used;
}

// ------------------------------------------------------------------ inflate (two passes)
// k_inflate_sym   Huffman decode, one thread per stream with its code tables in shared memory: the bit-serial
//                 part of inflate touches only the bit reader and the tables and appends LZ77 symbols to a
//                 per-stream list;
// k_lz_expand     LZ77 expansion, one warp per stream: 32 symbols at a time, output offsets by a warp scan,
//                 literals stored in parallel, matches copied by the whole warp in symbol order;
// k_adler_check   Adler-32 of the expanded bytes against the stream trailer (CTA per stream).
struct StreamSrc {  // where stream i lives: arena + offsets (with a header to skip) or fixed stride + lengths
    const uint8_t *base; const int64_t *off; int64_t stride; const int32_t *len; int skip;
    __device__ const uint8_t *ptr(int i) const { return base + (off ? off[i] : (int64_t)i * stride) + skip; }
    __device__ int size(int i) const { return (off ? (int)(off[i + 1] - off[i]) : len[i]) - skip; }
};

// One stream per CTA of kSubLanes threads, all lanes decoding (inflate_core.cuh, "Lane-parallel Huffman decode");
// the stream's tables, the staged bits of the round and the few words the lanes exchange live in shared memory.
#ifndef CCT_INF_CTAS
#define CCT_INF_CTAS 10  // resident CTAs per SM the register budget is set for (shared memory allows ten)
#endif
constexpr int kInfWarps = kSubLanes / 32;
constexpr unsigned kFull = 0xffffffffu;

struct StreamBits {  // the stream as aligned 32-bit words; positions are bits from the aligned origin
    const uint8_t *in; int in_len;       // zlib stream (header, deflate data, Adler-32)
    const uint32_t *words; uint32_t skew;  // words = in rounded down to 4 bytes, skew = bits between them
    __device__ __forceinline__ uint32_t word(uint32_t w) const {
        const uint8_t *b = reinterpret_cast<const uint8_t *>(words + w);
        if (b >= in && b + 4 <= in + in_len) return __ldg(words + w);
        uint32_t v = 0;  // the first and last few bytes: only what belongs to the stream, zeros elsewhere
        for (int k = 0; k < 4; k++) if (b + k >= in && b + k < in + in_len) v |= (uint32_t)b[k] << (8 * k);
        return v;
    }
};

struct InflateShared {
    FastTables tb;
    uint32_t ends[kSubLanes];                      // where each lane's last symbol ended
#ifndef CCT_INF_MARKS
#define CCT_INF_MARKS 0  // sync marks (inflate_core.cuh) save a tenth of the decoding but cost two resident CTAs per SM: 10.6 ms against 8.4
#endif
#if CCT_INF_MARKS
    uint32_t marks[kMarkWords][kSubLanes];         // sync marks of each lane's list (inflate_core.cuh)
#endif
    unsigned ok_mask[kInfWarps], stop_mask[kInfWarps];
    int warp_n[kInfWarps], warp_b[kInfWarps];      // symbols / bytes per warp (scan)
    int end_flag; uint32_t end_pos;                // of the lane that closes the round
    int type, nlen, ndist, last; uint32_t bp;      // block header, from thread 0
    int ok;
};

// Construction of one code's tables from tb.lens[first .. first + nsym) by the whole CTA.
__device__ __forceinline__ bool cta_build_code(InflateShared &sh, int first, int nsym, bool lit) {
    FastTables &tb = sh.tb;
    const int tid = threadIdx.x, lane = tid & 31;
    FastCanon &c = lit ? tb.lc : tb.dc;
    uint32_t *syms = lit ? tb.lsym : tb.dsym;
    if (tid < 16) c.count[tid] = 0;
    __syncthreads();
    fast_count(tb.lens + first, nsym, c, tid, kSubLanes);
    __syncthreads();
    if (tid == 0) sh.ok = fast_ranges(c);
    __syncthreads();
    if (!sh.ok) return false;
    if (tid < 32) {
        // symbols ordered by (length, symbol), 32 at a time by the first warp (fast_sort is the sequential
        // statement): a symbol's slot is its length's running cursor plus its rank among the lanes of equal length
        uint16_t *next = reinterpret_cast<uint16_t *>(tb.stage);
        if (lane < 16) next[lane] = c.offs[lane];
        __syncwarp();
        for (int base = 0; base < nsym; base += 32) {
            const int i = base + lane;
            const int l = i < nsym ? tb.lens[first + i] : 0;
            const unsigned peers = __match_any_sync(kFull, l);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            const int start = next[l];
            __syncwarp();
            if (l) {
                syms[start + rank] = lit ? lit_entry(i, 0) : dst_entry(i, 0);
                if (rank == 0) next[l] = (uint16_t)(start + __popc(peers));
            }
            __syncwarp();
        }
    }
    __syncthreads();
    fast_fill(c, syms, tb.tab + (lit ? 0 : kFDstAt), lit ? kFLitBits : kFDstBits, tid, kSubLanes);
    __syncthreads();
    return true;
}

// The symbols of one Huffman block, starting at bit `q` (a true symbol boundary).  Returns false on an error.
// Every thread of the CTA calls it; q, ns, op are kept identical in all threads.
__device__ __forceinline__ bool cta_block_symbols(const StreamBits &sb, uint32_t data_end, InflateShared &sh, uint32_t &q,
                                                  uint32_t *syms, int &ns, int &op, int sym_cap, int out_cap) {
    FastTables &tb = sh.tb;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (;;) {
        const uint32_t w0 = q >> 5, origin = w0 << 5;
        for (int j = tid; j < kStageWords; j += kSubLanes) tb.stage[stage_slot(j)] = sb.word(w0 + j);
        __syncthreads();
        const uint32_t limit = min(q + (uint32_t)(tid + 1) * kSubBits, data_end);
        uint32_t start = q + (uint32_t)tid * kSubBits, cur = 0xffffffffu;
        Span r{0u, 0, 0, kSpanRun};
        int confirmed = 0, last = -1;
        [[maybe_unused]] bool again = false;
        for (;;) {  // rounds: lanes at and after the first unconfirmed one decode from where their predecessor ended
            if (tid >= confirmed && start != cur) {
#if CCT_INF_MARKS
                decode_span_marks<kSubLanes>(tb.stage, origin, q + (uint32_t)tid * kSubBits, start, limit, data_end, tb,
                                             &sh.marks[0][tid], again, r);
#else
                decode_span<false>(tb.stage, origin, start, limit, data_end, tb, nullptr, 0, r);
#endif
                cur = start;
            }
            again = true;
            sh.ends[tid] = r.end;
            __syncthreads();
            const uint32_t prev_end = tid ? sh.ends[tid - 1] : 0u;
            const unsigned okm = __ballot_sync(kFull, tid == 0 || cur == prev_end);
            const unsigned stm = __ballot_sync(kFull, r.flag != kSpanRun);
            if (lane == 0) { sh.ok_mask[warp] = okm; sh.stop_mask[warp] = stm; }
            __syncthreads();
            int c = 0;  // lanes [0, c) decoded from true boundaries
            for (int w = 0; w < kInfWarps; w++) {
                const unsigned m = sh.ok_mask[w];
                if (m == kFull) c += 32; else { c += __ffs((int)~m) - 1; break; }
            }
            for (int w = 0; w < kInfWarps && last < 0 && 32 * w < c; w++) {  // the block (or the data) ends in a confirmed lane?
                const int k = c - 32 * w;
                const unsigned m = sh.stop_mask[w] & (k >= 32 ? kFull : ((1u << k) - 1u));
                if (m) last = 32 * w + __ffs((int)m) - 1;
            }
            if (last >= 0) break;
            if (c == kSubLanes) { last = kSubLanes - 1; break; }
            confirmed = c;
            if (tid >= c) start = prev_end;
        }
        if (tid == last) { sh.end_flag = r.flag; sh.end_pos = r.end; }
        const int n = tid <= last ? r.n : 0, bytes = tid <= last ? r.n + r.extra : 0;
        int in_n = n, in_b = bytes;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int tn = __shfl_up_sync(kFull, in_n, d), tb_ = __shfl_up_sync(kFull, in_b, d);
            if (lane >= d) { in_n += tn; in_b += tb_; }
        }
        if (lane == 31) { sh.warp_n[warp] = in_n; sh.warp_b[warp] = in_b; }
        __syncthreads();
        const int end_flag = sh.end_flag;
        if (end_flag == kSpanBad) return false;
        int off_n = 0, off_b = 0, tot_n = 0, tot_b = 0;
        for (int w = 0; w < kInfWarps; w++) {
            if (w < warp) { off_n += sh.warp_n[w]; off_b += sh.warp_b[w]; }
            tot_n += sh.warp_n[w]; tot_b += sh.warp_b[w];
        }
        if (ns + tot_n > sym_cap || op + tot_b > out_cap) return false;
        Span w{0u, 0, 0, kSpanRun};
        if (tid <= last) decode_span<true>(tb.stage, origin, cur, limit, data_end, tb, syms + ns + off_n + in_n - n, op + off_b + in_b - bytes, w);
        ns += tot_n; op += tot_b;
        q = sh.end_pos;
        if (__syncthreads_or(w.flag == kSpanFar)) return false;  // (also: the staging area is rewritten next)
        if (end_flag == kSpanEob) return q <= data_end;
    }
}

__global__ void __launch_bounds__(kSubLanes, CCT_INF_CTAS) k_inflate_sym(StreamSrc src, const uint8_t *__restrict__ fl, int B,
                                                           uint32_t *__restrict__ syms_all, int64_t sym_stride, int out_cap,
                                                           int32_t *__restrict__ nsym, int32_t *__restrict__ out_len,
                                                           uint32_t *__restrict__ adler_want, uint32_t *__restrict__ status) {
    extern __shared__ __align__(16) unsigned char inf_raw[];
    InflateShared &sh = *reinterpret_cast<InflateShared *>(inf_raw);
    const int i = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) { nsym[i] = 0; out_len[i] = 0; adler_want[i] = 1; }
    if (status[i] || (fl && !(fl[i] & 4u))) return;
    const uint8_t *in = src.ptr(i);
    const int in_len = src.size(i);
    uint32_t *syms = syms_all + (int64_t)i * sym_stride;
    const int sym_cap = (int)sym_stride;
    bool good = in_len >= 6 && (in[0] & 0x0F) == 8 && (((uint32_t)in[0] << 8) | in[1]) % 31 == 0 && !(in[1] & 0x20);
    const int dlen = in_len - 4;  // deflate data ends where the Adler-32 trailer starts
    StreamBits sb{in, in_len, reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(in) & ~(uintptr_t)3),
                  8u * (uint32_t)(reinterpret_cast<uintptr_t>(in) & 3u)};
    const uint32_t data_end = sb.skew + 8u * (uint32_t)(good ? dlen : 0);
    uint32_t bp = 16;  // bit position in the stream (after the zlib header)
    int ns = 0, op = 0;
    bool last = false;
    while (good && !last) {
        if (tid == 0) {
            BitReader br{in, dlen, 0, 0, 0, false};
            br.seek(bp);
            bool l = false;
            int nlen = 0, ndist = 0;
            int type = inflate_block_header(br, sh.tb, &l, &nlen, &ndist, false);
            if (br.bad) type = -1;
            sh.type = type; sh.nlen = nlen; sh.ndist = ndist; sh.last = l; sh.bp = br.tell();
        }
        __syncthreads();
        if (sh.type == 2) {  // the code lengths: every thread looks up what starts at its bits, thread 0 follows the chain
            uint32_t *stage = sh.tb.stage;
            const uint32_t hq = sh.bp + sb.skew, w0 = hq >> 5, sk = hq & 31u;
            for (int j = tid; j < kClWords; j += kSubLanes) stage[kClWordAt + j] = sb.word(w0 + j);
            __syncthreads();
            const uint16_t *prim = reinterpret_cast<const uint16_t *>(stage);
            uint16_t *ent = reinterpret_cast<uint16_t *>(stage + kClEntAt);
            for (int b = tid; b < kClBits; b += kSubLanes) {
                const uint32_t rel = sk + (uint32_t)b;
                ent[b] = cl_entry(prim, __funnelshift_r(stage[kClWordAt + (rel >> 5)], stage[kClWordAt + (rel >> 5) + 1], rel & 31u));
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t used = 0;
                if (!cl_walk(ent, sh.nlen + sh.ndist, sh.tb.lens, &used) || sh.bp + used > 8u * (uint32_t)dlen) sh.type = -1;
                sh.bp += used;
            }
            __syncthreads();
        }
        const int type = sh.type, nlen = sh.nlen, ndist = sh.ndist;
        last = sh.last != 0; bp = sh.bp;
        if (type < 0) { good = false; break; }
        if (type == 0) {  // stored: LEN bytes at a byte boundary become literal symbols
            const int at = (int)(bp >> 3);
            if (at + nlen > dlen || ns + nlen > sym_cap || op + nlen > out_cap) { good = false; break; }
            for (int k = tid; k < nlen; k += kSubLanes) syms[ns + k] = in[at + k];
            ns += nlen; op += nlen; bp += 8u * (uint32_t)nlen;
            __syncthreads();  // sh.type is rewritten by the next header
            continue;
        }
        if (!cta_build_code(sh, 0, nlen, true) || !cta_build_code(sh, nlen, ndist, false)) { good = false; break; }
        uint32_t q = bp + sb.skew;
        good = cta_block_symbols(sb, data_end, sh, q, syms, ns, op, sym_cap, out_cap);
        bp = q - sb.skew;
    }
    if (good && (int)((bp + 7) >> 3) != dlen) good = false;  // the trailer follows the last block at the next byte boundary
    if (tid) return;
    if (!good) { status[i] |= 16u; return; }  // CCT_ST_INFLATE
    const uint8_t *t = in + dlen;
    nsym[i] = ns; out_len[i] = op;
    adler_want[i] = ((uint32_t)t[0] << 24) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 8) | t[3];
}

constexpr int kLzWarps = 4;
constexpr int kLzOwn = 16;              // longest match a single lane copies on its own
constexpr int kLzPer = 4;               // symbols per lane per step
constexpr int kLzStep = 32 * kLzPer;
constexpr int kLzSpan = 2048;           // most bytes a step may produce through the ring (more: one group at a time, through global memory)
constexpr int kLzRing = 4096;           // bytes of recent output a warp keeps in shared memory
constexpr int kLzBack = 512;            // bytes put back into the ring after a step that went around it
static_assert(kLzRing >= kLzSpan + kLzBack && kLzBack >= 258 + 16, "a source that is not yet in global memory is in the ring");

// 32 symbols (one per lane) expanded straight into global memory at ob + base: literals in parallel, short matches whose
// source lies before the group by their own lane, the rest in symbol order by the whole warp.
__device__ __forceinline__ void lz_group_global(uint8_t *ob, uint32_t s, bool valid, int &base, int lane) {
    const int dist = (int)(s >> 8);
    const int len = valid ? (dist ? (int)(s & 255u) + 3 : 1) : 0;
    int incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    const int pos = base + incl - len;
    if (len && !dist) ob[pos] = (uint8_t)s;
    const bool own = dist != 0 && len <= kLzOwn && dist >= len && pos - dist + len <= base;
    if (own) {  // the first bytes are loaded before any is stored: the compiler cannot tell that the two never overlap
        const uint8_t *from = ob + pos - dist;
        const uint8_t b0 = from[0], b1 = from[1], b2 = from[2];
        uint8_t b3 = 0;
        if (len > 3) b3 = from[3];
        ob[pos] = b0; ob[pos + 1] = b1; ob[pos + 2] = b2;
        if (len > 3) ob[pos + 3] = b3;
        for (int j = 4; j < len; j++) ob[pos + j] = from[j];
    }
    unsigned mm = __ballot_sync(0xffffffffu, dist != 0 && !own);
    __syncwarp();
    while (mm) {
        const int l = __ffs((int)mm) - 1;
        mm &= mm - 1;
        const int mpos = __shfl_sync(0xffffffffu, pos, l), mlen = __shfl_sync(0xffffffffu, len, l);
        const int mdist = __shfl_sync(0xffffffffu, dist, l);
        const uint8_t *from = ob + mpos - mdist;
        if (mdist >= mlen) {
            for (int j = lane; j < mlen; j += 32) ob[mpos + j] = from[j];
        } else {  // overlapping copy: the source repeats with period mdist
            for (int j = lane; j < mlen; j += 32) ob[mpos + j] = from[j % mdist];
        }
        __syncwarp();
    }
    base += __shfl_sync(0xffffffffu, incl, 31);
}

// One warp per stream, kLzStep symbols per step (kLzPer consecutive ones per lane).  A step's bytes are assembled in a
// ring in shared memory and leave for global memory as whole 16-byte units; `fl` is how far that has got.  A match
// whose source ends at or before fl reads global memory -- all those of a step at once, one memory latency for the lot;
// the others (the source reaches into the step, or the last few bytes before it) are copied ring to ring in symbol
// order.  Positions are offsets from `ob`, the stream's first byte rounded down to 16 bytes.
__global__ void __launch_bounds__(32 * kLzWarps) k_lz_expand(const uint32_t *__restrict__ syms, int64_t sym_stride,
                                                             const int32_t *__restrict__ nsym, uint8_t *__restrict__ out,
                                                             int64_t out_stride, int B) {
    __shared__ __align__(16) uint8_t ring_all[kLzWarps][kLzRing];
    const int i = blockIdx.x * kLzWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= B) return;
    constexpr unsigned kAll = 0xffffffffu;
    constexpr int M = kLzRing - 1;
    uint8_t *ring = ring_all[threadIdx.x >> 5];
    const int ns = nsym[i];
    const uint32_t *sy = syms + (int64_t)i * sym_stride;  // 16-byte aligned (sym_stride is a multiple of 4 words)
    uint8_t *o = out + (int64_t)i * out_stride;
    const int a0 = (int)(reinterpret_cast<uintptr_t>(o) & 15u);
    uint8_t *ob = o - a0;
    asm volatile("" : "+l"(sy), "+l"(ob));  // (keeps the two bases in registers instead of re-deriving them from the parameters)
    int base = a0, fl = a0;  // [a0, base) is produced, [a0, fl) is in global memory
    uint4 nx = 4 * lane < ns ? __ldg(reinterpret_cast<const uint4 *>(sy + 4 * lane)) : make_uint4(0u, 0u, 0u, 0u);
    for (int g = 0; g < ns; g += kLzStep) {
        const int k0 = g + kLzPer * lane;
        const uint4 q = nx;
        nx = k0 + kLzStep < ns ? __ldg(reinterpret_cast<const uint4 *>(sy + k0 + kLzStep)) : make_uint4(0u, 0u, 0u, 0u);
        // (what follows the last symbol in its 16-byte word is not a symbol)
        const uint32_t s0 = q.x, s1 = k0 + 1 < ns ? q.y : 0u, s2 = k0 + 2 < ns ? q.z : 0u, s3 = k0 + 3 < ns ? q.w : 0u;
        const int d0 = (int)(s0 >> 8), d1 = (int)(s1 >> 8), d2 = (int)(s2 >> 8), d3 = (int)(s3 >> 8);
        const int l0 = k0 < ns ? (d0 ? (int)(s0 & 255u) + 3 : 1) : 0, l1 = k0 + 1 < ns ? (d1 ? (int)(s1 & 255u) + 3 : 1) : 0;
        const int l2 = k0 + 2 < ns ? (d2 ? (int)(s2 & 255u) + 3 : 1) : 0, l3 = k0 + 3 < ns ? (d3 ? (int)(s3 & 255u) + 3 : 1) : 0;
        const int tot = l0 + l1 + l2 + l3;
        int incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(kAll, incl, d);
            if (lane >= d) incl += t;
        }
        const int span = __shfl_sync(kAll, incl, 31), end = base + span;
        if (span > kLzSpan) {  // long matches all over: what the ring holds goes out, then group by group in global memory
            for (int p = fl + lane; p < base; p += 32) ob[p] = ring[p & M];
            __syncwarp();
            for (int t = 0; t < kLzPer; t++) {
                const int k = g + 32 * t + lane;
                lz_group_global(ob, k < ns ? __ldg(sy + k) : 0u, k < ns, base, lane);
            }
            fl = base;
            for (int p = max(a0, base - kLzBack) + lane; p < base; p += 32) ring[p & M] = ob[p];
            __syncwarp();
            continue;
        }
        const int p0 = base + incl - tot, p1 = p0 + l0, p2 = p1 + l1, p3 = p2 + l2;
        // literals and the short matches whose source is in global memory
        const bool f0 = d0 && l0 <= kLzOwn && p0 - d0 + l0 <= fl, f1 = d1 && l1 <= kLzOwn && p1 - d1 + l1 <= fl;
        const bool f2 = d2 && l2 <= kLzOwn && p2 - d2 + l2 <= fl, f3 = d3 && l3 <= kLzOwn && p3 - d3 + l3 <= fl;
        unsigned pm = (d0 && !f0 ? 1u : 0u) | (d1 && !f1 ? 2u : 0u) | (d2 && !f2 ? 4u : 0u) | (d3 && !f3 ? 8u : 0u);
        uint8_t b00 = 0, b01 = 0, b02 = 0, b03 = 0, b10 = 0, b11 = 0, b12 = 0, b13 = 0;
        uint8_t b20 = 0, b21 = 0, b22 = 0, b23 = 0, b30 = 0, b31 = 0, b32 = 0, b33 = 0;
        if (f0) { const uint8_t *f = ob + p0 - d0; b00 = f[0]; b01 = f[1]; b02 = f[2]; if (l0 > 3) b03 = f[3]; }
        if (f1) { const uint8_t *f = ob + p1 - d1; b10 = f[0]; b11 = f[1]; b12 = f[2]; if (l1 > 3) b13 = f[3]; }
        if (f2) { const uint8_t *f = ob + p2 - d2; b20 = f[0]; b21 = f[1]; b22 = f[2]; if (l2 > 3) b23 = f[3]; }
        if (f3) { const uint8_t *f = ob + p3 - d3; b30 = f[0]; b31 = f[1]; b32 = f[2]; if (l3 > 3) b33 = f[3]; }
        if (l0 && !d0) ring[p0 & M] = (uint8_t)s0;
        if (l1 && !d1) ring[p1 & M] = (uint8_t)s1;
        if (l2 && !d2) ring[p2 & M] = (uint8_t)s2;
        if (l3 && !d3) ring[p3 & M] = (uint8_t)s3;
#define CCT_LZ_FAR(f, p, d, l, x0, x1, x2, x3)                                                     \
        if (f) {                                                                                   \
            ring[p & M] = x0; ring[(p + 1) & M] = x1; ring[(p + 2) & M] = x2;                      \
            if (l > 3) ring[(p + 3) & M] = x3;                                                     \
            for (int t = 4; t < l; t++) ring[(p + t) & M] = ob[p - d + t];                         \
        }
        CCT_LZ_FAR(f0, p0, d0, l0, b00, b01, b02, b03)
        CCT_LZ_FAR(f1, p1, d1, l1, b10, b11, b12, b13)
        CCT_LZ_FAR(f2, p2, d2, l2, b20, b21, b22, b23)
        CCT_LZ_FAR(f3, p3, d3, l3, b30, b31, b32, b33)
#undef CCT_LZ_FAR
        __syncwarp();
        // the rest in symbol order, each by the whole warp
        unsigned bl = __ballot_sync(kAll, pm != 0);
        while (bl) {
            const int l = __ffs((int)bl) - 1;
            const unsigned pml = __shfl_sync(kAll, pm, l);
            const int j = __ffs((int)pml) - 1;
            const int pj = j == 0 ? p0 : j == 1 ? p1 : j == 2 ? p2 : p3;
            const int lj = j == 0 ? l0 : j == 1 ? l1 : j == 2 ? l2 : l3;
            const int dj = j == 0 ? d0 : j == 1 ? d1 : j == 2 ? d2 : d3;
            const int mpos = __shfl_sync(kAll, pj, l), mlen = __shfl_sync(kAll, lj, l), mdist = __shfl_sync(kAll, dj, l);
            const int src = mpos - mdist;
            if (src + mlen <= fl) {  // (a long one) from global memory
                for (int t = lane; t < mlen; t += 32) ring[(mpos + t) & M] = ob[src + t];
            } else if (mdist >= mlen) {
                for (int t = lane; t < mlen; t += 32) ring[(mpos + t) & M] = ring[(src + t) & M];
            } else {  // overlapping copy: the source repeats with period mdist
                for (int t = lane; t < mlen; t += 32) ring[(mpos + t) & M] = ring[(src + t % mdist) & M];
            }
            __syncwarp();
            if (lane == l) pm &= pm - 1;
            if ((pml & (pml - 1)) == 0) bl &= bl - 1;
        }
        // whole 16-byte units leave for global memory
        if (fl & 15) {  // (the stream does not start on a 16-byte boundary)
            const int h = min((fl + 15) & ~15, end);
            if (fl + lane < h) ob[fl + lane] = ring[(fl + lane) & M];
            fl = h;
        }
        for (int u = fl + 16 * lane; u + 16 <= end; u += 512)
            *reinterpret_cast<uint4 *>(ob + u) = *reinterpret_cast<const uint4 *>(ring + (u & M));
        if (!(fl & 15)) fl = max(fl, end & ~15);
        base = end;
        __syncwarp();
    }
    for (int p = fl + lane; p < base; p += 32) ob[p] = ring[p & M];
}

__global__ void __launch_bounds__(256) k_adler_check(const uint8_t *__restrict__ data, int64_t stride,
                                                     const int32_t *__restrict__ len, const uint32_t *__restrict__ want,
                                                     const uint8_t *__restrict__ fl, uint32_t *__restrict__ status) {
    const int i = blockIdx.x, tid = threadIdx.x;
    if (status[i] || (fl && !(fl[i] & 4u))) return;
    const int n = len[i];
    const uint8_t *w = data + (int64_t)i * stride;
    unsigned long long sa = 0, sb = 0;  // A = 1 + sum d ; B = n + sum d * (n - idx)   (mod 65521)
    adler_sums(w, n, tid, 256, sa, sb);
    __shared__ unsigned long long ra[8], rb[8];
    for (int o = 16; o; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
    if ((tid & 31) == 0) { ra[tid >> 5] = sa; rb[tid >> 5] = sb % 65521ull; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long A = 1, Bv = (unsigned long long)n;
        for (int k = 0; k < 8; k++) { A += ra[k]; Bv += rb[k]; }
        const uint32_t got = (uint32_t)(((Bv % 65521ull) << 16) | (A % 65521ull));
        if (got != want[i]) status[i] |= 16u;
    }
}

static void launch_inflate_passes(StreamSrc src, const uint8_t *fl, int B, const InflateWs &iw, uint8_t *out,
                                  int64_t out_stride, int32_t *out_len, uint32_t *status, cudaStream_t st) {
    { CCT_KERNEL(k_inflate_sym, st);
    k_inflate_sym<<<B, kSubLanes, sizeof(InflateShared), st>>>(src, fl, B, iw.syms, iw.sym_stride, (int)out_stride,
                                                             iw.nsym, out_len, iw.adler, status);
    }
    { CCT_KERNEL(k_lz_expand, st);
    k_lz_expand<<<(B + kLzWarps - 1) / kLzWarps, 32 * kLzWarps, 0, st>>>(iw.syms, iw.sym_stride, iw.nsym, out, out_stride, B);
    }
    { CCT_KERNEL(k_adler_check, st);
    k_adler_check<<<B, 256, 0, st>>>(out, out_stride, out_len, iw.adler, fl, status);
    }
}

size_t inflate_ws_carve(InflateWs *iw, void *mem, int B, int64_t out_stride) {
    char *p = (char *)mem;
    size_t used = 0;
    auto take = [&](size_t bytes) { void *r = p ? p + used : nullptr; used += (size_t)align_up((int64_t)bytes, 256); return r; };
    iw->sym_stride = align_up(out_stride + kInfSlack, 4);  // see inflate_symbols: capacity is checked per refill
    iw->syms = (uint32_t *)take((size_t)B * iw->sym_stride * sizeof(uint32_t));
    iw->nsym = (int32_t *)take((size_t)B * 4);
    iw->adler = (uint32_t *)take((size_t)B * 4);
    return used;
}

int launch_inflate(const uint8_t *in, int64_t in_stride, const int32_t *in_len, int B, uint8_t *out, int64_t out_stride,
                   int32_t *out_len, uint32_t *status, const InflateWs &iw, cudaStream_t st) {
    if (B == 0) return 0;
    StreamSrc src{in, nullptr, in_stride, in_len, 0};
    launch_inflate_passes(src, nullptr, B, iw, out, out_stride, out_len, status, st);
    return 0;
}

// ------------------------------------------------------------------ .cct decode
// status bits follow include/compact_b200.h
__global__ void k_dec_header(const uint8_t *__restrict__ files, const int64_t *__restrict__ offsets, int B, int rows,
                             int cols, uint32_t *__restrict__ status, uint8_t *__restrict__ fl) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const uint8_t *f = files + offsets[i];
    int64_t len = offsets[i + 1] - offsets[i];
    uint32_t st = 0;
    uint8_t flags = 0;
    if (len < kHeaderBytes || f[0] != 'p' || f[1] != 'a' || f[2] != 'c' || f[3] != 't') st |= 4u;  // core.py:388-389
    else {
        int w = (f[4] << 8) | f[5], h = (f[6] << 8) | f[7];
        if (w != rows || h != cols) st |= 8u;
        flags = (uint8_t)((f[10] ? 1 : 0) | (f[11] ? 2 : 0) | (f[12] ? 4 : 0));  // bool(reader.read()), core.py:397-401
    }
    status[i] = st;
    fl[i] = flags;
}

// deflate flag off: the payload follows the header as is (core.py:420)
__global__ void __launch_bounds__(128) k_dec_rawcopy(const uint8_t *__restrict__ files, const int64_t *__restrict__ offsets,
                                                     const uint8_t *__restrict__ fl, uint8_t *__restrict__ payload,
                                                     int64_t pstride, int32_t *__restrict__ plen, uint32_t *__restrict__ status) {
    const int i = blockIdx.x;
    if (status[i] || (fl[i] & 4u)) return;
    const uint8_t *src = files + offsets[i] + kHeaderBytes;
    const int len = (int)(offsets[i + 1] - offsets[i]) - kHeaderBytes;
    const int n = len < (int)pstride ? len : (int)pstride;
    uint8_t *dst = payload + (int64_t)i * pstride;
    for (int k = threadIdx.x; k < n; k += 128) dst[k] = src[k];
    if (threadIdx.x == 0) { plen[i] = n; if (n < len) status[i] |= 64u; }
}

// Op stream -> values, in three parallel passes over fixed chunks (inflate_core.cuh states them sequentially:
// ops_scan_chunk / ops_link_chunks / ops_emit_chunk).  Here a WARP owns a chunk and takes it 128 bytes at a time, a
// lane four bytes (one word).  Byte t starts an op iff the number of FULL-tagged bytes (1110xxxx) immediately before
// it is even: the byte after a byte without that tag always starts an op, and inside a run of tagged bytes starts
// alternate.  So a lane needs one bit from outside its word -- the parity of the tagged run that reaches its first
// byte -- and that is the parity of the tagged tail of the nearest lower lane that is not tagged throughout (whole
// tagged words in between add four each), found with two ballots.  Op indices and values then come from one warp
// prefix sum over the lanes' packed (delta sum, op count).  The only state crossing a step (and a chunk) is "the first
// byte is the low byte of a FULL op"; k_dec_ops_scan finds the state entering its chunk by looking BACK over the
// tagged bytes before it, so a chunk is summarised once, not once per possible entry state.
constexpr int kOpsWarps = 4;
constexpr int kOpsGroup = 128;  // bytes per warp step
constexpr int kOpsLookBack = 2048;  // bytes a chunk looks back for its entry state before it is summarised for both

__device__ __forceinline__ int op_delta(uint32_t b, uint32_t nb, bool full) {
    if (full) { const int u = (int)(((b & 0xFu) << 8) | nb); return u > 2048 ? u - 4096 : u; }  // signed(x, 12), core.py:500-506
    if (b & 0x80u) return 0;  // any other tag: the reference leaves the pixel unchanged
    const int x = (int)(b & 0x7Fu);
    return x > 64 ? x - 128 : x;  // signed(x, 7), core.py:513-516
}

// One lane's share of a step: word = payload bytes [pos, pos + 4), nb0 = the byte after them.
struct OpsLane {
    uint32_t ops;    // bit j: byte j starts a pixel op
    uint32_t jumps;  // bit j: byte j is a JUMP
    int d[4];        // delta of the op starting at byte j (0 where none)
};
// kEdge: the step may reach `stop` or the last byte before `end` (otherwise every byte is valid and every FULL op whole)
template <bool kEdge>
__device__ __forceinline__ OpsLane ops_lane(uint32_t word, uint32_t nb0, int pos, int stop, int end, uint32_t &carry, int lane) {
    uint32_t tags = 0, valid = 0xFu;
    if (kEdge) {
        valid = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) if (pos + j < stop) valid |= 1u << j;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t bj = (word >> (8 * j)) & 0xFFu;
        if ((bj & 0xF0u) == 0xE0u) tags |= 1u << j;
    }
    tags &= valid;
    // parity of the tagged run reaching byte 0 of this lane
    const uint32_t tail = (uint32_t)__clz((int)~(tags << 28));  // tagged bytes at the top of the word (4: all of them)
    const uint32_t open = __ballot_sync(0xffffffffu, tags != 0xFu), odd = __ballot_sync(0xffffffffu, tail & 1u);
    const uint32_t lower = open & ((1u << lane) - 1u);
    uint32_t e = lower ? (odd >> (31 - __clz((int)lower))) & 1u : carry;
    carry = open ? (odd >> (31 - __clz((int)open))) & 1u : carry;
    OpsLane r;
    r.ops = r.jumps = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t bj = (word >> (8 * j)) & 0xFFu, nbj = j < 3 ? (word >> (8 * j + 8)) & 0xFFu : nb0;
        const bool tag = (tags >> j) & 1u, start = !e && ((valid >> j) & 1u);
        e = (!e && tag) ? 1u : 0u;
        const bool jump = start && (bj & 0xC0u) == 0x80u;
        const bool full = tag && (!kEdge || pos + j + 1 < end);
        const bool op = start && !jump && (!tag || full);
        r.d[j] = op ? op_delta(bj, nbj, full) : 0;
        if (op) r.ops |= 1u << j;
        if (jump) r.jumps |= 1u << j;
    }
    return r;
}

// The words of a chunk, 128 bytes per step: the lane's word and the byte after it.
struct OpsReader {
    const uint32_t *row;
    int row_words, lo, lane;
    uint32_t word;
    __device__ __forceinline__ uint32_t fetch(int g) const {
        const int wi = ((lo + kOpsGroup * g) >> 2) + lane;
        return wi < row_words ? __ldg(row + wi) : 0u;
    }
    __device__ __forceinline__ void init(const uint8_t *payload_row, int64_t pstride, int lo_, int lane_) {
        row = reinterpret_cast<const uint32_t *>(payload_row);  // 16-byte aligned, padded
        row_words = (int)(pstride >> 2); lo = lo_; lane = lane_;
        word = fetch(0);
    }
    // returns the lane's word of group g and sets nb0; prefetches group g + 1
    __device__ __forceinline__ uint32_t step(int g, uint32_t &nb0) {
        const uint32_t cur = word, next = fetch(g + 1);
        const uint32_t up = __shfl_down_sync(0xffffffffu, cur, 1), wrap = __shfl_sync(0xffffffffu, next, 0);
        nb0 = (lane == 31 ? wrap : up) & 0xFFu;
        word = next;
        return cur;
    }
};

__global__ void __launch_bounds__(32 * kOpsWarps) k_dec_ops_scan(const uint8_t *__restrict__ payload, int64_t pstride,
                                                                 const int32_t *__restrict__ plen, OpsChunkSum *__restrict__ cs,
                                                                 int max_chunks, const uint32_t *__restrict__ status) {
    const int i = blockIdx.y, lane = threadIdx.x & 31;
    const int c = blockIdx.x * kOpsWarps + (threadIdx.x >> 5);
    const int n = plen[i], lo = c * kOpsChunk;
    if (status[i] || lo >= n) return;
    const int end = n - 1, hi = lo + kOpsChunk, stop = hi < end ? hi : end;
    const uint8_t *row = payload + (int64_t)i * pstride;
    // state entering the chunk: parity of the run of tagged bytes that ends at lo - 1.  The look-back is bounded
    // (a crafted payload of FULL-tag look-alikes would otherwise make every chunk rescan all that precedes it): when
    // kOpsLookBack bytes do not settle the parity, the chunk is summarised for both entry states and k_dec_ops_link,
    // which walks the chunks in order, picks the true one.
    uint32_t carry = 0;
    bool both = lo > 0;
    for (int back = lo - 1, step = 0; back >= 0 && step < kOpsLookBack / 32; back -= 32, step++) {
        const int p = back - lane;
        const uint32_t t = __ballot_sync(0xffffffffu, p >= 0 && (__ldg(row + p) & 0xF0u) == 0xE0u);
        const int run = __ffs((int)~t) - 1;  // -1: all 32 tagged
        if (run >= 0) { carry ^= (uint32_t)run & 1u; both = false; break; }
    }
    for (int e = 0; e < (both ? 2 : 1); e++) {
        if (both) carry = (uint32_t)e;
        const uint32_t entry = carry;
        OpsReader rd;
        rd.init(row, pstride, lo, lane);
        int cnt = 0, sum = 0;
        for (int g = 0; g < kOpsChunk / kOpsGroup && lo + kOpsGroup * g < stop; g++) {
            uint32_t nb0;
            const uint32_t word = rd.step(g, nb0);
            const int pos0 = lo + kOpsGroup * g;
            const bool edge = pos0 + kOpsGroup > stop || pos0 + kOpsGroup >= end;  // (warp-uniform)
            const OpsLane L = edge ? ops_lane<true>(word, nb0, pos0 + 4 * lane, stop, end, carry, lane)
                                   : ops_lane<false>(word, nb0, pos0 + 4 * lane, stop, end, carry, lane);
            cnt += __popc(L.ops);
            sum += L.d[0] + L.d[1] + L.d[2] + L.d[3];
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { cnt += __shfl_xor_sync(0xffffffffu, cnt, o); sum += __shfl_xor_sync(0xffffffffu, sum, o); }
        if (lane == 0) {
            // a settled chunk stores its one true summary in both entry slots (ops_link_chunks starts from state 0 and
            // follows exit_state, which reproduces `entry` for every chunk); a FULL op begun on the last byte of the
            // chunk ends in the next one
            OpsChunkSum &o = cs[(int64_t)i * max_chunks + c];
            const uint8_t ex = (uint8_t)(hi < end ? carry : 0u);
            for (uint32_t slot = both ? entry : 0u; slot <= (both ? entry : 1u); slot++) { o.nops[slot] = cnt; o.dsum[slot] = sum; o.exit_state[slot] = ex; }
        }
    }
}

// The chunk summaries of an image are chained by a warp: a chunk is a map from its entry state s (0/1) to (exit state,
// ops, delta sum), maps compose associatively, so 32 chunks are linked by one warp scan of such maps and the result is
// applied to the state, op count and value carried in from the chunks before (ops_link_chunks is the sequential
// statement of the same thing).
struct OpsMap { uint32_t next; int n0, n1, d0, d1; };  // next: bit s = exit state for entry state s
__device__ __forceinline__ OpsMap ops_then(const OpsMap &a, const OpsMap &b) {  // a first, then b
    const uint32_t s0 = a.next & 1u, s1 = (a.next >> 1) & 1u;
    OpsMap r;
    r.next = ((b.next >> s0) & 1u) | (((b.next >> s1) & 1u) << 1);
    r.n0 = a.n0 + (s0 ? b.n1 : b.n0); r.d0 = a.d0 + (s0 ? b.d1 : b.d0);
    r.n1 = a.n1 + (s1 ? b.n1 : b.n0); r.d1 = a.d1 + (s1 ? b.d1 : b.d0);
    return r;
}
constexpr int kLinkWarps = 4;
__global__ void __launch_bounds__(32 * kLinkWarps) k_dec_ops_link(const int32_t *__restrict__ plen, OpsChunkSum *__restrict__ cs_all,
                                                                  int max_chunks, int n_px, uint32_t *__restrict__ status, int B) {
    const int i = blockIdx.x * kLinkWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= B || status[i]) return;
    const int n_chunks = (plen[i] + kOpsChunk - 1) / kOpsChunk;
    OpsChunkSum *cs = cs_all + (int64_t)i * max_chunks;
    uint32_t state = 0;
    int ops = 0, val = 0;
    for (int base = 0; base < n_chunks; base += 32) {
        const int c = base + lane;
        OpsMap m{2u, 0, 0, 0, 0};  // identity past the end
        if (c < n_chunks) {
            const OpsChunkSum v = cs[c];
            m = OpsMap{(uint32_t)(v.exit_state[0] & 1u) | ((uint32_t)(v.exit_state[1] & 1u) << 1), v.nops[0], v.nops[1], v.dsum[0], v.dsum[1]};
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            OpsMap a;
            a.next = __shfl_up_sync(0xffffffffu, m.next, d);
            a.n0 = __shfl_up_sync(0xffffffffu, m.n0, d); a.n1 = __shfl_up_sync(0xffffffffu, m.n1, d);
            a.d0 = __shfl_up_sync(0xffffffffu, m.d0, d); a.d1 = __shfl_up_sync(0xffffffffu, m.d1, d);
            if (lane >= d) m = ops_then(a, m);
        }
        OpsMap ex;  // the chunks of this group before the lane's own
        ex.next = __shfl_up_sync(0xffffffffu, m.next, 1);
        ex.n0 = __shfl_up_sync(0xffffffffu, m.n0, 1); ex.n1 = __shfl_up_sync(0xffffffffu, m.n1, 1);
        ex.d0 = __shfl_up_sync(0xffffffffu, m.d0, 1); ex.d1 = __shfl_up_sync(0xffffffffu, m.d1, 1);
        if (lane == 0) ex = OpsMap{2u, 0, 0, 0, 0};
        if (c < n_chunks) {
            cs[c].entry = (uint8_t)((ex.next >> state) & 1u);
            cs[c].op_base = ops + (state ? ex.n1 : ex.n0);
            cs[c].val_base = val + (state ? ex.d1 : ex.d0);
        }
        const uint32_t tn = __shfl_sync(0xffffffffu, m.next, 31);
        const int t_n = __shfl_sync(0xffffffffu, state ? m.n1 : m.n0, 31), t_d = __shfl_sync(0xffffffffu, state ? m.d1 : m.d0, 31);
        ops += t_n; val += t_d;
        state = (tn >> state) & 1u;
    }
    if (lane == 0 && ops < n_px) status[i] |= 64u;
}

__global__ void __launch_bounds__(32 * kOpsWarps) k_dec_ops_emit(const uint8_t *__restrict__ payload, int64_t pstride,
                                                                 const int32_t *__restrict__ plen, const OpsChunkSum *__restrict__ cs,
                                                                 int max_chunks, int n_px, uint16_t *__restrict__ vals,
                                                                 uint8_t *__restrict__ jflag, uint32_t *__restrict__ status) {
    __shared__ uint16_t stage[kOpsWarps][kOpsGroup];  // values of one step in op order, written out coalesced
    const int i = blockIdx.y, lane = threadIdx.x & 31;
    const int c = blockIdx.x * kOpsWarps + (threadIdx.x >> 5);
    const int n = plen[i], lo = c * kOpsChunk;
    if ((status[i] & ~32u) || lo >= n) return;
    const int end = n - 1, hi = lo + kOpsChunk, stop = hi < end ? hi : end;
    const OpsChunkSum sum = cs[(int64_t)i * max_chunks + c];
    uint32_t carry = sum.entry;
    int k_base = sum.op_base, v_base = sum.val_base;  // ops and value before the step
    uint16_t *va = vals + (int64_t)i * n_px;
    uint8_t *jf = jflag + (int64_t)i * (n_px / kBlock);
    uint16_t *sv = stage[threadIdx.x >> 5];
    uint32_t bad = 0;
    OpsReader rd;
    rd.init(payload + (int64_t)i * pstride, pstride, lo, lane);
    for (int g = 0; g < kOpsChunk / kOpsGroup && lo + kOpsGroup * g < stop && k_base < n_px; g++) {
        uint32_t nb0;
        const uint32_t word = rd.step(g, nb0);
        const int pos0 = lo + kOpsGroup * g;
        const bool edge = pos0 + kOpsGroup > stop || pos0 + kOpsGroup >= end;  // (warp-uniform)
        const OpsLane L = edge ? ops_lane<true>(word, nb0, pos0 + 4 * lane, stop, end, carry, lane)
                               : ops_lane<false>(word, nb0, pos0 + 4 * lane, stop, end, carry, lane);
        // inclusive warp scan of (delta sum << 8 | op count): at most 128 ops per step, so the count never carries
        const int mine = (L.d[0] + L.d[1] + L.d[2] + L.d[3]) * 256 + __popc(L.ops);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const int total = __shfl_sync(0xffffffffu, incl, 31), excl = incl - mine;
        int kk = excl & 0xFF, value = v_base + (excl >> 8);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if ((L.jumps >> j) & 1u) {
                const int k = k_base + kk;
                if (k < n_px) {
                    if ((k & 15) == 0) jf[k >> 4] = (uint8_t)((word >> (8 * j)) & 0x3Fu);  // JUMP distance (core.py:484-494)
                    else bad |= 64u;  // a JUMP inside a unit is never produced by the encoder
                }
            }
            if ((L.ops >> j) & 1u) {
                value += L.d[j];
                if (k_base + kk < n_px && (value < 0 || value > 65535)) bad |= 32u;  // reference: OverflowError (core.py:506,516)
                sv[kk++] = (uint16_t)value;
            }
        }
        __syncwarp();
        const int cnt = min(total & 0xFF, n_px - k_base);
        for (int t = lane; t < cnt; t += 32) va[k_base + t] = sv[t];
        __syncwarp();
        k_base += total & 0xFF;
        v_base += total >> 8;
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    if (bad && lane == 0) atomicOr(status + i, bad);
}

// JUMP flags -> unit table (the first-unconsumed-block walk of core.py:459-464; resolve_units in inflate_core.cuh is
// the sequential statement), one warp per image.  The walk state (next slot q, next block cur, 64-bit window of
// blocks already consumed as partners) is kept warp-uniform.  While the window is empty and the next slots carry
// no JUMP -- the common case -- 32 slots are filled per step; otherwise one unit is resolved per step.
constexpr int kUnitWarps = 4;
__global__ void __launch_bounds__(32 * kUnitWarps) k_dec_units(const uint8_t *__restrict__ jflag, int n_px, int32_t *__restrict__ slotA,
                                                               int32_t *__restrict__ slotB, uint32_t *__restrict__ status, int B) {
    const int i = blockIdx.x * kUnitWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= B || (status[i] & ~32u)) return;
    const int nb = n_px / kBlock;
    const uint8_t *jf = jflag + (int64_t)i * nb;
    int32_t *sa = slotA + (int64_t)i * nb, *sb = slotB + (int64_t)i * nb;
    asm volatile("" : "+l"(jf), "+l"(sa), "+l"(sb));  // (kept in registers: the loops below would re-derive them from the kernel parameters)
    uint64_t win = 0;  // bit k: block cur + k already consumed as a partner
    int cur = 0, q = 0;
    bool bad = false;
    while (q < nb) {
        const int j = q + lane < nb ? jf[q + lane] : 1;  // past the end counts as "not plain"
        const unsigned jm = __ballot_sync(0xffffffffu, j != 0);
        if (win == 0 && !(jm & 1u)) {  // a stretch of plain units on untouched blocks
            const int k = jm ? __ffs((int)jm) - 1 : 32;
            if (lane < k) { sa[q + lane] = cur + lane; sb[q + lane] = -1; }
            q += k; cur += k;
            continue;
        }
        while (win & 1ull) { win >>= 1; cur++; }
        const int j0 = __shfl_sync(0xffffffffu, j, 0);
        if (j0) {
            const int Bk = cur + j0;
            if (Bk >= nb || q + 1 >= nb || ((win >> j0) & 1ull)) { bad = true; break; }
            if (lane < 2) { sa[q + lane] = cur; sb[q + lane] = lane ? (Bk | kSlotSecond) : Bk; }
            win |= 1ull << j0;
            q += 2;
        } else {
            if (lane == 0) { sa[q] = cur; sb[q] = -1; }
            q += 1;
        }
        win >>= 1; cur++;
    }
    if (bad && lane == 0) status[i] |= 64u;
}

__global__ void __launch_bounds__(256) k_dec_scatter(const uint16_t *__restrict__ vals, const int32_t *__restrict__ slotA,
                                                     const int32_t *__restrict__ slotB, const int32_t *__restrict__ order,
                                                     const uint8_t *__restrict__ fl, const uint32_t *__restrict__ status,
                                                     uint16_t *__restrict__ px, int n_px) {
    // four consecutive ops per thread (same unit slot): one 8-byte load of the values, and for a plain unit one
    // 16-byte load of the four traversal entries
    const int i = blockIdx.y;
    const int k0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (k0 >= n_px || (status[i] & ~32u)) return;  // overflow still writes (mod 2^16), everything else leaves zeros
    const int nb = n_px / kBlock;
    const int q = k0 >> 4, t0 = k0 & 15;
    const int32_t A = slotA[(int64_t)i * nb + q], Bf = slotB[(int64_t)i * nb + q];
    const bool fractal = (fl[i] & 1u) != 0;  // fractal flag off: raster order (core.py:426-427)
    const uint2 v2 = *reinterpret_cast<const uint2 *>(vals + (int64_t)i * n_px + k0);
    const uint16_t v[4] = {(uint16_t)(v2.x & 0xFFFFu), (uint16_t)(v2.x >> 16), (uint16_t)(v2.y & 0xFFFFu), (uint16_t)(v2.y >> 16)};
    int32_t r[4];  // raster indices
    if (Bf < 0) {
        const int32_t pos = A * kBlock + t0;  // position along the (un-meshed) traversal
        if (fractal) { const int4 o4 = __ldg(reinterpret_cast<const int4 *>(order + pos)); r[0] = o4.x; r[1] = o4.y; r[2] = o4.z; r[3] = o4.w; }
        else { r[0] = pos; r[1] = pos + 1; r[2] = pos + 2; r[3] = pos + 3; }
    } else {
        const int m0 = t0 + ((Bf & kSlotSecond) ? 16 : 0);  // ops of a meshed pair alternate between its two blocks
        const int32_t pa = A * kBlock + (m0 >> 1), pb = (Bf & ~kSlotSecond) * kBlock + (m0 >> 1);
        r[0] = pa; r[1] = pb; r[2] = pa + 1; r[3] = pb + 1;
        if (fractal) {
#pragma unroll
            for (int e = 0; e < 4; e++) r[e] = __ldg(order + r[e]);
        }
    }
    // The kernel is bound by the sectors its scattered 2-byte stores touch.  Four consecutive positions of a Hilbert
    // traversal are a 2x2 cell, so the four pixels are two horizontal pairs {even index, +1} in one of three
    // matchings: those go out as two 4-byte stores (selects, not branches: neighbouring lanes differ in orientation).
    uint16_t *dst = px + (int64_t)i * n_px;
    const bool m1 = ((r[0] ^ r[1]) == 1) & ((r[2] ^ r[3]) == 1);  // (0,1) (2,3)
    const bool m2 = ((r[1] ^ r[2]) == 1) & ((r[3] ^ r[0]) == 1);  // (1,2) (3,0)
    const bool m3 = ((r[0] ^ r[2]) == 1) & ((r[1] ^ r[3]) == 1);  // (0,2) (1,3): meshed units
    if (m1 | m2 | m3) {
        const int32_t a0 = m2 ? r[1] : r[0], a1 = m1 ? r[1] : r[2];
        const int32_t b0 = m1 ? r[2] : (m2 ? r[3] : r[1]), b1 = m2 ? r[0] : r[3];
        const uint32_t va0 = m2 ? v[1] : v[0], va1 = m1 ? v[1] : v[2];
        const uint32_t vb0 = m1 ? v[2] : (m2 ? v[3] : v[1]), vb1 = m2 ? v[0] : v[3];
        // the even index is the low half of the word
        *reinterpret_cast<uint32_t *>(dst + (a0 & ~1)) = (a0 & 1) ? (va1 | (va0 << 16)) : (va0 | (va1 << 16));
        *reinterpret_cast<uint32_t *>(dst + (b0 & ~1)) = (b0 & 1) ? (vb1 | (vb0 << 16)) : (vb0 | (vb1 << 16));
        (void)a1; (void)b1;
    } else {
#pragma unroll
        for (int e = 0; e < 4; e++) dst[r[e]] = v[e];
    }
}

void launch_decode(const uint8_t *files, const int64_t *offsets, int B, int rows, int cols, const int32_t *order_fractal,
                   uint16_t *px, uint32_t *status, const DecodeWs &ws, cudaStream_t st) {
    if (B == 0) return;
    const int n_px = rows * cols, nb = n_px / kBlock;
    { CCT_KERNEL(m_zero_px, st);
    launch_zero(ws.jflag, (size_t)B * nb, st);
    launch_zero(px, (size_t)B * n_px * sizeof(uint16_t), st);  // np.zeros(size), core.py:436
    }
    { CCT_KERNEL(k_dec_header, st);
    k_dec_header<<<(B + 127) / 128, 128, 0, st>>>(files, offsets, B, rows, cols, status, ws.fl);
    }
    StreamSrc src{files, offsets, 0, nullptr, kHeaderBytes};
    launch_inflate_passes(src, ws.fl, B, ws.iw, ws.payload, ws.pstride, ws.plen, status, st);
    { CCT_KERNEL(k_dec_rawcopy, st);
    k_dec_rawcopy<<<B, 128, 0, st>>>(files, offsets, ws.fl, ws.payload, ws.pstride, ws.plen, status);
    }
    dim3 gs((unsigned)((ws.max_chunks + kOpsWarps - 1) / kOpsWarps), (unsigned)B), ge = gs;
    { CCT_KERNEL(k_dec_ops_scan, st);
    k_dec_ops_scan<<<gs, 32 * kOpsWarps, 0, st>>>(ws.payload, ws.pstride, ws.plen, ws.chunks, ws.max_chunks, status);
    }
    { CCT_KERNEL(k_dec_ops_link, st);
    k_dec_ops_link<<<(B + kLinkWarps - 1) / kLinkWarps, 32 * kLinkWarps, 0, st>>>(ws.plen, ws.chunks, ws.max_chunks, n_px, status, B);
    }
    { CCT_KERNEL(k_dec_ops_emit, st);
    k_dec_ops_emit<<<ge, 32 * kOpsWarps, 0, st>>>(ws.payload, ws.pstride, ws.plen, ws.chunks, ws.max_chunks, n_px, ws.vals, ws.jflag, status);
    }
    { CCT_KERNEL(k_dec_units, st);
    k_dec_units<<<(B + kUnitWarps - 1) / kUnitWarps, 32 * kUnitWarps, 0, st>>>(ws.jflag, n_px, ws.slotA, ws.slotB, status, B);
    }
    dim3 g((unsigned)((n_px / 4 + 255) / 256), (unsigned)B);
    { CCT_KERNEL(k_dec_scatter, st);
    k_dec_scatter<<<g, 256, 0, st>>>(ws.vals, ws.slotA, ws.slotB, order_fractal, ws.fl, status, px, n_px);
    }
}

}  // namespace cct
