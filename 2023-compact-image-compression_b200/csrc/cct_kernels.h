// cct_kernels.h — launcher prototypes shared by the .cu files and the C-ABI layer (cct_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cct_common.cuh"

namespace cct {

struct BlockDesc;
struct SortEnt;
struct Walk;
struct OpsChunkSum;

// ---- launch accounting (cct_api.cu): every kernel launch is counted; with profiling on it is also bracketed
// by CUDA events on its own stream so bench.py can report per-kernel device time from the timed region.
void prof_before(const char *name, cudaStream_t st);
void prof_after(cudaStream_t st);
struct ProfScope {
    cudaStream_t st;
    ProfScope(const char *name, cudaStream_t s) : st(s) { prof_before(name, s); }
    ~ProfScope() { prof_after(st); }
};
#define CCT_KERNEL(name, st) ::cct::ProfScope prof_scope_##name(#name, st)

// ---- cct_encode.cu
// Zero-fill by a kernel.  (cudaMemsetAsync may be carried out by a copy engine: inside the host-buffer entry points it
// then queues behind whatever upload or download is in flight -- measured: every encode pass started only when the
// download of the pass two before it had finished.)
void launch_zero(void *p, size_t bytes, cudaStream_t st);
void launch_curve_pow2(int32_t *d_order, int n, cudaStream_t st);
void launch_iota(int32_t *d_order, int64_t n, cudaStream_t st);
void launch_gather_stats(const uint16_t *px, const int32_t *order, uint16_t *hv, uint8_t *info, int B, int64_t n_px,
                         cudaStream_t st);
void launch_mesh(const uint16_t *hv, const uint8_t *info, uint8_t *jump, uint8_t *taken, int32_t *stats, int B,
                 int64_t n_px, cudaStream_t st);
void launch_emit(const uint16_t *hv, const uint8_t *jump, const uint8_t *taken, uint8_t *payload, int64_t pstride,
                 int32_t *plen, uint32_t *status, int32_t *stats, int B, int64_t n_px, cudaStream_t st);

// ---- cct_deflate.cu
struct DeflateWs {  // device scratch for one pass over B streams of capacity max_len
    SortEnt *ent;        // [B][stride] chain members sorted by (hash, position)
    uint32_t *runs;      // [B][stride] equal bytes after / before every position
    SortEnt *p1;         // [B][stride] (position, bytes) split by the top hash digit (first pass of the sort; aliases mtab + mhi)
    uint32_t *counters;  // [64] work counters of the persistent kernels (zeroed before each)
    uint32_t *thist;     // [B][max_tiles][128] per-tile bin counts, then offsets
    uint32_t *binbase;   // [B][129] bin starts
    int max_tiles;
    uint32_t *mtab, *mhi;  // [B][stride] probe table: entries for chain budget 4096 / (where flagged) 1024
    uint32_t *syms;      // [B][stride] dense LZ77 symbols
    uint32_t *rec;       // [B][stride] symbol a walk standing at each position emits (k_next; slow steps recorded by the walks)
    uint32_t *mask;      // [B][mask_stride] bit p = a symbol starts at position p
    int64_t mask_stride; // words per stream
    Walk *exits;         // [B][max_win] state each speculative window was left with
    int max_win;
    int32_t *blk_end;    // [B][max_blocks]
    int32_t *nsym;       // [B]
    int32_t *nflush;     // [B]
    uint32_t *adler;     // [B]
    BlockDesc *desc;     // [B][max_blocks]
    int64_t stride;      // elements per stream in dist16/mtab/syms (>= max_len, multiple of 16)
    int max_blocks;
};
size_t deflate_ws_bytes(int B, int64_t max_len);
// carve `ws` out of raw memory; returns bytes used
size_t deflate_ws_carve(DeflateWs *ws, void *mem, int B, int64_t max_len);
// zlib streams of in[i*in_stride .. +in_len[i]) -> out[i*out_stride ..], lengths -> out_len[i].
// `out` must be zero-filled by the caller for out_stride*B bytes (the bit packer ORs into it).
void launch_deflate(const uint8_t *in, int64_t in_stride, const int32_t *in_len, int B, int64_t max_len, uint8_t *out,
                    int64_t out_stride, int32_t *out_len, const DeflateWs &ws, cudaStream_t st);

// sorted chain members and probe tables only (stage tap): fills ws.ent, ws.mtab, ws.mhi
void launch_probe_tables(const uint8_t *in, int64_t in_stride, const int32_t *in_len, int B, int64_t max_len,
                         const DeflateWs &ws, cudaStream_t st);

// ---- cct_decode.cu
struct InflateWs {
    uint32_t *syms;      // [B][sym_stride] LZ77 symbols of each stream
    int32_t *nsym;       // [B]
    uint32_t *adler;     // [B] Adler-32 read from each trailer
    int64_t sym_stride;
};
size_t inflate_ws_carve(InflateWs *iw, void *mem, int B, int64_t out_stride);
int launch_inflate(const uint8_t *in, int64_t in_stride, const int32_t *in_len, int B, uint8_t *out, int64_t out_stride,
                   int32_t *out_len, uint32_t *status, const InflateWs &iw, cudaStream_t st);
struct DecodeWs {
    uint8_t *payload;  // [B][pstride]
    int32_t *plen;     // [B]
    uint16_t *vals;    // [B][n_px]
    uint8_t *jflag;    // [B][nb]
    int32_t *slotA, *slotB;  // [B][nb]
    uint8_t *fl;       // [B] transform flags from each header
    InflateWs iw;
    OpsChunkSum *chunks;  // [B][max_chunks] op-stream chunk summaries
    int max_chunks;
    int64_t pstride;
};
size_t decode_ws_bytes(int B, int64_t n_px);
size_t decode_ws_carve(DecodeWs *ws, void *mem, int B, int64_t n_px);
void launch_decode(const uint8_t *files, const int64_t *offsets, int B, int rows, int cols, const int32_t *order_fractal,
                   uint16_t *px, uint32_t *status, const DecodeWs &ws, cudaStream_t st);

// ---- assembly (cct_deflate.cu)
// file lengths -> offsets[first+1 .. first+B] (running from offsets[first]) and copy header + body into the arena
void launch_assemble(const uint8_t *body, int64_t body_stride, const int32_t *body_len, int B, int rows, int cols,
                     uint32_t flags, uint8_t *arena, int64_t arena_cap, int64_t *offsets, int first, uint32_t *status,
                     cudaStream_t st);

}  // namespace cct
