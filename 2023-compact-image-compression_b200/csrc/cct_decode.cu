// cct_decode.cu — decoder: header check, inflate, op stream -> values, JUMP units, scatter to raster.
//
// Reference path restated (reference src/codec/core.py):
//   core.py:385-402   read_header (magic, dims, flags)              -> k_dec_header
//   core.py:420-421   zlib.decompress(file_bytes[13:])              -> k_dec_inflate (one stream per warp slot)
//   core.py:453-520   padded-order walk with JUMP splicing          -> k_dec_ops (ops -> running values, unit table)
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
    ws->max_chunks = (int)(ws->pstride / kOpsChunk + 1);
    ws->chunks = (OpsChunkSum *)take((size_t)B * ws->max_chunks * sizeof(OpsChunkSum));
    return used;
}

// ------------------------------------------------------------------ generic batched inflate
__global__ void __launch_bounds__(32) k_inflate(const uint8_t *__restrict__ in, const int64_t *__restrict__ in_off,
                                                int64_t in_stride, const int32_t *__restrict__ in_len, int hdr_skip,
                                                uint8_t *__restrict__ out, int64_t out_stride, int32_t *__restrict__ out_len,
                                                uint32_t *__restrict__ status) {
    __shared__ InflateTables tb;
    const int i = blockIdx.x;
    if (threadIdx.x != 0) return;
    const uint8_t *src = in + (in_off ? in_off[i] : (int64_t)i * in_stride) + hdr_skip;
    int len = (in_off ? (int)(in_off[i + 1] - in_off[i]) : in_len[i]) - hdr_skip;
    int r = len > 0 ? inflate_stream(src, len, out + (int64_t)i * out_stride, (int)out_stride, tb) : -1;
    if (r < 0) { if (status) status[i] |= 16u; r = 0; }  // CCT_ST_INFLATE
    out_len[i] = r;
}

int launch_inflate(const uint8_t *in, const int64_t *in_off, int64_t in_stride, const int32_t *in_len, int hdr_skip,
                   int B, uint8_t *out, int64_t out_stride, int32_t *out_len, uint32_t *status, cudaStream_t st) {
    if (B == 0) return 0;
    { CCT_KERNEL(k_inflate, st);
    k_inflate<<<B, 32, 0, st>>>(in, in_off, in_stride, in_len, hdr_skip, out, out_stride, out_len, status);
    }
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

__global__ void __launch_bounds__(32) k_dec_inflate(const uint8_t *__restrict__ files, const int64_t *__restrict__ offsets,
                                                    const uint8_t *__restrict__ fl, uint8_t *__restrict__ payload,
                                                    int64_t pstride, int32_t *__restrict__ plen, uint32_t *__restrict__ status) {
    __shared__ InflateTables tb;
    const int i = blockIdx.x, lane = threadIdx.x;
    if (status[i]) { if (lane == 0) plen[i] = 0; return; }
    const uint8_t *src = files + offsets[i] + kHeaderBytes;
    const int len = (int)(offsets[i + 1] - offsets[i]) - kHeaderBytes;
    uint8_t *dst = payload + (int64_t)i * pstride;
    if (fl[i] & 4u) {
        if (lane != 0) return;
        int r = len > 0 ? inflate_stream(src, len, dst, (int)pstride, tb) : -1;
        if (r < 0) { status[i] |= 16u; r = 0; }
        plen[i] = r;
    } else {  // deflate flag off: the payload follows the header as is (core.py:420)
        int n = len < (int)pstride ? len : (int)pstride;
        for (int k = lane; k < n; k += 32) dst[k] = src[k];
        if (lane == 0) { plen[i] = n; if (n < len) status[i] |= 64u; }
    }
}

// Op stream -> values, in three parallel passes over fixed chunks (see inflate_core.cuh).
__global__ void __launch_bounds__(128) k_dec_ops_scan(const uint8_t *__restrict__ payload, int64_t pstride,
                                                      const int32_t *__restrict__ plen, OpsChunkSum *__restrict__ cs,
                                                      int max_chunks, const uint32_t *__restrict__ status) {
    const int i = blockIdx.y;
    const int t = blockIdx.x * 128 + threadIdx.x;  // chunk * 2 + entry state
    const int c = t >> 1, e = t & 1;
    const int n = plen[i];
    if (status[i] || c * kOpsChunk >= n) return;
    int nops, dsum, ex;
    ops_scan_chunk(payload + (int64_t)i * pstride, c * kOpsChunk, (c + 1) * kOpsChunk, n - 1, e, &nops, &dsum, &ex);
    OpsChunkSum &o = cs[(int64_t)i * max_chunks + c];
    o.nops[e] = nops; o.dsum[e] = dsum; o.exit_state[e] = (uint8_t)ex;
}

__global__ void __launch_bounds__(128) k_dec_ops_link(const int32_t *__restrict__ plen, OpsChunkSum *__restrict__ cs,
                                                      int max_chunks, int n_px, uint32_t *__restrict__ status, int B) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= B || status[i]) return;
    const int n_chunks = (plen[i] + kOpsChunk - 1) / kOpsChunk;
    status[i] |= ops_link_chunks(cs + (int64_t)i * max_chunks, n_chunks, n_px);
}

__global__ void __launch_bounds__(128) k_dec_ops_emit(const uint8_t *__restrict__ payload, int64_t pstride,
                                                      const int32_t *__restrict__ plen, const OpsChunkSum *__restrict__ cs,
                                                      int max_chunks, int n_px, uint16_t *__restrict__ vals,
                                                      uint8_t *__restrict__ jflag, uint32_t *__restrict__ status) {
    const int i = blockIdx.y;
    const int c = blockIdx.x * 128 + threadIdx.x;
    const int n = plen[i];
    if ((status[i] & ~32u) || c * kOpsChunk >= n) return;
    const uint32_t st = ops_emit_chunk(payload + (int64_t)i * pstride, c * kOpsChunk, (c + 1) * kOpsChunk, n - 1,
                                       cs[(int64_t)i * max_chunks + c], n_px, vals + (int64_t)i * n_px,
                                       jflag + (int64_t)i * (n_px / kBlock));
    if (st) atomicOr(status + i, st);
}

// JUMP flags -> unit table, one thread per image (sequential first-unconsumed-block walk, core.py:459-464)
__global__ void __launch_bounds__(32) k_dec_units(const uint8_t *__restrict__ jflag, int n_px, int32_t *__restrict__ slotA,
                                                  int32_t *__restrict__ slotB, uint32_t *__restrict__ status, int B) {
    const int i = blockIdx.x * 32 + threadIdx.x;
    if (i >= B || (status[i] & ~32u)) return;
    const int nb = n_px / kBlock;
    const uint32_t st = resolve_units(jflag + (int64_t)i * nb, nb, slotA + (int64_t)i * nb, slotB + (int64_t)i * nb);
    if (st) status[i] |= st;
}

__global__ void __launch_bounds__(256) k_dec_scatter(const uint16_t *__restrict__ vals, const int32_t *__restrict__ slotA,
                                                     const int32_t *__restrict__ slotB, const int32_t *__restrict__ order,
                                                     const uint8_t *__restrict__ fl, const uint32_t *__restrict__ status,
                                                     uint16_t *__restrict__ px, int n_px) {
    const int i = blockIdx.y;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= n_px || (status[i] & ~32u)) return;  // overflow still writes (mod 2^16), everything else leaves zeros
    const int nb = n_px / kBlock;
    const int32_t *sa = slotA + (int64_t)i * nb, *sb = slotB + (int64_t)i * nb;
    int q = k >> 4, t = k & 15;
    int32_t A = sa[q], Bf = sb[q];
    int32_t pos;  // position along the (un-meshed) traversal
    if (Bf < 0) pos = A * kBlock + t;
    else {
        int m = t + ((Bf & kSlotSecond) ? 16 : 0);
        pos = ((m & 1) ? (Bf & ~kSlotSecond) : A) * kBlock + (m >> 1);
    }
    int32_t r = (fl[i] & 1u) ? __ldg(order + pos) : pos;  // fractal flag off: raster order (core.py:426-427)
    px[(int64_t)i * n_px + r] = vals[(int64_t)i * n_px + k];
}

void launch_decode(const uint8_t *files, const int64_t *offsets, int B, int rows, int cols, const int32_t *order_fractal,
                   uint16_t *px, uint32_t *status, const DecodeWs &ws, cudaStream_t st) {
    if (B == 0) return;
    const int n_px = rows * cols, nb = n_px / kBlock;
    cudaMemsetAsync(ws.jflag, 0, (size_t)B * nb, st);
    cudaMemsetAsync(px, 0, (size_t)B * n_px * sizeof(uint16_t), st);  // np.zeros(size), core.py:436
    { CCT_KERNEL(k_dec_header, st);
    k_dec_header<<<(B + 127) / 128, 128, 0, st>>>(files, offsets, B, rows, cols, status, ws.fl);
    }
    { CCT_KERNEL(k_dec_inflate, st);
    k_dec_inflate<<<B, 32, 0, st>>>(files, offsets, ws.fl, ws.payload, ws.pstride, ws.plen, status);
    }
    dim3 gs((unsigned)((2 * ws.max_chunks + 127) / 128), (unsigned)B), ge((unsigned)((ws.max_chunks + 127) / 128), (unsigned)B);
    { CCT_KERNEL(k_dec_ops_scan, st);
    k_dec_ops_scan<<<gs, 128, 0, st>>>(ws.payload, ws.pstride, ws.plen, ws.chunks, ws.max_chunks, status);
    }
    { CCT_KERNEL(k_dec_ops_link, st);
    k_dec_ops_link<<<(B + 127) / 128, 128, 0, st>>>(ws.plen, ws.chunks, ws.max_chunks, n_px, status, B);
    }
    { CCT_KERNEL(k_dec_ops_emit, st);
    k_dec_ops_emit<<<ge, 128, 0, st>>>(ws.payload, ws.pstride, ws.plen, ws.chunks, ws.max_chunks, n_px, ws.vals, ws.jflag, status);
    }
    { CCT_KERNEL(k_dec_units, st);
    k_dec_units<<<(B + 31) / 32, 32, 0, st>>>(ws.jflag, n_px, ws.slotA, ws.slotB, status, B);
    }
    dim3 g((unsigned)((n_px + 255) / 256), (unsigned)B);
    { CCT_KERNEL(k_dec_scatter, st);
    k_dec_scatter<<<g, 256, 0, st>>>(ws.vals, ws.slotA, ws.slotB, order_fractal, ws.fl, status, px, n_px);
    }
}

}  // namespace cct
