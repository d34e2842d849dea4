/*
 * compact_b200.h — C ABI of libcompact_b200.so: the B200 (sm_100a) replacement for the CompaCT codec hot path.
 *
 * The reference has no FFI layer: its hot path is the Python classes `Encoder` / `Decoder`
 * (reference src/codec/core.py:170-365 and core.py:367-543).  These entry points are what a
 * ctypes binding of those two classes calls; the host-side mirror in
 * 2023-compact-image-compression_b200/codec/core.py is exactly such a binding.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch/CUDA types in signatures (`stream` is a cudaStream_t
 *     passed as void*, NULL = the legacy default stream);
 *   - "d_" pointers are device memory owned by the caller, "h_" pointers are host memory;
 *   - every function returns 0 (CCT_OK) or a negative CCT_E* code and never throws;
 *   - functions taking a stream only enqueue work: results are valid after the stream is synchronised;
 *   - re-entrant per (device, stream); the only shared state is a per-device cache of traversal
 *     tables keyed by (rows, cols), guarded by a mutex.
 *
 * Image geometry: `rows, cols` = image.shape of the reference's 2-D uint16 array.  The reference calls
 * them width, height (core.py:179) and stores them in that order in the header (core.py:199-200).
 */
#ifndef COMPACT_B200_H
#define COMPACT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* transform flags = the three header bytes 10..12 (core.py:207-209) */
#define CCT_FLAG_FRACTAL 1u      /* config.encoder.transforms.fractal      */
#define CCT_FLAG_SEGMENTATION 2u /* config.encoder.transforms.segmentation */
#define CCT_FLAG_DEFLATE 4u      /* config.encoder.deflate_compression     */
#define CCT_FLAGS_DEFAULT 7u

#define CCT_HEADER_BYTES 13 /* core.py:193-210 */
#define CCT_BLOCK_SIZE 16   /* config.json:9   */

/* return codes */
#define CCT_OK 0
#define CCT_E_ARG (-1)       /* bad argument (NULL pointer, B < 0, ...) */
#define CCT_E_SHAPE (-2)     /* rows*cols not a multiple of 16: reference raises ValueError (core.py:245) */
#define CCT_E_WORKSPACE (-3) /* workspace smaller than cct_*_workspace_bytes(1, ...) */
#define CCT_E_CUDA (-4)      /* a CUDA runtime call failed; cct_last_cuda_error() has the text */
#define CCT_E_TOO_LARGE (-5) /* raw size > 4e11 bytes: reference raises MemoryError (core.py:218-219) */
#define CCT_E_NO_DEVICE (-6) /* no CUDA device: there is no CPU fallback */
#define CCT_E_CAPACITY (-7)  /* output arena too small (host-buffer entry points only) */

/* per-image status bits written to d_status[i] */
#define CCT_ST_DOMAIN 1u     /* encode: some step left [-2047,+2048] or first pixel > 2048: the reference
                                decoder cannot reconstruct this image; bytes are still exact */
#define CCT_ST_CAPACITY 2u   /* encode: the output arena was too small for this image */
#define CCT_ST_BAD_MAGIC 4u  /* decode: reference raises ValueError (core.py:388-389) */
#define CCT_ST_BAD_SHAPE 8u  /* decode: header dimensions differ from the rows/cols passed */
#define CCT_ST_INFLATE 16u   /* decode: corrupt zlib stream / Adler-32 mismatch (reference: zlib.error) */
#define CCT_ST_OVERFLOW 32u  /* decode: a value left [0,65535]: reference raises OverflowError (core.py:506,516) */
#define CCT_ST_TRUNCATED 64u /* decode: payload ended early */

/* per-image counters written to d_stats[4*i .. 4*i+3] by encode (Encoder.info, core.py:186,317,322) */
#define CCT_STAT_DELTA 0
#define CCT_STAT_FULL 1
#define CCT_STAT_JUMP 2
#define CCT_STAT_DIFFICULT 3

int cct_version(void);
const char *cct_strerror(int code);
const char *cct_last_cuda_error(void);
int cct_device_count(void);

/* Worst-case .cct bytes for one image (header + zlib stream of the worst-case payload). */
size_t cct_file_bound(int rows, int cols);

/* Device workspace needed to process `B` images per pass.  A smaller workspace is accepted as long as
 * it holds one image: the batch is then processed in several passes on the same stream. */
size_t cct_encode_workspace_bytes(int B, int rows, int cols, uint32_t flags);
size_t cct_decode_workspace_bytes(int B, int rows, int cols);

/*
 * Encoder.encode() (core.py:212-365) for B images of identical shape.
 *   d_px       [B][rows][cols] uint16, C order (image.tobytes(), core.py:177)
 *   d_out      byte arena; image i's .cct file is d_out[d_offsets[i] .. d_offsets[i+1])
 *   out_cap    arena capacity in bytes (B * cct_file_bound() always suffices)
 *   d_offsets  [B+1] int64, written by the call (d_offsets[0] = 0)
 *   d_status   [B] uint32 CCT_ST_* bits, written by the call
 *   d_stats    [B][4] int32 counters or NULL
 */
int cct_encode_batch(const uint16_t *d_px, int B, int rows, int cols, uint32_t flags, uint8_t *d_out, size_t out_cap,
                     int64_t *d_offsets, uint32_t *d_status, int32_t *d_stats, void *d_ws, size_t ws_bytes,
                     void *stream);

/*
 * Decoder.decode() (core.py:404-543) for B files holding images of identical shape.
 *   d_in       byte arena; file i is d_in[d_offsets[i] .. d_offsets[i+1])
 *   d_px       [B][rows][cols] uint16 output (output.tobytes(), core.py:543)
 * The transform flags are read from each file's header (core.py:397-401), as the reference does.
 * Batches of 512 images and more are decoded as two halves that overlap on `stream` and on a library-owned stream
 * (forked from and joined to `stream` with events: work queued on `stream` after the call still sees the finished
 * pixels); cct_decode_workspace_bytes covers both halves.  CCT_DEC_SPLIT=0 in the environment turns this off.
 */
int cct_decode_batch(const uint8_t *d_in, const int64_t *d_offsets, int B, int rows, int cols, uint16_t *d_px,
                     uint32_t *d_status, void *d_ws, size_t ws_bytes, void *stream);

/* Parse a .cct header on the host (core.py:385-402).  Returns CCT_OK or CCT_E_ARG (bad magic / short). */
int cct_read_header(const uint8_t *h_file, size_t len, int *rows, int *cols, int *channels, int *bytes_per_channel,
                    uint32_t *flags);

/*
 * Host-buffer entry points (what `Encoder(config, ndarray).encode()` binds to): pixels and files live in
 * host memory; the call owns a per-device context (device buffers, workspace, streams), page-locks pageable caller
 * buffers for its duration (cudaHostRegister; already pinned memory is used as it is; CCT_HOST_PIN=0 disables this),
 * copies in, runs the kernels, copies out and synchronises.  device < 0 selects the current device; the caller's
 * current device is restored on return, and calls for different devices may run concurrently from different threads.
 *   h_out / out_cap as above; returns CCT_E_CAPACITY if the arena is too small (h_offsets is still valid
 *   so the caller can resize and retry).
 */
int cct_encode_host(const uint16_t *h_px, int B, int rows, int cols, uint32_t flags, uint8_t *h_out, size_t out_cap,
                    int64_t *h_offsets, uint32_t *h_status, int32_t *h_stats, int device);
int cct_decode_host(const uint8_t *h_in, const int64_t *h_offsets, int B, int rows, int cols, uint16_t *h_px,
                    uint32_t *h_status, int device);
/* Release the per-device contexts and traversal-table caches created by the calls above. */
void cct_release(void);

/*
 * Stage taps for the parity tests (tests/ compare each intermediate with the oracle's).  Same inputs as
 * cct_encode_batch for ONE image; every output pointer is device memory and may be NULL.
 *   d_order     [N]  int32  traversal order (curve.py:61-138)
 *   d_jump      [N/16] uint8 JUMP distance of each block or 0 (cluster.py:166)
 *   d_payload   [cct_payload_bound] uint8 op bytes + EOF (core.py:274-330); *d_payload_len int32
 */
size_t cct_payload_bound(int rows, int cols);
int cct_debug_encode_stages(const uint16_t *d_px, int rows, int cols, uint32_t flags, int32_t *d_order,
                            uint8_t *d_jump, uint8_t *d_payload, int32_t *d_payload_len, void *d_ws, size_t ws_bytes,
                            void *stream);
/* zlib-level-9 stream (bit-exact with libz 1.3 compress2(...,9)) of B byte strings of equal capacity:
 * string i = d_in[i*in_stride .. +d_in_len[i]); stream i is written at d_out[i*out_stride ..], length d_out_len[i].
 * d_in must be 16-byte aligned and in_stride a multiple of 16. */
size_t cct_deflate_workspace_bytes(int B, size_t max_len);
size_t cct_deflate_bound(size_t len);
int cct_deflate_batch(const uint8_t *d_in, size_t in_stride, const int32_t *d_in_len, int B, size_t max_len,
                      uint8_t *d_out, size_t out_stride, int32_t *d_out_len, void *d_ws, size_t ws_bytes, void *stream);
/* Stage tap: the match finder of the DEFLATE stage (zlib's longest_match, reference src/codec/core.py:340 via libz).
 *   d_ent [B][stride] 8-byte entries: the chain members of string i (positions 1 .. len-3) sorted by (3-byte hash,
 *         position); low word = position | 0x80000000 (starts a byte triple c,c,c) | 0x40000000 (first of its hash
 *         bucket), high word = the four bytes at the position (0 past the end)
 *   d_lo  [B][stride] uint32 per position: longest_match from a cold start with chain budget 4096 as
 *         len | dist << 9 (0 = none); bit 30 set = the budget-1024 result differs and is in d_hi at the same index
 *   d_hi  [B][stride] uint32, defined only where d_lo has bit 30 set
 * Any of the three may be NULL; stride (elements) >= max_len; the workspace is cct_deflate_workspace_bytes(B, max_len). */
int cct_debug_match_tables(const uint8_t *d_in, size_t in_stride, const int32_t *d_in_len, int B, size_t max_len,
                           uint64_t *d_ent, uint32_t *d_lo, uint32_t *d_hi, size_t stride, void *d_ws, size_t ws_bytes,
                           void *stream);
/* Host logic tap (needs no device): the pass schedule cct_encode_host (decode = 0) / cct_decode_host (decode = 1)
 * use for B images of n_px pixels.  Writes the first image of every pass and then B to starts[0 .. n_pass] and returns
 * n_pass, or CCT_E_ARG when `cap` entries do not hold them. */
int cct_debug_host_schedule(int B, size_t n_px, int decode, int32_t *starts, int cap);
/* RFC 1950 inflate of B streams (any conformant zlib stream). d_status[i] gets CCT_ST_INFLATE on error
 * (invalid code, output longer than out_stride, truncated input, Adler-32 mismatch). */
size_t cct_inflate_workspace_bytes(int B, size_t out_stride);
int cct_inflate_batch(const uint8_t *d_in, size_t in_stride, const int32_t *d_in_len, int B, uint8_t *d_out,
                      size_t out_stride, int32_t *d_out_len, uint32_t *d_status, void *d_ws, size_t ws_bytes,
                      void *stream);

/* Launch accounting for bench.py: number of kernels launched by this library since the last reset, and
 * optional per-kernel device timing (CUDA events recorded on the launching stream around every kernel).
 * cct_profile_read synchronises the device and writes "<kernel> <launches> <total_ms>\n" lines. */
long long cct_launch_count(int reset);
int cct_profile_enable(int on);
/* Time only the kernels named here (comma- or space-separated; NULL or "" = every kernel again). */
int cct_profile_select(const char *names);
int cct_profile_read(char *buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* COMPACT_B200_H */
