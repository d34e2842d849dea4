// cct_api.cu — the C ABI of libcompact_b200.so (include/compact_b200.h): argument checking, workspace
// carving, traversal-table cache, pass scheduling over the caller's stream, host-buffer entry points.
// No codec arithmetic lives here; see cct_encode.cu / cct_deflate.cu / cct_decode.cu.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/compact_b200.h"
#include "cct_kernels.h"
#include "deflate_core.cuh"

using namespace cct;

// ---------------------------------------------------------------- launch accounting / profiling
namespace {
std::mutex g_prof_mu;
bool g_prof_on = false;
long long g_launches = 0;
struct ProfRec { const char *name; cudaEvent_t e0, e1; };
std::vector<ProfRec> g_prof;
thread_local cudaEvent_t g_pending_e1 = nullptr;
}  // namespace

namespace cct {
void prof_before(const char *name, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_launches++;
    g_pending_e1 = nullptr;
    if (!g_prof_on) return;
    ProfRec r{name, nullptr, nullptr};
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, st);
    g_pending_e1 = r.e1;
    g_prof.push_back(r);
}
void prof_after(cudaStream_t st) {
    if (g_pending_e1) { cudaEventRecord(g_pending_e1, st); g_pending_e1 = nullptr; }
}
}  // namespace cct

namespace {

thread_local std::string g_cuda_err;

bool cuda_ok(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    g_cuda_err = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}
#define CK(call) do { if (!cuda_ok((call), #call)) return CCT_E_CUDA; } while (0)

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------- generalized Hilbert curve (host)
// Same traversal as the reference's GeneralizedHilbertCurve (curve.py:61-138): a rectangle described by
// its major axis a and orthogonal axis b is either emitted as a straight run (one side of length 1),
// split in two along a long major axis, or split in three (up / across / back down).  Written with an
// explicit work stack; halving uses floor division like Python's // on negative components.
struct Rect { int x, y, ax, ay, bx, by; };
inline int isgn(int v) { return (v > 0) - (v < 0); }
inline int floor_half(int v) { return v >= 0 ? v / 2 : -((1 - v) / 2); }

void gilbert_order(int width, int height, std::vector<int32_t> &order) {
    order.clear();
    order.reserve((size_t)width * height);
    std::vector<Rect> stack;
    if (width >= height) stack.push_back({0, 0, width, 0, 0, height});
    else stack.push_back({0, 0, 0, height, width, 0});
    while (!stack.empty()) {
        Rect r = stack.back();
        stack.pop_back();
        const int w = std::abs(r.ax + r.ay), h = std::abs(r.bx + r.by);
        const int dax = isgn(r.ax), day = isgn(r.ay), dbx = isgn(r.bx), dby = isgn(r.by);
        if (h == 1 || w == 1) {
            const int steps = h == 1 ? w : h, sx = h == 1 ? dax : dbx, sy = h == 1 ? day : dby;
            int x = r.x, y = r.y;
            for (int i = 0; i < steps; i++, x += sx, y += sy) order.push_back(y * width + x);
            continue;
        }
        int ax2 = floor_half(r.ax), ay2 = floor_half(r.ay), bx2 = floor_half(r.bx), by2 = floor_half(r.by);
        const int w2 = std::abs(ax2 + ay2), h2 = std::abs(bx2 + by2);
        if (2 * w > 3 * h) {
            if ((w2 & 1) && w > 2) { ax2 += dax; ay2 += day; }
            // pushed in reverse so the first part is processed first
            stack.push_back({r.x + ax2, r.y + ay2, r.ax - ax2, r.ay - ay2, r.bx, r.by});
            stack.push_back({r.x, r.y, ax2, ay2, r.bx, r.by});
        } else {
            if ((h2 & 1) && h > 2) { bx2 += dbx; by2 += dby; }
            stack.push_back({r.x + (r.ax - dax) + (bx2 - dbx), r.y + (r.ay - day) + (by2 - dby), -bx2, -by2, -(r.ax - ax2), -(r.ay - ay2)});
            stack.push_back({r.x + bx2, r.y + by2, r.ax, r.ay, r.bx - bx2, r.by - by2});
            stack.push_back({r.x, r.y, bx2, by2, ax2, ay2});
        }
    }
}

// ---------------------------------------------------------------- traversal-table cache
std::mutex g_mu;
std::map<std::tuple<int, int, int, int>, int32_t *> g_orders;  // (device, rows, cols, fractal)

int get_order(int rows, int cols, bool fractal, cudaStream_t st, const int32_t **out) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_tuple(dev, rows, cols, fractal ? 1 : 0);
    auto it = g_orders.find(key);
    if (it != g_orders.end()) { *out = it->second; return CCT_OK; }
    const int64_t n = (int64_t)rows * cols;
    int32_t *d = nullptr;
    CK(cudaMalloc(&d, (size_t)std::max<int64_t>(n, 1) * sizeof(int32_t)));
    if (!fractal) launch_iota(d, n, st);
    else if (rows == cols && rows >= 2 && (rows & (rows - 1)) == 0) launch_curve_pow2(d, rows, st);
    else {
        std::vector<int32_t> h;
        gilbert_order(rows, cols, h);  // reference: width = image.shape[0] (core.py:179,235)
        CK(cudaMemcpyAsync(d, h.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));  // one-time: the table may be used from any stream afterwards
    g_orders[key] = d;
    *out = d;
    return CCT_OK;
}

// ---------------------------------------------------------------- encode workspace
struct EncodeWs {
    uint16_t *hv; uint8_t *info, *jump, *taken, *payload, *zbuf;
    int32_t *plen, *zlen;
    int64_t pstride, zstride;
    DeflateWs dws;
};

size_t encode_ws_carve(EncodeWs *ws, void *mem, int B, int64_t n_px, uint32_t flags) {
    char *p = (char *)mem;
    size_t used = 0;
    auto take = [&](size_t bytes) { void *r = p ? p + used : nullptr; used += (size_t)align_up((int64_t)bytes, 256); return r; };
    const int64_t nb = n_px / kBlock;
    ws->pstride = payload_bound(n_px) + 16;
    ws->zstride = align_up(deflate_bound(ws->pstride), 16);
    ws->hv = (uint16_t *)take((size_t)B * n_px * 2);
    ws->info = (uint8_t *)take((size_t)B * nb);
    ws->jump = (uint8_t *)take((size_t)B * nb);
    ws->taken = (uint8_t *)take((size_t)B * nb);
    ws->payload = (uint8_t *)take((size_t)B * ws->pstride);
    ws->plen = (int32_t *)take((size_t)B * 4);
    if (flags & CCT_FLAG_DEFLATE) {
        ws->zbuf = (uint8_t *)take((size_t)B * ws->zstride);
        ws->zlen = (int32_t *)take((size_t)B * 4);
        used += deflate_ws_carve(&ws->dws, p ? p + used : nullptr, B, ws->pstride);
    } else { ws->zbuf = nullptr; ws->zlen = nullptr; }
    return used;
}

int check_shape(int B, int rows, int cols) {
    if (B < 0 || rows <= 0 || cols <= 0 || rows > 65535 || cols > 65535) return CCT_E_ARG;
    const int64_t n = (int64_t)rows * cols;
    if (n * 2 > 400000000000LL) return CCT_E_TOO_LARGE;  // core.py:218-219
    if (n % kBlock) return CCT_E_SHAPE;                  // core.py:245 reshape -> ValueError
    if (n > (1LL << 26)) return CCT_E_TOO_LARGE;         // 32-bit payload offsets in the kernels
    return CCT_OK;
}

int fit_batch(size_t ws_bytes, int B, const std::function<size_t(int)> &need) {
    if (B == 0) return 0;
    if (need(B) <= ws_bytes) return B;
    if (need(1) > ws_bytes) return 0;
    int lo = 1, hi = B;  // need(lo) fits, need(hi) does not
    while (hi - lo > 1) {
        int mid = lo + (hi - lo) / 2;
        if (need(mid) <= ws_bytes) lo = mid; else hi = mid;
    }
    return lo;
}
}  // namespace

extern "C" {

int cct_version(void) { return 100; }

const char *cct_strerror(int code) {
    switch (code) {
        case CCT_OK: return "ok";
        case CCT_E_ARG: return "bad argument";
        case CCT_E_SHAPE: return "rows*cols is not a multiple of the block size 16";
        case CCT_E_WORKSPACE: return "workspace too small for one image";
        case CCT_E_CUDA: return "CUDA runtime error";
        case CCT_E_TOO_LARGE: return "image too large";
        case CCT_E_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
        case CCT_E_CAPACITY: return "output arena too small";
        default: return "unknown error";
    }
}

const char *cct_last_cuda_error(void) { return g_cuda_err.c_str(); }

long long cct_launch_count(int reset) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    long long v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

int cct_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
    return CCT_OK;
}

// Synchronises the device, then writes one line per kernel name: "<name> <launches> <total_ms>\n".
int cct_profile_read(char *buf, size_t cap) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    CK(cudaDeviceSynchronize());
    std::map<std::string, std::pair<int, double>> agg;
    for (auto &r : g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { auto &a = agg[r.name]; a.first++; a.second += ms; }
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    g_prof.clear();
    std::string out;
    for (auto &kv : agg) { char line[160]; snprintf(line, sizeof line, "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second); out += line; }
    if (buf && cap) { size_t n = std::min(cap - 1, out.size()); memcpy(buf, out.data(), n); buf[n] = 0; }
    return CCT_OK;
}

int cct_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

size_t cct_payload_bound(int rows, int cols) { return (size_t)payload_bound((int64_t)rows * cols) + 16; }
size_t cct_file_bound(int rows, int cols) {
    return (size_t)(kHeaderBytes + deflate_bound((int64_t)cct_payload_bound(rows, cols)));
}
size_t cct_deflate_bound(size_t len) { return (size_t)align_up(deflate_bound((int64_t)len), 16); }

size_t cct_encode_workspace_bytes(int B, int rows, int cols, uint32_t flags) {
    if (check_shape(B, rows, cols) != CCT_OK) return 0;
    EncodeWs ws;
    return encode_ws_carve(&ws, nullptr, std::max(B, 1), (int64_t)rows * cols, flags);
}

size_t cct_decode_workspace_bytes(int B, int rows, int cols) {
    if (check_shape(B, rows, cols) != CCT_OK) return 0;
    return decode_ws_bytes(std::max(B, 1), (int64_t)rows * cols);
}

}  // extern "C"

namespace {
// Called on the host around every pass of a batch call: the host-buffer entry points use them to make the
// compute stream wait for the pass's input copy and to start the pass's output copy.
struct PassHooks {
    std::function<int(int first, int count)> before, after;
    int max_pass = 0;  // 0 = as many images per pass as the workspace holds
};

int encode_impl(const uint16_t *d_px, int B, int rows, int cols, uint32_t flags, uint8_t *d_out, size_t out_cap,
                int64_t *d_offsets, uint32_t *d_status, int32_t *d_stats, void *d_ws, size_t ws_bytes, cudaStream_t st,
                const PassHooks *hooks);
int decode_impl(const uint8_t *d_in, const int64_t *d_offsets, int B, int rows, int cols, uint16_t *d_px,
                uint32_t *d_status, void *d_ws, size_t ws_bytes, cudaStream_t st, const PassHooks *hooks);
}  // namespace

extern "C" {

int cct_encode_batch(const uint16_t *d_px, int B, int rows, int cols, uint32_t flags, uint8_t *d_out, size_t out_cap,
                     int64_t *d_offsets, uint32_t *d_status, int32_t *d_stats, void *d_ws, size_t ws_bytes,
                     void *stream) {
    return encode_impl(d_px, B, rows, cols, flags, d_out, out_cap, d_offsets, d_status, d_stats, d_ws, ws_bytes,
                       (cudaStream_t)stream, nullptr);
}

int cct_decode_batch(const uint8_t *d_in, const int64_t *d_offsets, int B, int rows, int cols, uint16_t *d_px,
                     uint32_t *d_status, void *d_ws, size_t ws_bytes, void *stream) {
    return decode_impl(d_in, d_offsets, B, rows, cols, d_px, d_status, d_ws, ws_bytes, (cudaStream_t)stream, nullptr);
}

}  // extern "C"

namespace {

int encode_impl(const uint16_t *d_px, int B, int rows, int cols, uint32_t flags, uint8_t *d_out, size_t out_cap,
                int64_t *d_offsets, uint32_t *d_status, int32_t *d_stats, void *d_ws, size_t ws_bytes, cudaStream_t st,
                const PassHooks *hooks) {
    int rc = check_shape(B, rows, cols);
    if (rc != CCT_OK) return rc;
    if (!d_offsets || (B > 0 && (!d_px || !d_out || !d_status || !d_ws))) return CCT_E_ARG;
    if (cct_device_count() == 0) return CCT_E_NO_DEVICE;
    const int64_t n_px = (int64_t)rows * cols, nb = n_px / kBlock;
    CK(cudaMemsetAsync(d_offsets, 0, sizeof(int64_t), st));
    if (B == 0) return CCT_OK;
    int Bc = fit_batch(ws_bytes, B, [&](int b) { EncodeWs w; return encode_ws_carve(&w, nullptr, b, n_px, flags); });
    if (Bc == 0) return CCT_E_WORKSPACE;
    if (hooks && hooks->max_pass > 0) Bc = std::min(Bc, hooks->max_pass);
    const int32_t *order = nullptr;
    rc = get_order(rows, cols, (flags & CCT_FLAG_FRACTAL) != 0, st, &order);
    if (rc != CCT_OK) return rc;
    if (d_stats) CK(cudaMemsetAsync(d_stats, 0, (size_t)B * 4 * sizeof(int32_t), st));
    EncodeWs ws;
    encode_ws_carve(&ws, d_ws, Bc, n_px, flags);
    for (int first = 0; first < B; first += Bc) {
        const int b = std::min(Bc, B - first);
        const uint16_t *px = d_px + (int64_t)first * n_px;
        int32_t *stats = d_stats ? d_stats + (int64_t)first * 4 : nullptr;
        if (hooks && hooks->before && (rc = hooks->before(first, b)) != CCT_OK) return rc;
        { CCT_KERNEL(m_zero_jump, st); CK(cudaMemsetAsync(ws.jump, 0, (size_t)b * nb, st)); CK(cudaMemsetAsync(ws.taken, 0, (size_t)b * nb, st)); }
        launch_gather_stats(px, order, ws.hv, ws.info, b, n_px, st);
        if (flags & CCT_FLAG_SEGMENTATION) launch_mesh(ws.hv, ws.info, ws.jump, ws.taken, stats, b, n_px, st);
        launch_emit(ws.hv, ws.jump, ws.taken, ws.payload, ws.pstride, ws.plen, d_status + first, stats, b, n_px, st);
        if (flags & CCT_FLAG_DEFLATE) {
            { CCT_KERNEL(m_zero_zbuf, st); CK(cudaMemsetAsync(ws.zbuf, 0, (size_t)b * ws.zstride, st)); }
            launch_deflate(ws.payload, ws.pstride, ws.plen, b, ws.pstride, ws.zbuf, ws.zstride, ws.zlen, ws.dws, st);
            launch_assemble(ws.zbuf, ws.zstride, ws.zlen, b, rows, cols, flags, d_out, (int64_t)out_cap, d_offsets, first,
                            d_status, st);
        } else {
            launch_assemble(ws.payload, ws.pstride, ws.plen, b, rows, cols, flags, d_out, (int64_t)out_cap, d_offsets, first,
                            d_status, st);
        }
        CK(cudaGetLastError());
        if (hooks && hooks->after && (rc = hooks->after(first, b)) != CCT_OK) return rc;
    }
    return CCT_OK;
}

int decode_impl(const uint8_t *d_in, const int64_t *d_offsets, int B, int rows, int cols, uint16_t *d_px,
                uint32_t *d_status, void *d_ws, size_t ws_bytes, cudaStream_t st, const PassHooks *hooks) {
    int rc = check_shape(B, rows, cols);
    if (rc != CCT_OK) return rc;
    if (B > 0 && (!d_in || !d_offsets || !d_px || !d_status || !d_ws)) return CCT_E_ARG;
    if (cct_device_count() == 0) return CCT_E_NO_DEVICE;
    if (B == 0) return CCT_OK;
    const int64_t n_px = (int64_t)rows * cols;
    int Bc = fit_batch(ws_bytes, B, [&](int b) { return decode_ws_bytes(b, n_px); });
    if (Bc == 0) return CCT_E_WORKSPACE;
    if (hooks && hooks->max_pass > 0) Bc = std::min(Bc, hooks->max_pass);
    const int32_t *order = nullptr;
    rc = get_order(rows, cols, true, st, &order);
    if (rc != CCT_OK) return rc;
    DecodeWs ws;
    decode_ws_carve(&ws, d_ws, Bc, n_px);
    for (int first = 0; first < B; first += Bc) {
        const int b = std::min(Bc, B - first);
        if (hooks && hooks->before && (rc = hooks->before(first, b)) != CCT_OK) return rc;
        launch_decode(d_in, d_offsets + first, b, rows, cols, order, d_px + (int64_t)first * n_px, d_status + first, ws, st);
        CK(cudaGetLastError());
        if (hooks && hooks->after && (rc = hooks->after(first, b)) != CCT_OK) return rc;
    }
    return CCT_OK;
}

}  // namespace

extern "C" {

int cct_read_header(const uint8_t *f, size_t len, int *rows, int *cols, int *channels, int *bytes_per_channel,
                    uint32_t *flags) {
    if (!f || len < (size_t)kHeaderBytes || memcmp(f, "pact", 4) != 0) return CCT_E_ARG;
    if (rows) *rows = (f[4] << 8) | f[5];
    if (cols) *cols = (f[6] << 8) | f[7];
    if (channels) *channels = f[8];
    if (bytes_per_channel) *bytes_per_channel = f[9];
    if (flags) *flags = (f[10] ? CCT_FLAG_FRACTAL : 0u) | (f[11] ? CCT_FLAG_SEGMENTATION : 0u) | (f[12] ? CCT_FLAG_DEFLATE : 0u);
    return CCT_OK;
}

// ---------------------------------------------------------------- stage taps
int cct_debug_encode_stages(const uint16_t *d_px, int rows, int cols, uint32_t flags, int32_t *d_order, uint8_t *d_jump,
                            uint8_t *d_payload, int32_t *d_payload_len, void *d_ws, size_t ws_bytes, void *stream) {
    int rc = check_shape(1, rows, cols);
    if (rc != CCT_OK) return rc;
    if (!d_px || !d_ws) return CCT_E_ARG;
    if (cct_device_count() == 0) return CCT_E_NO_DEVICE;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_px = (int64_t)rows * cols, nb = n_px / kBlock;
    EncodeWs ws;
    if (encode_ws_carve(&ws, nullptr, 1, n_px, flags & ~CCT_FLAG_DEFLATE) > ws_bytes) return CCT_E_WORKSPACE;
    encode_ws_carve(&ws, d_ws, 1, n_px, flags & ~CCT_FLAG_DEFLATE);
    const int32_t *order = nullptr;
    rc = get_order(rows, cols, (flags & CCT_FLAG_FRACTAL) != 0, st, &order);
    if (rc != CCT_OK) return rc;
    CK(cudaMemsetAsync(ws.jump, 0, (size_t)nb, st));
    CK(cudaMemsetAsync(ws.taken, 0, (size_t)nb, st));
    launch_gather_stats(d_px, order, ws.hv, ws.info, 1, n_px, st);
    if (flags & CCT_FLAG_SEGMENTATION) launch_mesh(ws.hv, ws.info, ws.jump, ws.taken, nullptr, 1, n_px, st);
    launch_emit(ws.hv, ws.jump, ws.taken, ws.payload, ws.pstride, ws.plen, nullptr, nullptr, 1, n_px, st);
    if (d_order) CK(cudaMemcpyAsync(d_order, order, (size_t)n_px * 4, cudaMemcpyDeviceToDevice, st));
    if (d_jump) CK(cudaMemcpyAsync(d_jump, ws.jump, (size_t)nb, cudaMemcpyDeviceToDevice, st));
    if (d_payload) CK(cudaMemcpyAsync(d_payload, ws.payload, (size_t)ws.pstride, cudaMemcpyDeviceToDevice, st));
    if (d_payload_len) CK(cudaMemcpyAsync(d_payload_len, ws.plen, 4, cudaMemcpyDeviceToDevice, st));
    CK(cudaGetLastError());
    return CCT_OK;
}

size_t cct_deflate_workspace_bytes(int B, size_t max_len) { return deflate_ws_bytes(std::max(B, 1), (int64_t)max_len); }

int cct_deflate_batch(const uint8_t *d_in, size_t in_stride, const int32_t *d_in_len, int B, size_t max_len,
                      uint8_t *d_out, size_t out_stride, int32_t *d_out_len, void *d_ws, size_t ws_bytes, void *stream) {
    if (B < 0 || (B > 0 && (!d_in || !d_in_len || !d_out || !d_out_len || !d_ws))) return CCT_E_ARG;
    if (out_stride % 4 || in_stride % 16 || out_stride < cct_deflate_bound(max_len) || max_len > (1u << 27)) return CCT_E_ARG;
    if (cct_device_count() == 0) return CCT_E_NO_DEVICE;
    if (deflate_ws_bytes(std::max(B, 1), (int64_t)max_len) > ws_bytes) return CCT_E_WORKSPACE;
    if (B == 0) return CCT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DeflateWs ws;
    deflate_ws_carve(&ws, d_ws, B, (int64_t)max_len);
    CK(cudaMemsetAsync(d_out, 0, (size_t)B * out_stride, st));
    launch_deflate(d_in, (int64_t)in_stride, d_in_len, B, (int64_t)max_len, d_out, (int64_t)out_stride, d_out_len, ws, st);
    CK(cudaGetLastError());
    return CCT_OK;
}

int cct_debug_chains(const uint8_t *d_in, size_t in_stride, const int32_t *d_in_len, int B, uint16_t *d_dist16,
                     size_t stride, void *stream) {
    if (B < 0 || (B > 0 && (!d_in || !d_in_len || !d_dist16)) || in_stride % 16) return CCT_E_ARG;
    if (cct_device_count() == 0) return CCT_E_NO_DEVICE;
    launch_chains(d_in, (int64_t)in_stride, d_in_len, B, d_dist16, (int64_t)stride, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return CCT_OK;
}

size_t cct_inflate_workspace_bytes(int B, size_t out_stride) {
    InflateWs iw;
    return inflate_ws_carve(&iw, nullptr, std::max(B, 1), (int64_t)out_stride);
}

int cct_inflate_batch(const uint8_t *d_in, size_t in_stride, const int32_t *d_in_len, int B, uint8_t *d_out,
                      size_t out_stride, int32_t *d_out_len, uint32_t *d_status, void *d_ws, size_t ws_bytes, void *stream) {
    if (B < 0 || (B > 0 && (!d_in || !d_in_len || !d_out || !d_out_len || !d_status || !d_ws))) return CCT_E_ARG;
    if (cct_device_count() == 0) return CCT_E_NO_DEVICE;
    if (cct_inflate_workspace_bytes(B, out_stride) > ws_bytes) return CCT_E_WORKSPACE;
    if (B == 0) return CCT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(d_status, 0, (size_t)B * 4, st));
    InflateWs iw;
    inflate_ws_carve(&iw, d_ws, B, (int64_t)out_stride);
    launch_inflate(d_in, (int64_t)in_stride, d_in_len, B, d_out, (int64_t)out_stride, d_out_len, d_status, iw, st);
    CK(cudaGetLastError());
    return CCT_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- host-buffer entry points
namespace {

struct HostCtx {
    int dev = -1;
    cudaStream_t st = nullptr, copy = nullptr;
    std::vector<cudaEvent_t> events;
    cudaEvent_t event(size_t i) {
        while (events.size() <= i) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); events.push_back(e); }
        return events[i];
    }
    void *buf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[6] = {0, 0, 0, 0, 0, 0};
    int ensure(int slot, size_t bytes) {
        if (cap[slot] >= bytes) return CCT_OK;
        if (buf[slot]) { cudaFree(buf[slot]); buf[slot] = nullptr; cap[slot] = 0; }
        bytes = (size_t)align_up((int64_t)bytes, 1 << 20);
        CK(cudaMalloc(&buf[slot], bytes));
        cap[slot] = bytes;
        return CCT_OK;
    }
};

std::mutex g_ctx_mu;
std::map<int, HostCtx> g_ctx;

int get_ctx(int device, HostCtx **out) {
    if (cct_device_count() == 0) return CCT_E_NO_DEVICE;
    if (device < 0) CK(cudaGetDevice(&device));
    CK(cudaSetDevice(device));
    HostCtx &c = g_ctx[device];
    if (c.dev < 0) {
        c.dev = device;
        CK(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c.copy, cudaStreamNonBlocking));
    }
    *out = &c;
    return CCT_OK;
}

// workspace budget for the host entry points: as many images per pass as fit in ~40 % of free memory
size_t ws_budget(size_t want) {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return want;
    return std::min(want, (size_t)(free_b * 0.4));
}

}  // namespace

extern "C" {

int cct_encode_host(const uint16_t *h_px, int B, int rows, int cols, uint32_t flags, uint8_t *h_out, size_t out_cap,
                    int64_t *h_offsets, uint32_t *h_status, int32_t *h_stats, int device) {
    int rc = check_shape(B, rows, cols);
    if (rc != CCT_OK) return rc;
    if (!h_offsets || (B > 0 && (!h_px || !h_out || !h_status))) return CCT_E_ARG;
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    HostCtx *c = nullptr;
    rc = get_ctx(device, &c);
    if (rc != CCT_OK) return rc;
    h_offsets[0] = 0;
    if (B == 0) return CCT_OK;
    const size_t n_px = (size_t)rows * cols;
    const size_t arena = std::min(out_cap, (size_t)B * cct_file_bound(rows, cols));
    const size_t one = cct_encode_workspace_bytes(1, rows, cols, flags);
    size_t ws_bytes = std::max(one, ws_budget(cct_encode_workspace_bytes(B, rows, cols, flags)));
    if ((rc = c->ensure(0, (size_t)B * n_px * 2))) return rc;
    if ((rc = c->ensure(1, std::max<size_t>(arena, 16)))) return rc;
    if ((rc = c->ensure(2, (size_t)(B + 1) * 8 + (size_t)B * 4 + (size_t)B * 16))) return rc;
    if ((rc = c->ensure(3, ws_bytes))) return rc;
    uint16_t *d_px = (uint16_t *)c->buf[0];
    uint8_t *d_out = (uint8_t *)c->buf[1];
    int64_t *d_off = (int64_t *)c->buf[2];
    uint32_t *d_status = (uint32_t *)(d_off + B + 1);
    int32_t *d_stats = (int32_t *)(d_status + B);
    // pixels go up in slices on the copy stream; each pass of the encoder waits only for its own slices, so the
    // upload of pass k+1 overlaps the kernels of pass k
    const int slice = std::max(1, std::min(B, 128));
    const int n_slices = (B + slice - 1) / slice;
    for (int k = 0; k < n_slices; k++) {
        const size_t lo = (size_t)k * slice, cnt = std::min<size_t>(slice, (size_t)B - lo);
        CK(cudaMemcpyAsync(d_px + lo * n_px, h_px + lo * n_px, cnt * n_px * 2, cudaMemcpyHostToDevice, c->copy));
        CK(cudaEventRecord(c->event(k), c->copy));
    }
    PassHooks hooks;  // max_pass stays 0: the per-stream kernels need the whole batch resident to fill the GPU
    hooks.before = [&](int first, int count) -> int {
        const int last_slice = (first + count - 1) / slice;
        CK(cudaStreamWaitEvent(c->st, c->event(last_slice), 0));  // slices are recorded in order on one stream
        return CCT_OK;
    };
    rc = encode_impl(d_px, B, rows, cols, flags, d_out, arena, d_off, d_status, d_stats, c->buf[3], c->cap[3], c->st, &hooks);
    if (rc != CCT_OK) { cudaStreamSynchronize(c->copy); cudaStreamSynchronize(c->st); return rc; }
    CK(cudaMemcpyAsync(h_offsets, d_off, (size_t)(B + 1) * 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(h_status, d_status, (size_t)B * 4, cudaMemcpyDeviceToHost, c->st));
    if (h_stats) CK(cudaMemcpyAsync(h_stats, d_stats, (size_t)B * 16, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    const int64_t total = h_offsets[B];
    if ((size_t)total > out_cap) return CCT_E_CAPACITY;
    CK(cudaMemcpyAsync(h_out, d_out, (size_t)total, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return CCT_OK;
}

int cct_decode_host(const uint8_t *h_in, const int64_t *h_offsets, int B, int rows, int cols, uint16_t *h_px,
                    uint32_t *h_status, int device) {
    int rc = check_shape(B, rows, cols);
    if (rc != CCT_OK) return rc;
    if (B > 0 && (!h_in || !h_offsets || !h_px || !h_status)) return CCT_E_ARG;
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    HostCtx *c = nullptr;
    rc = get_ctx(device, &c);
    if (rc != CCT_OK) return rc;
    if (B == 0) return CCT_OK;
    const size_t n_px = (size_t)rows * cols;
    const int64_t base = h_offsets[0], total = h_offsets[B] - base;
    if (total < 0) return CCT_E_ARG;
    const size_t one = cct_decode_workspace_bytes(1, rows, cols);
    size_t ws_bytes = std::max(one, ws_budget(cct_decode_workspace_bytes(B, rows, cols)));
    if ((rc = c->ensure(0, (size_t)B * n_px * 2))) return rc;
    if ((rc = c->ensure(1, std::max<size_t>((size_t)total, 16)))) return rc;
    if ((rc = c->ensure(2, (size_t)(B + 1) * 8 + (size_t)B * 4))) return rc;
    if ((rc = c->ensure(4, ws_bytes))) return rc;
    uint16_t *d_px = (uint16_t *)c->buf[0];
    uint8_t *d_in = (uint8_t *)c->buf[1];
    int64_t *d_off = (int64_t *)c->buf[2];
    uint32_t *d_status = (uint32_t *)(d_off + B + 1);
    std::vector<int64_t> rel((size_t)B + 1);
    for (int i = 0; i <= B; i++) rel[i] = h_offsets[i] - base;
    CK(cudaMemcpyAsync(d_in, h_in + base, (size_t)total, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(d_off, rel.data(), (size_t)(B + 1) * 8, cudaMemcpyHostToDevice, c->st));
    // pixels come down pass by pass on the copy stream while the next pass decodes
    PassHooks hooks;  // (splitting the batch to overlap more of the copy was measured: 26 % slower end to end)
    int n_ev = 0;
    hooks.after = [&](int first, int count) -> int {
        cudaEvent_t e = c->event(n_ev++);
        CK(cudaEventRecord(e, c->st));
        CK(cudaStreamWaitEvent(c->copy, e, 0));
        CK(cudaMemcpyAsync(h_px + (size_t)first * n_px, d_px + (size_t)first * n_px, (size_t)count * n_px * 2,
                           cudaMemcpyDeviceToHost, c->copy));
        return CCT_OK;
    };
    rc = decode_impl(d_in, d_off, B, rows, cols, d_px, d_status, c->buf[4], c->cap[4], c->st, &hooks);
    if (rc != CCT_OK) { cudaStreamSynchronize(c->copy); cudaStreamSynchronize(c->st); return rc; }
    CK(cudaMemcpyAsync(h_status, d_status, (size_t)B * 4, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    CK(cudaStreamSynchronize(c->copy));
    return CCT_OK;
}

void cct_release(void) {
    {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        for (auto &kv : g_ctx) {
            cudaSetDevice(kv.first);
            for (int i = 0; i < 6; i++) if (kv.second.buf[i]) cudaFree(kv.second.buf[i]);
            for (cudaEvent_t e : kv.second.events) cudaEventDestroy(e);
            if (kv.second.st) cudaStreamDestroy(kv.second.st);
            if (kv.second.copy) cudaStreamDestroy(kv.second.copy);
        }
        g_ctx.clear();
    }
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto &kv : g_orders) { cudaSetDevice(std::get<0>(kv.first)); cudaFree(kv.second); }
    g_orders.clear();
}

}  // extern "C"
