// cct_api.cu — the C ABI of libcompact_b200.so (include/compact_b200.h): argument checking, workspace
// carving, traversal-table cache, pass scheduling over the caller's stream, host-buffer entry points.
// No codec arithmetic lives here; see cct_encode.cu / cct_deflate.cu / cct_decode.cu.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
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
std::vector<std::string> g_prof_only;  // non-empty: only these kernels are bracketed by events
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
    if (!g_prof_only.empty() && std::find(g_prof_only.begin(), g_prof_only.end(), name) == g_prof_only.end()) return;
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

// Images per pass: several kernels index the image with gridDim.y, which is limited to 65535.
constexpr int kMaxPass = 65535;
int fit_batch(size_t ws_bytes, int B, const std::function<size_t(int)> &need) {
    if (B == 0) return 0;
    B = std::min(B, kMaxPass);
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

// Restricts the event timing to the kernels named in `names` (separated by commas or spaces; NULL or "" = all): two
// event records around every launch of a step cost about 2 % of its time, so a benchmark times everything during its
// warm-up and only the kernels that matter inside the timed region.
int cct_profile_select(const char *names) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_only.clear();
    std::string cur;
    for (const char *q = names ? names : ""; ; q++) {
        if (*q && *q != ',' && *q != ' ') { cur += *q; continue; }
        if (!cur.empty()) { g_prof_only.push_back(cur); cur.clear(); }
        if (!*q) break;
    }
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
    const int b = std::max(B, 1);
    const int64_t n_px = (int64_t)rows * cols;
    return std::max(decode_ws_bytes(b, n_px), 2 * decode_ws_bytes((b + 1) / 2, n_px));  // one pass, or two halves side by side
}

}  // extern "C"

namespace {
// Called on the host around every pass of a batch call: the host-buffer entry points use them to make the
// compute stream wait for the pass's input copy and to start the pass's output copy.
struct PassHooks {
    std::function<int(int pass, int first, int count, cudaStream_t st)> before, after;
    int max_pass = 0;  // 0 = as many images per pass as the workspace holds
    const std::vector<int> *starts = nullptr;  // optional schedule: pass k covers [starts[k], starts[k + 1]), each <= max_pass
    // optional second lane: odd passes run on `alt_stream` with their own workspace, so that the kernels of two
    // neighbouring passes overlap on the GPU (per-stream kernels of one pass leave SMs idle that the other fills)
    cudaStream_t alt_stream = nullptr;
    void *alt_ws = nullptr;
    // ... and a third (passes 2, 5, 8 ...), used only together with the second
    cudaStream_t alt_stream2 = nullptr;
    void *alt_ws2 = nullptr;
    cudaEvent_t ready = nullptr, carry[3] = {nullptr, nullptr, nullptr};  // set-up done on the main stream; offsets of a pass written
    int lanes() const { return alt_stream && alt_ws ? (alt_stream2 && alt_ws2 ? 3 : 2) : 1; }
    cudaStream_t lane_stream(int lane, cudaStream_t main) const { return lane == 0 ? main : (lane == 1 ? alt_stream : alt_stream2); }
};

// Library-owned second stream of a device for cct_decode_batch (one per device, created on first use).
constexpr int kSplitMin = 512;  // images; below that every kernel is latency-bound and a split only adds launches
struct SideLane { cudaStream_t stream = nullptr; cudaEvent_t ready = nullptr; std::mutex mu; };  // mu: one call at a time queues work on it
std::map<int, SideLane> g_side;  // guarded by g_mu
int get_side(SideLane **out) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    SideLane &s = g_side[dev];
    if (!s.stream) {
        CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
    }
    *out = &s;
    return CCT_OK;
}

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
    // A large batch is decoded as two halves on two streams (the caller's and a library-owned one, forked from and
    // joined to it with events): the bit-serial stages are bound by the latency of one stream, not by the GPU --
    // k_lz_expand keeps one warp per stream busy, k_dec_units one warp per image -- so the halves fill each other's
    // idle issue slots (the half that finished k_inflate_sym expands while the other still decodes Huffman codes).
    // CCT_DEC_SPLIT=0 turns it off.
    static const int parts = [] { const char *e = getenv("CCT_DEC_SPLIT"); return e ? atoi(e) : 2; }();
    if (parts > 1 && B >= kSplitMin && (B + parts - 1) / parts <= kMaxPass && check_shape(B, rows, cols) == CCT_OK && d_ws) {
        const int each = (B + parts - 1) / parts;
        const size_t one = decode_ws_bytes(each, (int64_t)rows * cols);
        SideLane *side = nullptr;
        if (2 * one <= ws_bytes && get_side(&side) == CCT_OK) {
            std::vector<int> starts;
            for (int k = 0; k < parts; k++) starts.push_back(std::min(B, k * each));
            starts.push_back(B);
            PassHooks hooks;
            hooks.max_pass = each;
            hooks.starts = &starts;
            hooks.alt_stream = side->stream;
            hooks.alt_ws = (char *)d_ws + one;
            hooks.ready = side->ready;
            std::lock_guard<std::mutex> lk(side->mu);  // the fork/join event is shared: host threads queue their calls in turn
            return decode_impl(d_in, d_offsets, B, rows, cols, d_px, d_status, d_ws, one, (cudaStream_t)stream, &hooks);
        }
    }
    return decode_impl(d_in, d_offsets, B, rows, cols, d_px, d_status, d_ws, ws_bytes, (cudaStream_t)stream, nullptr);
}

}  // extern "C"

namespace {

int encode_impl(const uint16_t *d_px, int B, int rows, int cols, uint32_t flags, uint8_t *d_out, size_t out_cap,
                int64_t *d_offsets, uint32_t *d_status, int32_t *d_stats, void *d_ws, size_t ws_bytes, cudaStream_t st_main,
                const PassHooks *hooks) {
    cudaStream_t st = st_main;
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
    EncodeWs wss[3];
    encode_ws_carve(&wss[0], d_ws, Bc, n_px, flags);
    const int n_lanes = hooks ? hooks->lanes() : 1;
    const bool two = n_lanes > 1;
    if (two) {
        encode_ws_carve(&wss[1], hooks->alt_ws, Bc, n_px, flags);
        if (n_lanes > 2) encode_ws_carve(&wss[2], hooks->alt_ws2, Bc, n_px, flags);
        CK(cudaEventRecord(hooks->ready, st_main));
        for (int l = 1; l < n_lanes; l++) CK(cudaStreamWaitEvent(hooks->lane_stream(l, st_main), hooks->ready, 0));
    }
    if (hooks && hooks->starts) for (size_t k = 0; k + 1 < hooks->starts->size(); k++) if ((*hooks->starts)[k + 1] - (*hooks->starts)[k] > Bc) return CCT_E_WORKSPACE;
    int pass = 0;
    for (int first = 0, b = 0; first < B; first += b, pass++) {
        b = hooks && hooks->starts ? (*hooks->starts)[pass + 1] - first : std::min(Bc, B - first);
        const int lane = pass % n_lanes;
        const EncodeWs &ws = wss[lane];
        cudaStream_t st = two ? hooks->lane_stream(lane, st_main) : st_main;
        const uint16_t *px = d_px + (int64_t)first * n_px;
        int32_t *stats = d_stats ? d_stats + (int64_t)first * 4 : nullptr;
        if (hooks && hooks->before && (rc = hooks->before(pass, first, b, st)) != CCT_OK) return rc;
        { CCT_KERNEL(m_zero_jump, st); launch_zero(ws.jump, (size_t)b * nb, st); launch_zero(ws.taken, (size_t)b * nb, st); }
        launch_gather_stats(px, order, ws.hv, ws.info, b, n_px, st);
        if (flags & CCT_FLAG_SEGMENTATION) launch_mesh(ws.hv, ws.info, ws.jump, ws.taken, stats, b, n_px, st);
        launch_emit(ws.hv, ws.jump, ws.taken, ws.payload, ws.pstride, ws.plen, d_status + first, stats, b, n_px, st);
        const uint8_t *body = ws.payload; int64_t body_stride = ws.pstride; const int32_t *body_len = ws.plen;
        if (flags & CCT_FLAG_DEFLATE) {
            { CCT_KERNEL(m_zero_zbuf, st); launch_zero(ws.zbuf, (size_t)b * ws.zstride, st); }
            launch_deflate(ws.payload, ws.pstride, ws.plen, b, ws.pstride, ws.zbuf, ws.zstride, ws.zlen, ws.dws, st);
            body = ws.zbuf; body_stride = ws.zstride; body_len = ws.zlen;
        }
        // the running offset comes from the previous pass, which ran on the other stream
        if (two && pass > 0) CK(cudaStreamWaitEvent(st, hooks->carry[(pass - 1) % n_lanes], 0));
        launch_assemble(body, body_stride, body_len, b, rows, cols, flags, d_out, (int64_t)out_cap, d_offsets, first, d_status, st);
        if (two) CK(cudaEventRecord(hooks->carry[pass % n_lanes], st));
        CK(cudaGetLastError());
        if (hooks && hooks->after && (rc = hooks->after(pass, first, b, st)) != CCT_OK) return rc;
    }
    for (int l = 1; l < n_lanes; l++) {  // the caller's stream ends after everything
        CK(cudaEventRecord(hooks->ready, hooks->lane_stream(l, st_main)));
        CK(cudaStreamWaitEvent(st_main, hooks->ready, 0));
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
    DecodeWs wss[3];
    decode_ws_carve(&wss[0], d_ws, Bc, n_px);
    const int n_lanes = hooks ? hooks->lanes() : 1;
    if (n_lanes > 1) {
        decode_ws_carve(&wss[1], hooks->alt_ws, Bc, n_px);
        if (n_lanes > 2) decode_ws_carve(&wss[2], hooks->alt_ws2, Bc, n_px);
        CK(cudaEventRecord(hooks->ready, st));
        for (int l = 1; l < n_lanes; l++) CK(cudaStreamWaitEvent(hooks->lane_stream(l, st), hooks->ready, 0));
    }
    if (hooks && hooks->starts) for (size_t k = 0; k + 1 < hooks->starts->size(); k++) if ((*hooks->starts)[k + 1] - (*hooks->starts)[k] > Bc) return CCT_E_WORKSPACE;
    int pass = 0;
    for (int first = 0, b = 0; first < B; first += b, pass++) {
        b = hooks && hooks->starts ? (*hooks->starts)[pass + 1] - first : std::min(Bc, B - first);
        const int lane = pass % n_lanes;
        cudaStream_t ps = n_lanes > 1 ? hooks->lane_stream(lane, st) : st;
        if (hooks && hooks->before && (rc = hooks->before(pass, first, b, ps)) != CCT_OK) return rc;
        launch_decode(d_in, d_offsets + first, b, rows, cols, order, d_px + (int64_t)first * n_px, d_status + first, wss[lane], ps);
        CK(cudaGetLastError());
        if (hooks && hooks->after && (rc = hooks->after(pass, first, b, ps)) != CCT_OK) return rc;
    }
    for (int l = 1; l < n_lanes; l++) {
        CK(cudaEventRecord(hooks->ready, hooks->lane_stream(l, st)));
        CK(cudaStreamWaitEvent(st, hooks->ready, 0));
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

int cct_debug_match_tables(const uint8_t *d_in, size_t in_stride, const int32_t *d_in_len, int B, size_t max_len,
                           uint64_t *d_ent, uint32_t *d_lo, uint32_t *d_hi, size_t stride, void *d_ws, size_t ws_bytes,
                           void *stream) {
    if (B < 0 || (B > 0 && (!d_in || !d_in_len || !d_ws)) || in_stride % 16 || max_len > (1u << 27) || stride < max_len) return CCT_E_ARG;
    if (cct_device_count() == 0) return CCT_E_NO_DEVICE;
    if (deflate_ws_bytes(std::max(B, 1), (int64_t)max_len) > ws_bytes) return CCT_E_WORKSPACE;
    if (B == 0) return CCT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DeflateWs ws;
    deflate_ws_carve(&ws, d_ws, B, (int64_t)max_len);
    launch_probe_tables(d_in, (int64_t)in_stride, d_in_len, B, (int64_t)max_len, ws, st);
    CK(cudaGetLastError());
    if (d_ent) CK(cudaMemcpy2DAsync(d_ent, stride * 8, ws.ent, (size_t)ws.stride * 8, max_len * 8, (size_t)B, cudaMemcpyDeviceToDevice, st));
    if (d_lo) CK(cudaMemcpy2DAsync(d_lo, stride * 4, ws.mtab, (size_t)ws.stride * 4, max_len * 4, (size_t)B, cudaMemcpyDeviceToDevice, st));
    if (d_hi) CK(cudaMemcpy2DAsync(d_hi, stride * 4, ws.mhi, (size_t)ws.stride * 4, max_len * 4, (size_t)B, cudaMemcpyDeviceToDevice, st));
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
    std::mutex mu;  // one host-buffer call at a time per device; calls to different devices run side by side
    int dev = -1;
    cudaStream_t st = nullptr, st2 = nullptr, st3 = nullptr, up = nullptr, down = nullptr;  // kernels (up to three lanes), host->device copies, device->host copies
    cudaEvent_t ready = nullptr, carry[3] = {nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> events;
    cudaEvent_t event(size_t i) {
        while (events.size() <= i) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); events.push_back(e); }
        return events[i];
    }
    void *buf[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[7] = {0, 0, 0, 0, 0, 0, 0};
    int ensure(int slot, size_t bytes) {
        if (cap[slot] >= bytes) return CCT_OK;
        if (buf[slot]) { cudaFree(buf[slot]); buf[slot] = nullptr; cap[slot] = 0; }
        bytes = (size_t)align_up((int64_t)bytes, 1 << 20);
        CK(cudaMalloc(&buf[slot], bytes));
        cap[slot] = bytes;
        return CCT_OK;
    }
    // CCT_HOST_TRACE=1: device timeline of a host-buffer call (copies and passes), printed to stderr at its end
    std::vector<std::pair<std::string, cudaEvent_t>> marks;
    bool tracing = false;
    std::vector<cudaEvent_t> mark_pool;  // (created before the call starts: creating an event may synchronise)
    void mark(const char *what, int k, cudaStream_t s) {
        if (!tracing || marks.size() >= mark_pool.size()) return;
        cudaEvent_t e = mark_pool[marks.size()];
        cudaEventRecord(e, s);
        marks.emplace_back(std::string(what) + " " + std::to_string(k), e);
    }
    void trace_begin(cudaStream_t s) {
        const char *e = getenv("CCT_HOST_TRACE");
        tracing = e && atoi(e) > 0;
        if (tracing) while (mark_pool.size() < 512) { cudaEvent_t ev; cudaEventCreate(&ev); mark_pool.push_back(ev); }
        mark("begin", 0, s);
    }
    void trace_end(const char *call) {
        if (!tracing) return;
        for (size_t i = 1; i < marks.size(); i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, marks[0].second, marks[i].second);
            fprintf(stderr, "[cct trace] %s %-14s %8.3f ms\n", call, marks[i].first.c_str(), ms);
        }
        marks.clear();
        tracing = false;
    }
    void drain() { cudaStreamSynchronize(up); cudaStreamSynchronize(st); cudaStreamSynchronize(st2); cudaStreamSynchronize(st3); cudaStreamSynchronize(down); }
    void more_lanes(PassHooks &h, int n_lanes) const {  // workspaces of lanes 1 and 2: slots 5 and 6
        h.alt_stream = st2; h.alt_ws = buf[5]; h.ready = ready;
        if (n_lanes > 2) { h.alt_stream2 = st3; h.alt_ws2 = buf[6]; }
        for (int l = 0; l < 3; l++) h.carry[l] = carry[l];
    }
};

std::mutex g_ctx_mu;  // guards the map only
std::map<int, HostCtx> g_ctx;

// The host entry points switch to the requested device and give the caller's current device back when they return.
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Pageable caller memory is page-locked for the duration of a call (cudaHostRegister): copies from pageable memory go
// through the driver's staging buffer and block the host, which would run the passes of the pipeline one after the
// other.  Memory that is already pinned (cudaHostAlloc, torch pin_memory) is left alone.  CCT_HOST_PIN=0 turns it off.
struct HostPin {
    void *p = nullptr;
    HostPin(const void *ptr, size_t bytes) {
        static const bool on = [] { const char *e = getenv("CCT_HOST_PIN"); return !e || atoi(e) != 0; }();
        if (!on || !ptr || bytes < ((size_t)1 << 20)) return;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return; }
        if (a.type != cudaMemoryTypeUnregistered) return;
        if (cudaHostRegister(const_cast<void *>(ptr), bytes, cudaHostRegisterDefault) == cudaSuccess) p = const_cast<void *>(ptr);
        else cudaGetLastError();
    }
    ~HostPin() { if (p) cudaHostUnregister(p); }
    HostPin(const HostPin &) = delete;
    HostPin &operator=(const HostPin &) = delete;
};

int get_ctx(int device, HostCtx **out) {
    if (cct_device_count() == 0) return CCT_E_NO_DEVICE;
    if (device < 0) CK(cudaGetDevice(&device));
    CK(cudaSetDevice(device));
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    HostCtx &c = g_ctx[device];
    if (c.dev < 0) {
        c.dev = device;
        CK(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c.st2, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c.st3, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c.ready, cudaEventDisableTiming));
        for (int l = 0; l < 3; l++) CK(cudaEventCreateWithFlags(&c.carry[l], cudaEventDisableTiming));
        CK(cudaStreamCreateWithFlags(&c.up, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c.down, cudaStreamNonBlocking));
    }
    *out = &c;
    return CCT_OK;
}

// Pass schedule of the host entry points: about 512 MiB of raw pixels per pass, so that the copies of one pass hide
// behind the kernels of its neighbours while every kernel still sees enough streams to fill the GPU.  The batch opens
// with shorter passes, because the upload (encode) / the kernels (decode) of the first pass are the part nothing can
// hide: a quarter and a half pass for encode, whose bound is the kernels; an eighth, a quarter and a half for decode,
// whose bound is the download of 2 bytes per pixel -- it has to start as early as possible and then never wait (the
// first pixels cannot come before ~8 ms: one stream alone takes that long through the bit-serial inflate).
// Encode also closes with a half and a quarter pass: the download of the last pass's bytes is the tail of the call.
// CCT_HOST_PASS overrides the pass size (0 = one pass), CCT_HOST_LEADS the opening ("8,4,2" = those fractions),
// CCT_HOST_TAILS the closing passes.
// Returns the largest pass.
int host_schedule(int B, size_t n_px, std::vector<int> &starts, bool decode) {
    int per = (int)std::max<size_t>(1, ((size_t)512 << 20) / (n_px * 2));
    if (const char *e = getenv("CCT_HOST_PASS")) { const int v = atoi(e); per = v > 0 ? v : B; }
    per = std::max(1, std::min(std::min(per, B), kMaxPass));
    std::vector<int> div = decode ? std::vector<int>{8, 4, 2} : std::vector<int>{4, 2};
    if (const char *e = getenv("CCT_HOST_LEADS")) {
        div.clear();
        for (const char *q = e; *q;) { const int v = atoi(q); if (v > 0) div.push_back(v); while (*q && *q != ',') q++; if (*q) q++; }
    }
    // closing passes (encode): the download of the last pass is the part nothing can hide either
    std::vector<int> tails = decode ? std::vector<int>{} : std::vector<int>{2, 4};
    if (const char *e = getenv("CCT_HOST_TAILS")) {
        tails.clear();
        for (const char *q = e; *q;) { const int v = atoi(q); if (v > 0) tails.push_back(v); while (*q && *q != ',') q++; if (*q) q++; }
    }
    starts.clear();
    starts.push_back(0);
    if (const char *e = getenv(decode ? "CCT_HOST_SIZES_DEC" : "CCT_HOST_SIZES")) {  // explicit pass sizes (experiments): used when they add up to B
        int at = 0, largest = 0;
        for (const char *q = e; *q;) {
            const int v = atoi(q);
            if (v > 0 && v <= kMaxPass) { at += v; largest = std::max(largest, v); starts.push_back(at); }
            while (*q && *q != ',') q++;
            if (*q) q++;
        }
        if (at == B) return largest;
        starts.resize(1);
    }
    if (B > per) {
        int at = 0, largest = 0;
        for (int d : div) {
            const int lead = std::max(1, per / d);
            if (at + lead >= B) break;
            at += lead; largest = std::max(largest, lead);
            starts.push_back(at);
        }
        std::vector<int> closing;  // sizes of the closing passes, in order
        int tail_sum = 0;
        for (int d : tails) {
            const int t = std::max(1, per / d);
            if (at + tail_sum + t + per > B) break;  // keep at least one full pass in the middle
            closing.push_back(t); tail_sum += t; largest = std::max(largest, t);
        }
        const int rest = B - at - tail_sum, n = (rest + per - 1) / per, each = (rest + n - 1) / n;  // equal passes in between
        for (int k = 1; k < n; k++) starts.push_back(at + k * each);
        int pos = at + rest;
        for (int t : closing) { starts.push_back(pos); pos += t; }
        per = std::max(largest, each);
    }
    starts.push_back(B);
    return per;
}
// Neighbouring passes run on two streams (CCT_HOST_LANES = 1, 2 or 3).
int host_lanes(int n_pass) {
    int n = 2;
    if (const char *e = getenv("CCT_HOST_LANES")) n = std::min(std::max(atoi(e), 1), 3);
    return std::min(n, std::max(n_pass, 1));
}

}  // namespace

extern "C" {

int cct_encode_host(const uint16_t *h_px, int B, int rows, int cols, uint32_t flags, uint8_t *h_out, size_t out_cap,
                    int64_t *h_offsets, uint32_t *h_status, int32_t *h_stats, int device) {
    int rc = check_shape(B, rows, cols);
    if (rc != CCT_OK) return rc;
    if (!h_offsets || (B > 0 && (!h_px || !h_out || !h_status))) return CCT_E_ARG;
    DeviceGuard restore;
    HostCtx *c = nullptr;
    rc = get_ctx(device, &c);
    if (rc != CCT_OK) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    h_offsets[0] = 0;
    if (B == 0) return CCT_OK;
    const size_t n_px = (size_t)rows * cols;
    std::vector<int> starts;
    const int per = host_schedule(B, n_px, starts, false), n_pass = (int)starts.size() - 1;
    const size_t arena = std::min(out_cap, (size_t)B * cct_file_bound(rows, cols));
    HostPin pin_px(h_px, (size_t)B * n_px * 2), pin_out(h_out, arena), pin_off(h_offsets, (size_t)(B + 1) * 8);
    const size_t ws_bytes = cct_encode_workspace_bytes(per, rows, cols, flags);
    if ((rc = c->ensure(0, (size_t)B * n_px * 2))) return rc;
    if ((rc = c->ensure(1, std::max<size_t>(arena, 16)))) return rc;
    if ((rc = c->ensure(2, (size_t)(B + 1) * 8 + (size_t)B * 4 + (size_t)B * 16))) return rc;
    if ((rc = c->ensure(3, ws_bytes))) return rc;
    const int n_lanes = host_lanes(n_pass);
    for (int l = 1; l < n_lanes; l++) if ((rc = c->ensure(4 + l, ws_bytes))) return rc;
    uint16_t *d_px = (uint16_t *)c->buf[0];
    uint8_t *d_out = (uint8_t *)c->buf[1];
    int64_t *d_off = (int64_t *)c->buf[2];
    uint32_t *d_status = (uint32_t *)(d_off + B + 1);
    int32_t *d_stats = (int32_t *)(d_status + B);
    // Three streams: pixels of every pass go up on `up` (queued at once), the encoder's passes run on `st`, each
    // waiting only for its own pixels, and the bytes of pass k come down on `down` while pass k+1 is running.
    c->trace_begin(c->up);
    for (int k = 0; k < n_pass; k++) {
        const size_t lo = (size_t)starts[k], cnt = (size_t)(starts[k + 1] - starts[k]);
        CK(cudaMemcpyAsync(d_px + lo * n_px, h_px + lo * n_px, cnt * n_px * 2, cudaMemcpyHostToDevice, c->up));
        CK(cudaEventRecord(c->event(k), c->up));
        c->mark("up", k, c->up);
    }
    bool over = false;
    auto fetch = [&](int k) -> int {  // bytes of pass k -> host (its offsets have been copied on `st` before event n_pass + k)
        const int lo = starts[k], hi = starts[k + 1];
        CK(cudaEventSynchronize(c->event(n_pass + k)));
        const int64_t b0 = h_offsets[lo], b1 = h_offsets[hi];
        if (b1 > (int64_t)out_cap || b1 > (int64_t)arena) { over = true; return CCT_OK; }
        if (b1 > b0) CK(cudaMemcpyAsync(h_out + b0, d_out + b0, (size_t)(b1 - b0), cudaMemcpyDeviceToHost, c->down));
        c->mark("down", k, c->down);
        return CCT_OK;
    };
    PassHooks hooks;
    hooks.max_pass = per;
    hooks.starts = &starts;
    if (n_lanes > 1) c->more_lanes(hooks, n_lanes);
    const int back = std::max(n_lanes, 2);  // how far the downloads trail the pass being queued
    hooks.before = [&](int k, int, int, cudaStream_t ps) -> int {
        CK(cudaStreamWaitEvent(ps, c->event(k), 0));
        c->mark("start", k, ps);
        return CCT_OK;
    };
    hooks.after = [&](int k, int first, int count, cudaStream_t ps) -> int {
        c->mark("end", k, ps);
        CK(cudaMemcpyAsync(h_offsets + first + 1, d_off + first + 1, (size_t)count * 8, cudaMemcpyDeviceToHost, ps));
        CK(cudaEventRecord(c->event(n_pass + k), ps));
        // Passes up to k are queued: now wait for pass k - back and start its download.  (One pass further back than
        // the data needs: measured, the first kernel of whatever is queued AFTER a download has been issued does not
        // start before that download is through -- with k-1 here and two lanes, every pass began 2.6 ms after the one
        // before it on its lane had ended.  With `back` = the number of lanes the next pass is queued while the one
        // before it on its lane is still running.)
        return k >= back ? fetch(k - back) : CCT_OK;
    };
    rc = encode_impl(d_px, B, rows, cols, flags, d_out, arena, d_off, d_status, d_stats, c->buf[3], c->cap[3], c->st, &hooks);
    for (int k = std::max(0, n_pass - back); rc == CCT_OK && k < n_pass; k++) rc = fetch(k);
    if (rc != CCT_OK) { c->drain(); return rc; }
    CK(cudaMemcpyAsync(h_status, d_status, (size_t)B * 4, cudaMemcpyDeviceToHost, c->st));
    if (h_stats) CK(cudaMemcpyAsync(h_stats, d_stats, (size_t)B * 16, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    CK(cudaStreamSynchronize(c->down));
    c->trace_end("encode");
    return over ? CCT_E_CAPACITY : CCT_OK;
}

int cct_decode_host(const uint8_t *h_in, const int64_t *h_offsets, int B, int rows, int cols, uint16_t *h_px,
                    uint32_t *h_status, int device) {
    int rc = check_shape(B, rows, cols);
    if (rc != CCT_OK) return rc;
    if (B > 0 && (!h_in || !h_offsets || !h_px || !h_status)) return CCT_E_ARG;
    DeviceGuard restore;
    HostCtx *c = nullptr;
    rc = get_ctx(device, &c);
    if (rc != CCT_OK) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    if (B == 0) return CCT_OK;
    const size_t n_px = (size_t)rows * cols;
    const int64_t base = h_offsets[0], total = h_offsets[B] - base;
    if (total < 0) return CCT_E_ARG;
    HostPin pin_in(h_in + base, (size_t)total), pin_px(h_px, (size_t)B * n_px * 2);
    for (int i = 0; i < B; i++) if (h_offsets[i + 1] < h_offsets[i]) return CCT_E_ARG;
    std::vector<int> starts;
    const int per = host_schedule(B, n_px, starts, true), n_pass = (int)starts.size() - 1;
    const size_t ws_bytes = cct_decode_workspace_bytes(per, rows, cols);
    if ((rc = c->ensure(0, (size_t)B * n_px * 2))) return rc;
    if ((rc = c->ensure(1, std::max<size_t>((size_t)total, 16)))) return rc;
    if ((rc = c->ensure(2, (size_t)(B + 1) * 8 + (size_t)B * 4))) return rc;
    if ((rc = c->ensure(4, ws_bytes))) return rc;
    const int n_lanes = host_lanes(n_pass);
    for (int l = 1; l < n_lanes; l++) if ((rc = c->ensure(4 + l, ws_bytes))) return rc;
    uint16_t *d_px = (uint16_t *)c->buf[0];
    uint8_t *d_in = (uint8_t *)c->buf[1];
    int64_t *d_off = (int64_t *)c->buf[2];
    uint32_t *d_status = (uint32_t *)(d_off + B + 1);
    std::vector<int64_t> rel((size_t)B + 1);
    for (int i = 0; i <= B; i++) rel[i] = h_offsets[i] - base;
    c->trace_begin(c->up);
    CK(cudaMemcpyAsync(d_off, rel.data(), (size_t)(B + 1) * 8, cudaMemcpyHostToDevice, c->up));
    // files of every pass go up on `up` (queued at once); pass k decodes on `st` as soon as its files are there;
    // its pixels come down on `down` while pass k+1 decodes
    for (int k = 0; k < n_pass; k++) {
        const int lo = starts[k], hi = starts[k + 1];
        if (rel[hi] > rel[lo])
            CK(cudaMemcpyAsync(d_in + rel[lo], h_in + base + rel[lo], (size_t)(rel[hi] - rel[lo]), cudaMemcpyHostToDevice, c->up));
        CK(cudaEventRecord(c->event(k), c->up));
        c->mark("up", k, c->up);
    }
    PassHooks hooks;
    hooks.max_pass = per;
    hooks.starts = &starts;
    if (n_lanes > 1) c->more_lanes(hooks, n_lanes);
    hooks.before = [&](int k, int, int, cudaStream_t ps) -> int {
        CK(cudaStreamWaitEvent(ps, c->event(k), 0));
        c->mark("start", k, ps);
        return CCT_OK;
    };
    hooks.after = [&](int k, int first, int count, cudaStream_t ps) -> int {
        c->mark("end", k, ps);
        cudaEvent_t e = c->event(n_pass + k);
        CK(cudaEventRecord(e, ps));
        CK(cudaStreamWaitEvent(c->down, e, 0));
        CK(cudaMemcpyAsync(h_px + (size_t)first * n_px, d_px + (size_t)first * n_px, (size_t)count * n_px * 2,
                           cudaMemcpyDeviceToHost, c->down));
        c->mark("down", k, c->down);
        return CCT_OK;
    };
    rc = decode_impl(d_in, d_off, B, rows, cols, d_px, d_status, c->buf[4], c->cap[4], c->st, &hooks);
    if (rc != CCT_OK) { c->drain(); return rc; }
    CK(cudaMemcpyAsync(h_status, d_status, (size_t)B * 4, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->up));  // `rel` must outlive its copy
    CK(cudaStreamSynchronize(c->st));
    CK(cudaStreamSynchronize(c->down));
    c->trace_end("decode");
    return CCT_OK;
}

int cct_debug_host_schedule(int B, size_t n_px, int decode, int32_t *starts, int cap) {
    if (B < 0 || n_px == 0 || !starts) return CCT_E_ARG;
    std::vector<int> st;
    if (B == 0) { if (cap < 1) return CCT_E_ARG; starts[0] = 0; return 0; }
    host_schedule(B, n_px, st, decode != 0);
    if ((int)st.size() > cap) return CCT_E_ARG;
    for (size_t k = 0; k < st.size(); k++) starts[k] = st[k];
    return (int)st.size() - 1;
}

void cct_release(void) {
    {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        for (auto &kv : g_ctx) {
            cudaSetDevice(kv.first);
            for (int i = 0; i < 7; i++) if (kv.second.buf[i]) cudaFree(kv.second.buf[i]);
            for (cudaEvent_t e : kv.second.events) cudaEventDestroy(e);
            if (kv.second.st) cudaStreamDestroy(kv.second.st);
            if (kv.second.st2) cudaStreamDestroy(kv.second.st2);
            if (kv.second.st3) cudaStreamDestroy(kv.second.st3);
            if (kv.second.ready) cudaEventDestroy(kv.second.ready);
            for (cudaEvent_t e : kv.second.carry) if (e) cudaEventDestroy(e);
            if (kv.second.up) cudaStreamDestroy(kv.second.up);
            if (kv.second.down) cudaStreamDestroy(kv.second.down);
        }
        g_ctx.clear();
    }
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto &kv : g_orders) { cudaSetDevice(std::get<0>(kv.first)); cudaFree(kv.second); }
    g_orders.clear();
    for (auto &kv : g_side) { cudaSetDevice(kv.first); if (kv.second.stream) cudaStreamDestroy(kv.second.stream); if (kv.second.ready) cudaEventDestroy(kv.second.ready); }
    g_side.clear();
}

}  // extern "C"
