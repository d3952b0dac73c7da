// C-ABI glue: error reporting, the correlation dispatcher, the workspace-owning context
// and the batched scene solver that replaces ImageCutSolver._cut_and_pool / _solver /
// _execute_matching (misc/image_cut_solver.py:95-184).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "dm_common.cuh"
#include "dm_internal.h"

// ------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

void dm_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* dm_last_error(void) { return g_err; }

// programmatic dependent launch between the kernels of a chunk (dm_common.cuh); DM_PDL = mask of DM_PDL_*
#define DM_PDL_DEFAULT DM_PDL_CORR
static thread_local bool g_pdl_suppressed = false;
bool dm_pdl_enabled(int which) {
    static const int mask = getenv("DM_PDL") ? atoi(getenv("DM_PDL")) : DM_PDL_DEFAULT;
    return (mask & which) != 0 && !g_pdl_suppressed;
}
void dm_pdl_suppress(bool off) { g_pdl_suppressed = off; }
extern "C" int dm_version(void) { return 100; }

extern "C" int dm_device_cc(void) {
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return DM_ERR_CUDA;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return DM_ERR_CUDA;
    if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return DM_ERR_CUDA;
    return major * 10 + minor;
}

extern "C" int dm_correlation_set_pair_mode(int mode) {
    DM_REQUIRE(mode >= -1 && mode <= 1, DM_ERR_INVALID, "dm_correlation_set_pair_mode: mode %d", mode);
    dm_correlation_umma_set_pair_mode(mode);
    return DM_OK;
}

// ------------------------------------------------------------------ correlation dispatch
extern "C" int dm_correlation(const void* desc1_dev, const float* stat1_dev,
                              const void* desc2_dev, const float* stat2_dev,
                              int n_tiles, int p, int kpad, int ws, int method, int engine,
                              float* raw_dev, void* stream) {
    DM_REQUIRE(method == DM_TM_CCOEFF || method == DM_TM_CCOEFF_NORMED, DM_ERR_INVALID, "dm_correlation: invalid method %d", method);
    DM_REQUIRE(n_tiles > 0 && p > 0 && kpad == dm_kpad(ws), DM_ERR_INVALID, "dm_correlation: bad shape (kpad %d for ws %d)", kpad, ws);
    cudaStream_t st = (cudaStream_t)stream;
    const bool umma_ok = dm_correlation_umma_supported(p, kpad);
    const int kreal = dm_desc_kreal(ws);
    if (engine == 3) {      // undocumented measurement aid: MMA + TMEM drain only
        DM_REQUIRE(umma_ok, DM_ERR_UNSUPPORTED, "dm_correlation: unsupported shape");
        return dm_correlation_umma_null(desc1_dev, stat1_dev, desc2_dev, stat2_dev, n_tiles, p, kpad, kreal, raw_dev, st);
    }
    if (engine == 4) {      // undocumented test aid: pooled epilogue; raw_dev = [pooled n*P*P/4][rowmin n*P*4][rowmax n*P*4], square grid
        int t = 1; while (t * t < p) ++t;
        DM_REQUIRE(t * t == p && dm_correlation_umma_pool_supported(t, t, kpad), DM_ERR_UNSUPPORTED, "dm_correlation: unsupported shape for the pooled test engine");
        float* rmin = raw_dev + (size_t)n_tiles * p * (p / 4);
        return dm_correlation_umma_pool(desc1_dev, stat1_dev, desc2_dev, stat2_dev, n_tiles, t, t, kpad, kreal, method, 0, raw_dev, rmin, rmin + (size_t)n_tiles * p * 4, st);
    }
    if (engine == DM_CORR_UMMA) {
        DM_REQUIRE(umma_ok, DM_ERR_UNSUPPORTED, "dm_correlation: tcgen05 engine needs P %% 128 == 0 and kpad <= 256 (P=%d kpad=%d)", p, kpad);
        return dm_correlation_umma(desc1_dev, stat1_dev, desc2_dev, stat2_dev, n_tiles, p, kpad, kreal, method, raw_dev, st);
    }
    if (engine == DM_CORR_AUTO && umma_ok)
        return dm_correlation_umma(desc1_dev, stat1_dev, desc2_dev, stat2_dev, n_tiles, p, kpad, kreal, method, raw_dev, st);
    return dm_correlation_simt(desc1_dev, stat1_dev, desc2_dev, stat2_dev, n_tiles, p, kpad, ws, method, raw_dev, st);
}

// ------------------------------------------------------------------ context
extern "C" int dm_ctx_create(dm_ctx** out) {
    DM_REQUIRE(out != nullptr, DM_ERR_INVALID, "dm_ctx_create: null out");
    int cc = dm_device_cc();
    DM_REQUIRE(cc >= 100, DM_ERR_UNSUPPORTED, "dm_ctx_create: this library is built for sm_100a only (device cc %d)", cc);
    *out = new dm_ctx();
    cudaGetDevice(&(*out)->device);
    return DM_OK;
}

extern "C" void dm_ctx_destroy(dm_ctx* ctx) {
    if (!ctx) return;
    cudaFree(ctx->ws); cudaFree(ctx->scene1); cudaFree(ctx->scene2); cudaFree(ctx->planes);
    for (auto& v : ctx->ev) for (auto e : v) cudaEventDestroy(e);
    for (auto e : ctx->band_ev) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    for (int k = 0; k < 2; ++k) if (ctx->sp_stream[k]) cudaStreamDestroy(ctx->sp_stream[k]);
    cudaFree(ctx->sp_buf);
    if (ctx->up_ev) cudaEventDestroy(ctx->up_ev);
    for (auto& g : ctx->upper_graphs) cudaGraphExecDestroy(g.exec);
    if (ctx->capture_stream) cudaStreamDestroy(ctx->capture_stream);
    delete ctx;
}

extern "C" int dm_ctx_set_stream(dm_ctx* ctx, void* stream) { ctx->stream = (cudaStream_t)stream; return DM_OK; }
extern "C" int dm_ctx_set_workspace_limit(dm_ctx* ctx, size_t bytes) { ctx->ws_limit = bytes; return DM_OK; }
extern "C" size_t dm_ctx_workspace_bytes(const dm_ctx* ctx) { return ctx->ws_bytes; }
extern "C" int dm_ctx_enable_timing(dm_ctx* ctx, int on) { ctx->timing = on != 0; return DM_OK; }

static int ctx_reserve(dm_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->ws_bytes) return DM_OK;
    DM_REQUIRE(bytes <= ctx->ws_limit, DM_ERR_NOMEM, "workspace of %zu bytes exceeds the limit of %zu", bytes, ctx->ws_limit);
    DM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    // the cached graphs of the upper-pyramid launches bake in pointers into the old workspace
    for (auto& g : ctx->upper_graphs) cudaGraphExecDestroy(g.exec);
    ctx->upper_graphs.clear();
    if (ctx->ws) DM_CUDA_CHECK(cudaFree(ctx->ws));
    ctx->ws = nullptr; ctx->ws_bytes = 0;
    DM_CUDA_CHECK(cudaMalloc(&ctx->ws, bytes));
    ctx->ws_bytes = bytes;
    return DM_OK;
}

// The pixels a contiguous range of tiles [a,b) (row-major global index) OWNS -- the reference pastes
// tiles j-outer / i-inner, so a pixel belongs to the covering tile with the largest index
// (misc/image_cut_solver.py:165-175): tile row g owns output rows [s0 g, s0 (g+1)) (the last tile row
// up to out_h), tile column c owns columns [s1 c, s1 (c+1)) (the last one up to out_w).  At most three
// rectangles: the tail of the first tile row, whole tile rows, the head of the last tile row.
struct OwnedRect { int r0, r1, c0, c1; };
static int owned_rects(int len0, int len1, int s0, int s1, int out_h, int out_w, long long a, long long b, OwnedRect out[3]) {
    if (b <= a) return 0;
    auto row_lo = [&](long long g) { return (int)(s0 * g); };
    auto row_hi = [&](long long g) { return g >= len0 - 1 ? out_h : (int)(s0 * (g + 1)); };
    auto col_lo = [&](long long c) { return (int)(s1 * c); };
    auto col_hi = [&](long long c) { return c >= len1 - 1 ? out_w : (int)(s1 * (c + 1)); };
    const long long ga = a / len1, ca = a % len1, gb = (b - 1) / len1, cb = (b - 1) % len1;
    int n = 0;
    if (ga == gb) { out[n++] = {row_lo(ga), row_hi(ga), col_lo(ca), col_hi(cb)}; return n; }
    long long mid_a = ga, mid_b = gb;               // whole tile rows [mid_a, mid_b]
    if (ca != 0) { out[n++] = {row_lo(ga), row_hi(ga), col_lo(ca), out_w}; mid_a = ga + 1; }
    const bool tail = cb != len1 - 1;
    if (tail) mid_b = gb - 1;
    if (mid_b >= mid_a) out[n++] = {row_lo(mid_a), row_hi(mid_b), 0, out_w};
    if (tail) out[n++] = {row_lo(gb), row_hi(gb), 0, col_hi(cb)};
    return n;
}

// Copies what the tiles [rb.tiles_copied, tiles_done) own to the stream destination, on the copy stream
// and behind an event: a tile's pixels are written by that tile alone, so they are final as soon as
// the tile is.  The destination is page-locked host memory or the mosaic of a peer device (UVA decides).
int dm_readback_rows(dm_ctx* ctx, long long tiles_done) {
    dm_ctx::Readback& rb = ctx->rb;
    if (!rb.active || tiles_done <= rb.tiles_copied) return DM_OK;
    if (!ctx->copy_stream) DM_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (ctx->band_used >= ctx->band_ev.size()) {
        cudaEvent_t e;
        DM_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->band_ev.push_back(e);
    }
    cudaEvent_t ev = ctx->band_ev[ctx->band_used++];
    DM_CUDA_CHECK(cudaEventRecord(ev, ctx->stream));
    DM_CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ev, 0));
    const size_t plane = (size_t)rb.out_h * rb.out_w, pitch = (size_t)rb.out_w * sizeof(double);
    if (rb.n_scenes > 1) {
        // a batch of scenes ([scene][mode][plane] and [scene][plane]) leaves scene by scene: whole
        // scenes are two contiguous blocks each, no rectangles
        const long long tps = (long long)rb.len0 * rb.len1;
        const long long s_a = rb.tiles_copied / tps, s_b = tiles_done / tps;
        if (s_b > s_a) {
            const size_t cnt = (size_t)(s_b - s_a);
            DM_CUDA_CHECK(cudaMemcpyAsync(rb.dst_d_map + (size_t)s_a * rb.n_modes * plane, rb.d_d_map + (size_t)s_a * rb.n_modes * plane,
                                          cnt * rb.n_modes * plane * sizeof(double), cudaMemcpyDefault, ctx->copy_stream));
            DM_CUDA_CHECK(cudaMemcpyAsync(rb.dst_out_map + (size_t)s_a * plane, rb.d_out_map + (size_t)s_a * plane,
                                          cnt * plane * sizeof(double), cudaMemcpyDefault, ctx->copy_stream));
            rb.tiles_copied = s_b * tps;
        }
        return DM_OK;
    }
    OwnedRect rc[3];
    const int n = owned_rects(rb.len0, rb.len1, rb.s0, rb.s1, rb.out_h, rb.out_w, rb.tiles_copied, tiles_done, rc);
    for (int k = 0; k < n; ++k) {
        const size_t off = (size_t)rc[k].r0 * rb.out_w + rc[k].c0;
        const size_t width = (size_t)(rc[k].c1 - rc[k].c0) * sizeof(double), rows = (size_t)(rc[k].r1 - rc[k].r0);
        for (int m = 0; m <= rb.n_modes; ++m) {
            double* dst = (m < rb.n_modes ? rb.dst_d_map + m * plane : rb.dst_out_map) + off;
            const double* src = (m < rb.n_modes ? rb.d_d_map + m * plane : rb.d_out_map) + off;
            if (width == pitch) DM_CUDA_CHECK(cudaMemcpyAsync(dst, src, rows * pitch, cudaMemcpyDefault, ctx->copy_stream));
            else DM_CUDA_CHECK(cudaMemcpy2DAsync(dst, pitch, src, pitch, width, rows, cudaMemcpyDefault, ctx->copy_stream));
        }
    }
    rb.tiles_copied = tiles_done;
    return DM_OK;
}

// Host scenes go to the device in pieces, each piece ahead of the chunk of tiles that reads it
// (misc/raw_read.py:36-45 hands ImageCutSolver two host arrays).
// dm_upload_rows enqueues the copy of the stacked scene rows [rows_done, upto) of both images on the
// upload stream; dm_upload_wait makes the compute stream wait for what has been enqueued so far.
static int dm_upload_rows(dm_ctx* ctx, long long upto) {
    dm_ctx::Upload& up = ctx->up;
    if (!up.active) return DM_OK;
    if (upto > up.row_end) upto = up.row_end;
    if (upto <= up.rows_done) return DM_OK;
    if (!ctx->h2d_stream) DM_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    if (!ctx->up_ev) DM_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->up_ev, cudaEventDisableTiming));
    const size_t off = (size_t)up.rows_done * up.row_bytes, bytes = (size_t)(upto - up.rows_done) * up.row_bytes;
    DM_CUDA_CHECK(cudaMemcpyAsync(ctx->scene1 + off, up.h1 + off, bytes, cudaMemcpyHostToDevice, ctx->h2d_stream));
    DM_CUDA_CHECK(cudaMemcpyAsync(ctx->scene2 + off, up.h2 + off, bytes, cudaMemcpyHostToDevice, ctx->h2d_stream));
    DM_CUDA_CHECK(cudaEventRecord(ctx->up_ev, ctx->h2d_stream));
    up.rows_done = upto;
    up.pending = true;
    return DM_OK;
}
static int dm_upload_wait(dm_ctx* ctx) {
    if (!ctx->up.active || !ctx->up.pending) return DM_OK;
    DM_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->up_ev, 0));
    ctx->up.pending = false;
    return DM_OK;
}

int StageTimer::begin(size_t chunk) {
    if (!ctx->timing) return DM_OK;
    auto& v = ctx->ev[stage];
    slot = chunk * 2;
    while (v.size() < slot + 2) { cudaEvent_t e; DM_CUDA_CHECK(cudaEventCreate(&e)); v.push_back(e); }
    DM_CUDA_CHECK(cudaEventRecord(v[slot], ctx->stream));
    return DM_OK;
}
int StageTimer::end() {
    if (!ctx->timing) return DM_OK;
    DM_CUDA_CHECK(cudaEventRecord(ctx->ev[stage][slot + 1], ctx->stream));
    return DM_OK;
}

extern "C" int dm_ctx_stage_ms(dm_ctx* ctx, float* ms_out, int* launches_out) {
    DM_REQUIRE(ctx->timing, DM_ERR_INVALID, "dm_ctx_stage_ms: timing not enabled");
    DM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (int s = 0; s < DM_STAGE_COUNT; ++s) {
        float total = 0.f;
        for (size_t c = 0; c < ctx->chunks_last && 2 * c + 1 < ctx->ev[s].size(); ++c) {
            float ms = 0.f;
            DM_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev[s][2 * c], ctx->ev[s][2 * c + 1]));
            total += ms;
        }
        if (ms_out) ms_out[s] = total;
        if (launches_out) launches_out[s] = ctx->launches[s];
    }
    return DM_OK;
}

// ------------------------------------------------------------------ geometry
static int ilog2_exact(int v) { int l = 0; while ((1 << l) < v) ++l; return ((1 << l) == v) ? l : -1; }

// the tiles [a,b) (row-major index of one scene) a solve covers: a tile range, a strip of tile rows, or all
int dm_tile_range(const dm_scene_params* prm, int len0, int len1, long long* a, long long* b) {
    const long long all = (long long)len0 * len1;
    if (prm->tile_hi > 0) {
        DM_REQUIRE(prm->tile_row_hi <= 0, DM_ERR_INVALID, "tile_lo/tile_hi and tile_row_lo/tile_row_hi exclude each other");
        DM_REQUIRE(prm->tile_lo >= 0 && prm->tile_lo < prm->tile_hi && prm->tile_hi <= all, DM_ERR_INVALID,
                   "tile range [%d,%d) outside [0,%lld)", prm->tile_lo, prm->tile_hi, all);
        *a = prm->tile_lo; *b = prm->tile_hi;
        return DM_OK;
    }
    int lo = prm->tile_row_lo, hi = prm->tile_row_hi;
    if (hi <= 0) { lo = 0; hi = len0; }
    DM_REQUIRE(lo >= 0 && lo < hi && hi <= len0, DM_ERR_INVALID, "tile row strip [%d,%d) outside [0,%d)", lo, hi, len0);
    *a = (long long)lo * len1; *b = (long long)hi * len1;
    return DM_OK;
}

extern "C" int dm_scene_geometry(const dm_scene_params* prm, dm_scene_info* info) {
    DM_REQUIRE(prm && info, DM_ERR_INVALID, "dm_scene_geometry: null argument");
    DM_REQUIRE(prm->ws >= 1 && (prm->ws & 1) && prm->ws <= 31, DM_ERR_INVALID, "window_size must be odd and <= 31 (got %d)", prm->ws);
    DM_REQUIRE(prm->t0 > 0 && prm->t1 > 0 && prm->s0 > 0 && prm->s1 > 0, DM_ERR_INVALID, "image_size and stride must be positive");
    const int mn = prm->t0 < prm->t1 ? prm->t0 : prm->t1, mxs = prm->t0 < prm->t1 ? prm->t1 : prm->t0;
    const int lg = ilog2_exact(mn);
    // misc/Correlation_map.py:192 -- the patch grid's short side must be a power of two and
    // the long side a multiple of it; the reference fails with a
    // broadcast error otherwise.
    DM_REQUIRE(lg >= 0 && mxs % mn == 0, DM_ERR_INVALID, "image_size (%d,%d): short side must be a power of two dividing the long side", prm->t0, prm->t1);
    DM_REQUIRE(prm->method == DM_TM_CCOEFF || prm->method == DM_TM_CCOEFF_NORMED, DM_ERR_INVALID, "invalid feature method %d", prm->method);
    DM_REQUIRE(prm->n_modes >= 1 && prm->n_modes <= 4, DM_ERR_INVALID, "1..4 disparity modes supported (got %d)", prm->n_modes);
    for (int m = 0; m < prm->n_modes; ++m)
        DM_REQUIRE(prm->modes[m] >= 0 && prm->modes[m] <= 2, DM_ERR_INVALID, "invalid disparity mode %d", prm->modes[m]);
    memset(info, 0, sizeof(*info));
    const int e2 = prm->ws - 1;
    // misc/image_cut_solver.py:62: floor((S - trimmed) / stride) -- the last fitting tile is dropped
    const int d0 = prm->scene_h - (prm->t0 + e2), d1 = prm->scene_w - (prm->t1 + e2);
    info->len0 = d0 >= 0 ? d0 / prm->s0 : -1;
    info->len1 = d1 >= 0 ? d1 / prm->s1 : -1;
    DM_REQUIRE(info->len0 >= 1 && info->len1 >= 1, DM_ERR_INVALID, "scene %dx%d yields no tiles (len = %d,%d)", prm->scene_h, prm->scene_w, info->len0, info->len1);
    info->out_h = prm->s0 * (info->len0 - 1) + prm->t0;
    info->out_w = prm->s1 * (info->len1 - 1) + prm->t1;
    const int ns = prm->n_scenes > 1 ? prm->n_scenes : 1;
    DM_REQUIRE((long long)ns * info->len0 * info->len1 < (1LL << 31), DM_ERR_INVALID, "too many tiles");
    long long ta, tb;
    int rc = dm_tile_range(prm, info->len0, info->len1, &ta, &tb);
    if (rc != DM_OK) return rc;
    const int ga = (int)(ta / info->len1), gb = (int)((tb - 1) / info->len1);
    info->row_lo = prm->s0 * ga;
    info->row_hi = (gb == info->len0 - 1) ? info->out_h : prm->s0 * (gb + 1);
    DM_REQUIRE(ns == 1 || (ta == 0 && tb == (long long)info->len0 * info->len1), DM_ERR_INVALID, "a batch of scenes cannot be combined with a strip or a tile range");
    info->n_tiles = (int)(tb - ta) * ns;
    info->levels = lg + 1;
    info->n_map = mn;
    return DM_OK;
}

// ------------------------------------------------------------------ scene kernels
namespace {

// tile g of a (batch of) scene(s): scene = g / (len0*len1), then row-major (gi, gj); scenes are
// stacked along rows, so the origin of a tile is (scene*S0 + s0*gi, s1*gj)
__global__ void dm_tile_origin_kernel(int32_t* origin, int n, int first_tile, int len0, int len1, int s0, int s1, int scene_h) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int g = first_tile + t, tps = len0 * len1;
    const int sc = g / tps, r = g - sc * tps;
    origin[2 * t] = sc * scene_h + s0 * (r / len1);
    origin[2 * t + 1] = s1 * (r % len1);
}

struct PlaneArgs {
    int n_modes; int modes[4];
    int T0, T1, s0, s1, len0, len1, out_h, out_w;
    int first_tile; int sub_pix;
};

// misc/Matching.py:165-209 + misc/Calc_difference.py:25-49 + the paste of
// misc/image_cut_solver.py:165-175.  Overlaps: the reference pastes tiles j-outer/i-inner,
// so a pixel ends up owned by the covering tile with the largest (i,j); this kernel writes
// a pixel only from its owner, which makes the scatter race-free and order-free.
__global__ void __launch_bounds__(256)
dm_planes_kernel(const float* __restrict__ l0, long long total, PlaneArgs a,
                 const int32_t* __restrict__ match, const float* __restrict__ score,
                 double* __restrict__ d_map, double* __restrict__ out_map) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int P = a.T0 * a.T1;
    const long long n = idx / P;
    const int p = (int)(idx - n * P);
    const int i = p / a.T1, j = p - i * a.T1;
    const int g = a.first_tile + (int)n, tps = a.len0 * a.len1;
    const int sc = g / tps, gr = g - sc * tps;
    const int gi = gr / a.len1, gj = gr - gi * a.len1;
    const int Y = a.s0 * gi + i, X = a.s1 * gj + j;
    if (min(Y / a.s0, a.len0 - 1) != gi || min(X / a.s1, a.len1 - 1) != gj) return;
    const int c0 = match[(size_t)n * 2 * P + p];
    const int c1 = match[(size_t)n * 2 * P + P + p];
    double m0 = (double)c0, m1 = (double)c1;
    // numpy index rules as in dm_match_map_kernel (backtrack.cu): negative indices wrap, an
    // index >= size is the reference's swallowed IndexError (only reachable after a level-0 filter)
    if (a.sub_pix && dm_np_index_ok(c0, a.T0) && dm_np_index_ok(c1, a.T1)) {
        const float* map = l0 + ((size_t)n * P + p) * (size_t)P;
        const int w0 = dm_np_wrap(c0, a.T0), w1 = dm_np_wrap(c1, a.T1);
        const float r0 = map[(size_t)w0 * a.T1 + w1];
        if (dm_np_index_ok(c0 + 1, a.T0) && dm_np_index_ok(c0 - 1, a.T0)) {
            const float r1 = map[(size_t)dm_np_wrap(c0 + 1, a.T0) * a.T1 + w1];
            const float rm = map[(size_t)dm_np_wrap(c0 - 1, a.T0) * a.T1 + w1];
            if (r0 > r1 && r0 > rm) m0 += dm_parabola_shift(r0, r1, rm);
        }
        if (dm_np_index_ok(c1 + 1, a.T1) && dm_np_index_ok(c1 - 1, a.T1)) {
            const float r1 = map[(size_t)w0 * a.T1 + dm_np_wrap(c1 + 1, a.T1)];
            const float rm = map[(size_t)w0 * a.T1 + dm_np_wrap(c1 - 1, a.T1)];
            if (r0 > r1 && r0 > rm) m1 += dm_parabola_shift(r0, r1, rm);
        }
    }
    const double e0 = __dsub_rn((double)i, m0), e1 = __dsub_rn((double)j, m1);
    const size_t plane = (size_t)a.out_h * a.out_w, pix = (size_t)Y * a.out_w + X;
    for (int m = 0; m < a.n_modes; ++m) {
        double v = a.modes[m] == DM_MODE_ELEVATION ? e1
                 : a.modes[m] == DM_MODE_ELEVATION2 ? e0
                 : __dsqrt_rn(__fma_rn(e1, e1, __dmul_rn(e0, e0)));
        d_map[((size_t)sc * a.n_modes + m) * plane + pix] = v;
    }
    out_map[(size_t)sc * plane + pix] = (double)score[(size_t)n * P + p];
}

}  // namespace

// ------------------------------------------------------------------ materialising solver
struct Carve {
    char* base; size_t off;
    explicit Carve(char* b) : base(b), off(0) {}
    template <typename T> T* take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? (T*)(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

struct TileBuffers {
    int32_t* origin; void* desc1; void* desc2; float* stat1; float* stat2;
    std::vector<float*> level;
    int32_t* match[2]; float* score[2];
};

static size_t carve_tiles(char* base, int nt, int t0, int t1, int kpad, int levels, TileBuffers& tb) {
    Carve c(base);
    const size_t P = (size_t)t0 * t1;
    tb.origin = c.take<int32_t>((size_t)nt * 2);
    tb.desc1 = c.take<__nv_bfloat16>((size_t)nt * P * kpad);
    tb.desc2 = c.take<__nv_bfloat16>((size_t)nt * P * kpad);
    tb.stat1 = c.take<float>((size_t)nt * P * DM_STAT_FLOATS);
    tb.stat2 = c.take<float>((size_t)nt * P * DM_STAT_FLOATS);
    tb.level.clear();
    size_t a = t0, b = t1;
    for (int k = 0; k < levels; ++k) {
        tb.level.push_back(c.take<float>((size_t)nt * a * b * a * b));
        a >>= 1; b >>= 1;
    }
    for (int s = 0; s < 2; ++s) {
        tb.match[s] = c.take<int32_t>((size_t)nt * 2 * P);
        tb.score[s] = c.take<float>((size_t)nt * P);
    }
    return (c.off + 255) & ~(size_t)255;
}

extern "C" int dm_solve_scene(dm_ctx* ctx, const dm_scene_params* prm,
                              const uint8_t* img1_dev, const uint8_t* img2_dev,
                              double* d_map_dev, double* out_map_dev, dm_scene_info* info_out) {
    DM_REQUIRE(ctx && prm && img1_dev && img2_dev && d_map_dev && out_map_dev, DM_ERR_INVALID, "dm_solve_scene: null argument");
    dm_scene_info info;
    int rc = dm_scene_geometry(prm, &info);
    if (rc != DM_OK) return rc;
    cudaStream_t st = ctx->stream;
    const int t0 = prm->t0, t1 = prm->t1, P = t0 * t1, kpad = dm_kpad(prm->ws), L = info.levels;

    const int ns = prm->n_scenes > 1 ? prm->n_scenes : 1;
    const bool want_fused = prm->fused != 0;
    // displacement filter (misc/Matching.py:224-255): the first filter_num maps of the top-down
    // pass; when it reaches level 0 (filter_num >= levels) the materialising path runs, because
    // the fused final kernel produces the planes straight from the level-1 matches
    const int filter_num = prm->filter_num > 0 ? prm->filter_num : 0;
    const int filter_win = prm->filter_cfg & 0xff, filter_mode = (prm->filter_cfg >> 8) & 0xff;
    DM_REQUIRE(filter_num == 0 || (filter_win >= 1 && (filter_mode == DM_FILTER_MEDIAN || filter_mode == DM_FILTER_AVERAGE)),
               DM_ERR_INVALID, "dm_solve_scene: bad filter configuration 0x%x", prm->filter_cfg);
    DM_REQUIRE(filter_num == 0 || t0 == t1 || (t0 < filter_win || t1 < filter_win), DM_ERR_UNSUPPORTED,
               "Matching._filter is undefined on non-square patch grids (%d x %d)", t0, t1);
    const bool fused = want_fused && dm_fused_supported(t0, t1, kpad) && dm_fused_supported_ws(prm->ws) && filter_num < L;
    DM_REQUIRE(!(prm->fused == 1 && !fused), DM_ERR_UNSUPPORTED, "fused path does not support image_size (%d,%d) with window %d", t0, t1, prm->ws);

    // tiles per chunk from the workspace limit
    TileBuffers tb;
    const size_t per_tile = fused ? dm_fused_workspace(nullptr, 1, t0, t1, kpad, L, nullptr)
                                  : carve_tiles(nullptr, 1, t0, t1, kpad, L, tb);
    long long max_chunk = (long long)(ctx->ws_limit / (per_tile + 4096));
    if (max_chunk > 16384) max_chunk = 16384;
    DM_REQUIRE(max_chunk >= 1, DM_ERR_NOMEM, "one tile needs %zu bytes of workspace, limit is %zu", per_tile, ctx->ws_limit);
    // Chunks of tiles.  As few as the workspace limit allows -- except when the finished pixels are streamed
    // somewhere (host arrays, a peer's mosaic): then the strip is cut into a few chunks so that the copy of
    // chunk k rides behind the compute of chunk k+1 instead of behind the last stage alone.  Chunk boundaries
    // are multiples of the number of tiles that fill whole rounds of the persistent correlation grid.
    long long ta, tb_;
    if ((rc = dm_tile_range(prm, info.len0, info.len1, &ta, &tb_)) != DM_OK) return rc;
    const int granule = fused ? dm_correlation_round_tiles(P) : 1;
    int n_chunks = dm_div_up(info.n_tiles, max_chunk);
    if (ctx->rb.active && fused) {
        static const int want = getenv("DM_STREAM_CHUNKS") ? atoi(getenv("DM_STREAM_CHUNKS")) : 3;
        if (want > n_chunks && info.n_tiles / want >= 2 * granule) n_chunks = want;
    }
    int chunk = dm_div_up(info.n_tiles, n_chunks);
    if (chunk >= 2 * granule) {
        // whole rounds per chunk; the last chunk takes the remainder, if the workspace limit lets it
        const int whole = chunk / granule * granule;
        if (info.n_tiles - (long long)(n_chunks - 1) * whole <= max_chunk) chunk = whole;
    }
    auto chunk_begin = [&](int ck) -> int { return ck * chunk; };
    auto chunk_tiles = [&](int ck) -> int { return (ck == n_chunks - 1) ? info.n_tiles - ck * chunk : chunk; };
    int largest = 0;
    for (int ck = 0; ck < n_chunks; ++ck) if (chunk_tiles(ck) > largest) largest = chunk_tiles(ck);
    const size_t need = fused ? dm_fused_workspace(nullptr, largest, t0, t1, kpad, L, nullptr)
                              : carve_tiles(nullptr, largest, t0, t1, kpad, L, tb);
    rc = ctx_reserve(ctx, need);
    if (rc != DM_OK) return rc;
    for (int s = 0; s < DM_STAGE_COUNT; ++s) ctx->launches[s] = 0;
    ctx->chunks_last = n_chunks;

    PlaneArgs pa;
    pa.n_modes = prm->n_modes;
    for (int m = 0; m < 4; ++m) pa.modes[m] = prm->modes[m];
    pa.T0 = t0; pa.T1 = t1; pa.s0 = prm->s0; pa.s1 = prm->s1; pa.len0 = info.len0; pa.len1 = info.len1;
    pa.out_h = info.out_h; pa.out_w = info.out_w; pa.sub_pix = prm->sub_pix;

    int launches = 0;
    // stacked scene row (exclusive) up to which the tiles [.., last] read their input
    auto rows_needed = [&](long long last_tile) -> long long {
        const long long tps = (long long)info.len0 * info.len1, sc = last_tile / tps, r = last_tile - sc * tps;
        return sc * prm->scene_h + (long long)prm->s0 * (r / info.len1) + t0 + prm->ws - 1;
    };
    if ((rc = dm_upload_rows(ctx, rows_needed(ta + chunk_tiles(0) - 1))) != DM_OK) return rc;
    for (int ck = 0; ck < n_chunks; ++ck) {
        const int first = (int)ta + chunk_begin(ck);
        const int nt = chunk_tiles(ck);
        pa.first_tile = first;
        if ((rc = dm_upload_wait(ctx)) != DM_OK) return rc;
        // the next chunk's input rows travel while this chunk is being solved (enqueued after this
        // chunk's launches: a copy from pageable memory holds the host thread until it is staged)
        auto upload_next = [&]() -> int {
            return ck + 1 < n_chunks ? dm_upload_rows(ctx, rows_needed((long long)first + nt + chunk_tiles(ck + 1) - 1)) : DM_OK;
        };
        if (fused) {
            dm_fused_args fa;
            fa.img1 = img1_dev; fa.img2 = img2_dev; fa.scene_h = prm->scene_h; fa.scene_w = prm->scene_w; fa.n_scenes = ns;
            fa.t0 = t0; fa.t1 = t1; fa.ws = prm->ws; fa.kpad = kpad; fa.levels = L; fa.method = prm->method;
            fa.first_tile = first; fa.n_tiles = nt; fa.len0 = info.len0; fa.len1 = info.len1;
            fa.s0 = prm->s0; fa.s1 = prm->s1; fa.out_h = info.out_h; fa.out_w = info.out_w;
            fa.n_modes = prm->n_modes; for (int m = 0; m < 4; ++m) fa.modes[m] = prm->modes[m];
            fa.sub_pix = prm->sub_pix; fa.d_map = d_map_dev; fa.out_map = out_map_dev;
            fa.filter_num = filter_num; fa.filter_win = filter_win; fa.filter_mode = filter_mode;
            rc = dm_fused_solve_chunk(ctx, &fa, ck);
            if (rc != DM_OK) return rc;
            if ((rc = upload_next()) != DM_OK) return rc;
            continue;
        }
        carve_tiles(ctx->ws, nt, t0, t1, kpad, L, tb);
        {
            StageTimer tm(ctx, DM_STAGE_DESCRIPTORS);
            if ((rc = tm.begin(ck)) != DM_OK) return rc;
            dm_tile_origin_kernel<<<dm_div_up(nt, 128), 128, 0, st>>>(tb.origin, nt, first, info.len0, info.len1, prm->s0, prm->s1, prm->scene_h);
            DM_LAUNCH_CHECK();
            int n_desc = 0;
            if ((rc = dm_descriptors_both(img1_dev, img2_dev, prm->scene_h * ns, prm->scene_w, prm->scene_w, tb.origin, nt, t0, t1, prm->ws,
                                          tb.desc1, tb.stat1, tb.desc2, tb.stat2, st, &n_desc)) != DM_OK) return rc;
            ctx->launches[DM_STAGE_DESCRIPTORS] += 1 + n_desc;
            if ((rc = tm.end()) != DM_OK) return rc;
        }
        {
            StageTimer tm(ctx, DM_STAGE_CORRELATION);
            if ((rc = tm.begin(ck)) != DM_OK) return rc;
            if ((rc = dm_correlation(tb.desc1, tb.stat1, tb.desc2, tb.stat2, nt, P, kpad, prm->ws, prm->method, DM_CORR_AUTO, tb.level[0], st)) != DM_OK) return rc;
            ctx->launches[DM_STAGE_CORRELATION] += 1;
            if ((rc = tm.end()) != DM_OK) return rc;
        }
        {
            StageTimer tm(ctx, DM_STAGE_NORMALIZE);
            if ((rc = tm.begin(ck)) != DM_OK) return rc;
            if ((rc = dm_minmax_rectify(tb.level[0], (long long)nt * P, P, nullptr, tb.level[0], nullptr, nullptr, st)) != DM_OK) return rc;
            ctx->launches[DM_STAGE_NORMALIZE] += 1;
            if ((rc = tm.end()) != DM_OK) return rc;
        }
        {
            StageTimer tm(ctx, DM_STAGE_AGGREGATE);
            if ((rc = tm.begin(ck)) != DM_OK) return rc;
            for (int k = 0; k + 1 < L; ++k) {
                if ((rc = dm_aggregate(tb.level[k], nt, t0 >> k, t1 >> k, t0 >> k, t1 >> k, 1, tb.level[k + 1], st)) != DM_OK) return rc;
                ctx->launches[DM_STAGE_AGGREGATE] += 1;
            }
            if ((rc = tm.end()) != DM_OK) return rc;
        }
        int cur = 0, scur = 0;          // ping-pong index of the matches / of the scores (the filter moves only the matches)
        {
            StageTimer tm(ctx, DM_STAGE_BACKTRACK);
            if ((rc = tm.begin(ck)) != DM_OK) return rc;
            int filters_left = filter_num;
            auto maybe_filter = [&](int k) -> int {       // misc/Matching.py:91-93,136-138
                if (filters_left <= 0) return DM_OK;
                --filters_left;
                const int h = t0 >> k, w = t1 >> k;
                if (!(h >= filter_win && w >= filter_win)) return DM_OK;
                int r = dm_match_filter(tb.match[cur], nt, h, w, filter_win, filter_mode, tb.match[cur ^ 1], st);
                if (r != DM_OK) return r;
                cur ^= 1;
                ctx->launches[DM_STAGE_BACKTRACK] += 1;
                return DM_OK;
            };
            if ((rc = dm_backtrack_top(tb.level[L - 1], 0, nt, t0 >> (L - 1), t1 >> (L - 1), tb.match[cur], tb.score[scur], st)) != DM_OK) return rc;
            ctx->launches[DM_STAGE_BACKTRACK] += 1;
            if ((rc = maybe_filter(L - 1)) != DM_OK) return rc;
            for (int k = L - 2; k >= 0; --k) {
                if ((rc = dm_backtrack_level(tb.level[k], 0, nt, t0 >> k, t1 >> k, t0 >> k, t1 >> k, tb.match[cur], tb.match[cur ^ 1], tb.score[scur ^ 1], st)) != DM_OK) return rc;
                cur ^= 1; scur ^= 1;
                ctx->launches[DM_STAGE_BACKTRACK] += 1;
                if ((rc = maybe_filter(k)) != DM_OK) return rc;
            }
            if ((rc = tm.end()) != DM_OK) return rc;
        }
        {
            StageTimer tm(ctx, DM_STAGE_PLANES);
            if ((rc = tm.begin(ck)) != DM_OK) return rc;
            const long long total = (long long)nt * P;
            dm_planes_kernel<<<dm_div_up(total, 256), 256, 0, st>>>(tb.level[0], total, pa, tb.match[cur], tb.score[scur], d_map_dev, out_map_dev);
            DM_LAUNCH_CHECK();
            ctx->launches[DM_STAGE_PLANES] += 1;
            if ((rc = tm.end()) != DM_OK) return rc;
        }
        if ((rc = upload_next()) != DM_OK) return rc;
    }
    for (int s = 0; s < DM_STAGE_COUNT; ++s) launches += ctx->launches[s];
    info.used_fused = fused ? 1 : 0;
    info.chunk_tiles = largest;
    info.kernel_launches = launches;
    if (info_out) *info_out = info;
    return DM_OK;
}

extern "C" int dm_solve_scene_host(dm_ctx* ctx, const dm_scene_params* prm,
                                   const uint8_t* img1_host, const uint8_t* img2_host,
                                   double* d_map_host, double* out_map_host, dm_scene_info* info_out) {
    DM_REQUIRE(ctx && prm && img1_host && img2_host && d_map_host && out_map_host, DM_ERR_INVALID, "dm_solve_scene_host: null argument");
    dm_scene_info info;
    int rc = dm_scene_geometry(prm, &info);
    if (rc != DM_OK) return rc;
    cudaStream_t st = ctx->stream;
    const int ns = prm->n_scenes > 1 ? prm->n_scenes : 1;
    const size_t sb = (size_t)prm->scene_h * prm->scene_w * ns;
    if (sb > ctx->scene_bytes) {
        DM_CUDA_CHECK(cudaStreamSynchronize(st));
        cudaFree(ctx->scene1); cudaFree(ctx->scene2); ctx->scene1 = ctx->scene2 = nullptr; ctx->scene_bytes = 0;
        DM_CUDA_CHECK(cudaMalloc(&ctx->scene1, sb));
        DM_CUDA_CHECK(cudaMalloc(&ctx->scene2, sb));
        ctx->scene_bytes = sb;
    }
    const size_t plane = (size_t)info.out_h * info.out_w;
    const size_t pb = plane * (prm->n_modes + 1) * sizeof(double) * ns;
    if (pb > ctx->planes_bytes) {
        DM_CUDA_CHECK(cudaStreamSynchronize(st));
        cudaFree(ctx->planes); ctx->planes = nullptr; ctx->planes_bytes = 0;
        DM_CUDA_CHECK(cudaMalloc(&ctx->planes, pb));
        ctx->planes_bytes = pb;
    }
    // only the rows the tiles of the range read, uploaded chunk by chunk ahead of the compute (dm_upload_rows)
    long long ta, tb;
    if ((rc = dm_tile_range(prm, info.len0, info.len1, &ta, &tb)) != DM_OK) return rc;
    dm_ctx::Upload& up = ctx->up;
    up = dm_ctx::Upload();
    up.active = true; up.h1 = img1_host; up.h2 = img2_host; up.row_bytes = (size_t)prm->scene_w;
    up.rows_done = (long long)prm->s0 * (ta / info.len1);
    up.row_end = ns > 1 ? (long long)prm->scene_h * ns : (long long)prm->s0 * ((tb - 1) / info.len1) + prm->t0 + prm->ws - 1;
    double* d_map = ctx->planes;
    double* out_map = ctx->planes + plane * prm->n_modes * ns;
    if (prm->s0 > prm->t0 || prm->s1 > prm->t1)      // gaps between tiles: np.empty in the reference, zeros here
        DM_CUDA_CHECK(cudaMemsetAsync(ctx->planes, 0, pb, st));
    // what the finished tiles own streams back to the host behind the compute (a batch: scene after scene)
    dm_ctx::Readback& rb = ctx->rb;
    rb = dm_ctx::Readback();
    {
        rb.active = true;
        rb.dst_d_map = d_map_host; rb.dst_out_map = out_map_host; rb.d_d_map = d_map; rb.d_out_map = out_map;
        rb.n_modes = prm->n_modes; rb.out_h = info.out_h; rb.out_w = info.out_w; rb.len0 = info.len0; rb.len1 = info.len1;
        rb.s0 = prm->s0; rb.s1 = prm->s1; rb.n_scenes = ns;
        rb.tiles_copied = ta;
        ctx->band_used = 0;
    }
    rc = dm_solve_scene(ctx, prm, ctx->scene1, ctx->scene2, d_map, out_map, &info);
    if (rc == DM_OK) rc = dm_readback_rows(ctx, tb * ns);     // whatever has not left yet (everything on the materialising path)
    rb.active = false;
    up.active = false;
    if (rc != DM_OK) {
        // copies already queued must not keep reading / writing the caller's arrays after the
        // error has been reported (the caller may release them)
        if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
        if (ctx->h2d_stream) cudaStreamSynchronize(ctx->h2d_stream);
        cudaStreamSynchronize(st);
        return rc;
    }
    if (ctx->copy_stream && ctx->band_used > 0) DM_CUDA_CHECK(cudaStreamSynchronize(ctx->copy_stream));
    DM_CUDA_CHECK(cudaStreamSynchronize(st));
    if (info_out) *info_out = info;
    return DM_OK;
}

// Device-resident solve whose finished rows are also copied, band by band behind the final stage, to
// a second pair of arrays (see dmstereo.h).  The compute stream is made to wait for the last copy, so
// synchronising the ctx stream (or an event recorded on it) covers the copies as well.
extern "C" int dm_solve_scene_stream(dm_ctx* ctx, const dm_scene_params* prm,
                                     const uint8_t* img1_dev, const uint8_t* img2_dev,
                                     double* d_map_dev, double* out_map_dev,
                                     double* d_map_dst, double* out_map_dst, dm_scene_info* info_out) {
    DM_REQUIRE(ctx && prm && d_map_dst && out_map_dst, DM_ERR_INVALID, "dm_solve_scene_stream: null argument");
    DM_REQUIRE(prm->n_scenes <= 1, DM_ERR_UNSUPPORTED, "dm_solve_scene_stream: a batch of scenes has no row strips to stream");
    dm_scene_info info;
    int rc = dm_scene_geometry(prm, &info);
    if (rc != DM_OK) return rc;
    dm_ctx::Readback& rb = ctx->rb;
    rb = dm_ctx::Readback();
    rb.active = true;
    rb.dst_d_map = d_map_dst; rb.dst_out_map = out_map_dst; rb.d_d_map = d_map_dev; rb.d_out_map = out_map_dev;
    rb.n_modes = prm->n_modes; rb.out_h = info.out_h; rb.out_w = info.out_w; rb.len0 = info.len0; rb.len1 = info.len1;
    rb.s0 = prm->s0; rb.s1 = prm->s1;
    long long ta, tb;
    if ((rc = dm_tile_range(prm, info.len0, info.len1, &ta, &tb)) != DM_OK) return rc;
    rb.tiles_copied = ta;
    ctx->band_used = 0;
    ctx->up.active = false;
    rc = dm_solve_scene(ctx, prm, img1_dev, img2_dev, d_map_dev, out_map_dev, &info);
    if (rc == DM_OK) rc = dm_readback_rows(ctx, tb);     // whatever has not left yet (materialising path: everything)
    rb.active = false;
    if (ctx->copy_stream && ctx->band_used > 0) {
        // the compute stream waits for the copies: one stream to synchronise on for the caller
        cudaEvent_t ev = ctx->band_ev[ctx->band_used - 1];
        cudaError_t e1 = cudaEventRecord(ev, ctx->copy_stream);
        cudaError_t e2 = cudaStreamWaitEvent(ctx->stream, ev, 0);
        if (rc == DM_OK) { DM_CUDA_CHECK(e1); DM_CUDA_CHECK(e2); }
    }
    if (rc != DM_OK) return rc;
    if (info_out) *info_out = info;
    return DM_OK;
}

// sub_pix_cal (misc/sub_pix_cal.py:22-53) for a batch of planes in HOST memory (config 4: 64 pairs, two planes
// each).  The batch goes through the device in pieces on two streams, each with its own staging slot: the upload
// of piece k+1, the kernels of piece k and the download of piece k-1 share the PCIe link in both directions.
extern "C" int dm_sub_pix_cal_host_batch(dm_ctx* ctx, const double* d_maps_host, const double* co_maps_host, int n, int n_planes,
                                         int s0, int s1, const int32_t* directions, double ratio, double* out_host) {
    DM_REQUIRE(ctx && d_maps_host && co_maps_host && out_host && directions && n > 0 && n_planes > 0 && s0 > 0 && s1 > 0,
               DM_ERR_INVALID, "dm_sub_pix_cal_host_batch: bad arguments");
    for (int m = 0; m < n_planes; ++m)
        DM_REQUIRE(directions[m] == 0 || directions[m] == 1, DM_ERR_INVALID, "dm_sub_pix_cal_host_batch: direction %d", directions[m]);
    const size_t plane = (size_t)s0 * s1;
    int piece = (int)((((size_t)8 << 20) + plane * sizeof(double) - 1) / (plane * sizeof(double)));      // ~8 MB of scores per piece
    if (piece > n) piece = n;
    if (piece * 8 > n && n >= 8) piece = n / 8;                 // at least eight pieces in flight over the two streams
    if (piece < 1) piece = 1;
    const size_t slot = (size_t)piece * (2 * n_planes + 1) * plane;         // doubles: planes in, scores in, planes out
    if (2 * slot * sizeof(double) > ctx->sp_bytes) {
        for (int k = 0; k < 2; ++k) if (ctx->sp_stream[k]) DM_CUDA_CHECK(cudaStreamSynchronize(ctx->sp_stream[k]));
        cudaFree(ctx->sp_buf); ctx->sp_buf = nullptr; ctx->sp_bytes = 0;
        DM_CUDA_CHECK(cudaMalloc(&ctx->sp_buf, 2 * slot * sizeof(double)));
        ctx->sp_bytes = 2 * slot * sizeof(double);
    }
    for (int k = 0; k < 2; ++k)
        if (!ctx->sp_stream[k]) DM_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->sp_stream[k], cudaStreamNonBlocking));
    int rc = DM_OK;
    for (int b0 = 0, k = 0; b0 < n && rc == DM_OK; b0 += piece, ++k) {
        const int nb = b0 + piece <= n ? piece : n - b0;
        cudaStream_t st = ctx->sp_stream[k & 1];
        double* din = ctx->sp_buf + (size_t)(k & 1) * slot;
        double* cin = din + (size_t)piece * n_planes * plane;
        double* dout = cin + (size_t)piece * plane;
        DM_CUDA_CHECK(cudaMemcpyAsync(din, d_maps_host + (size_t)b0 * n_planes * plane, (size_t)nb * n_planes * plane * sizeof(double), cudaMemcpyHostToDevice, st));
        DM_CUDA_CHECK(cudaMemcpyAsync(cin, co_maps_host + (size_t)b0 * plane, (size_t)nb * plane * sizeof(double), cudaMemcpyHostToDevice, st));
        for (int b = 0; b < nb && rc == DM_OK; ++b)
            for (int m = 0; m < n_planes && rc == DM_OK; ++m)
                rc = dm_sub_pix_cal(din + ((size_t)b * n_planes + m) * plane, cin + (size_t)b * plane, s0, s1, directions[m], ratio,
                                    dout + ((size_t)b * n_planes + m) * plane, st);
        if (rc == DM_OK)
            DM_CUDA_CHECK(cudaMemcpyAsync(out_host + (size_t)b0 * n_planes * plane, dout, (size_t)nb * n_planes * plane * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    for (int k = 0; k < 2; ++k) {
        cudaError_t e = cudaStreamSynchronize(ctx->sp_stream[k]);
        if (rc == DM_OK) DM_CUDA_CHECK(e);
    }
    return rc;
}

// geometry only: the rectangles of the mosaic the tiles of `prm` own (rows [r0,r1) x columns [c0,c1) each)
extern "C" int dm_owned_rectangles(const dm_scene_params* prm, int32_t* rects /* [3][4] = {r0, r1, c0, c1} */, int32_t* n_rects) {
    DM_REQUIRE(prm && rects && n_rects, DM_ERR_INVALID, "dm_owned_rectangles: null argument");
    dm_scene_info info;
    int rc = dm_scene_geometry(prm, &info);
    if (rc != DM_OK) return rc;
    long long ta, tb;
    if ((rc = dm_tile_range(prm, info.len0, info.len1, &ta, &tb)) != DM_OK) return rc;
    OwnedRect r[3];
    const int n = owned_rects(info.len0, info.len1, prm->s0, prm->s1, info.out_h, info.out_w, ta, tb, r);
    for (int k = 0; k < n; ++k) { rects[4 * k] = r[k].r0; rects[4 * k + 1] = r[k].r1; rects[4 * k + 2] = r[k].c0; rects[4 * k + 3] = r[k].c1; }
    *n_rects = n;
    return DM_OK;
}
