// Several devices in one process, and mosaics shared between processes.
//
// The reference solves its tiles one after the other (misc/image_cut_solver.py:144-184, a serial
// trange over the tile list).  Tiles are independent and every output row is owned by exactly one
// tile row (the covering tile with the largest index, :165-175), so the tile rows are cut into
// contiguous strips and each device solves one strip with its own dm_ctx, stream and host thread.
// The strips meet
//   - in the caller's host arrays (dm_multi_solve_scene_host: every device copies its own rows
//     over its own PCIe link, no collective), or
//   - in the planes of a root device (dm_multi_solve_scene): finished bands streamed into the
//     root's memory over NVLink while the strip is still being solved (DM_GATHER_P2P), or one
//     grouped ncclSend / ncclRecv after the solve (DM_GATHER_NCCL; libnccl.so.2 is opened with
//     dlopen on first use, so the library has no link-time dependency on NCCL).
#include <dlfcn.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "dm_common.cuh"
#include "dm_internal.h"

namespace {

// ------------------------------------------------------------------ NCCL through dlopen
typedef void* nccl_comm_t;
enum { NCCL_FLOAT64 = 8 };          // ncclDataType_t::ncclFloat64 (nccl.h)
struct NcclApi {
    void* handle = nullptr;
    int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

int load_nccl(NcclApi& api) {
    if (api.handle) return DM_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    DM_REQUIRE(h != nullptr, DM_ERR_UNSUPPORTED, "DM_GATHER_NCCL: cannot load libnccl.so.2 (%s)", dlerror());
    bool ok = true;
    auto sym = [&](const char* name) -> void* { void* p = dlsym(h, name); if (!p) ok = false; return p; };
    api.CommInitAll = (int (*)(nccl_comm_t*, int, const int*))sym("ncclCommInitAll");
    api.CommDestroy = (int (*)(nccl_comm_t))sym("ncclCommDestroy");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.Send = (int (*)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t))sym("ncclSend");
    api.Recv = (int (*)(void*, size_t, int, int, nccl_comm_t, cudaStream_t))sym("ncclRecv");
    api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    DM_REQUIRE(ok, DM_ERR_UNSUPPORTED, "DM_GATHER_NCCL: libnccl.so.2 lacks a required symbol");
    api.handle = h;
    return DM_OK;
}

#define DM_NCCL_CHECK(api, expr)                                                              \
    do {                                                                                      \
        int r__ = (expr);                                                                     \
        if (r__ != 0) {                                                                       \
            dm_set_error("%s failed: %s", #expr, (api).GetErrorString ? (api).GetErrorString(r__) : "?"); \
            return DM_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

// ------------------------------------------------------------------ one host thread per device
struct Worker {
    int device = 0;
    dm_ctx* ctx = nullptr;
    cudaStream_t stream = nullptr;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = true, quit = false, ready = false;
    int rc = DM_OK;
    std::string err;

    void loop() {
        int rc0 = DM_OK;
        if (cudaSetDevice(device) != cudaSuccess) { rc0 = DM_ERR_CUDA; err = "cudaSetDevice failed"; }
        if (rc0 == DM_OK && (rc0 = dm_ctx_create(&ctx)) != DM_OK) err = dm_last_error();
        if (rc0 == DM_OK && cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) { rc0 = DM_ERR_CUDA; err = "cudaStreamCreate failed"; }
        if (rc0 == DM_OK) dm_ctx_set_stream(ctx, stream);
        {
            std::lock_guard<std::mutex> lk(mu);
            rc = rc0; ready = true;
        }
        cv.notify_all();
        for (;;) {
            std::function<int()> j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return has_job || quit; });
                if (quit) break;
                j = job; has_job = false;
            }
            int r = j();
            std::string e = r != DM_OK ? dm_last_error() : "";
            {
                std::lock_guard<std::mutex> lk(mu);
                rc = r; err = e; done = true;
            }
            cv.notify_all();
        }
        if (ctx) { cudaStreamSynchronize(stream); dm_ctx_destroy(ctx); }
        if (stream) cudaStreamDestroy(stream);
    }
    void submit(std::function<int()> j) {
        {
            std::lock_guard<std::mutex> lk(mu);
            job = std::move(j); has_job = true; done = false;
        }
        cv.notify_all();
    }
    int wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done; });
        return rc;
    }
};

}  // namespace

struct dm_multi {
    std::vector<Worker*> w;
    NcclApi nccl;
    std::vector<nccl_comm_t> comms;
    std::vector<std::vector<char>> peer_ok;     // peer_ok[r][root]: peer access from device r to root enabled
};

namespace {

// runs jobs[r] (empty = nothing to do) on worker r, returns the first failure
int run_all(dm_multi* m, std::vector<std::function<int()>>& jobs) {
    for (size_t r = 0; r < jobs.size(); ++r)
        if (jobs[r]) m->w[r]->submit(jobs[r]);
    int rc = DM_OK;
    for (size_t r = 0; r < jobs.size(); ++r) {
        if (!jobs[r]) continue;
        const int rr = m->w[r]->wait();
        if (rr != DM_OK && rc == DM_OK) {
            rc = rr;
            dm_set_error("device %d: %s", m->w[r]->device, m->w[r]->err.c_str());
        }
    }
    return rc;
}

void partition(int len0, int n, int32_t* lo, int32_t* hi) {
    const int base = len0 / n, extra = len0 % n;
    int at = 0;
    for (int r = 0; r < n; ++r) {
        lo[r] = at;
        at += base + (r < extra ? 1 : 0);
        hi[r] = at;
    }
}

}  // namespace

extern "C" int dm_partition_tile_rows(int len0, int n, int32_t* lo, int32_t* hi) {
    DM_REQUIRE(len0 >= 0 && n >= 1 && lo && hi, DM_ERR_INVALID, "dm_partition_tile_rows: bad arguments");
    partition(len0, n, lo, hi);
    return DM_OK;
}

extern "C" int dm_multi_create(const int* devices, int n_devices, dm_multi** out) {
    DM_REQUIRE(out != nullptr, DM_ERR_INVALID, "dm_multi_create: null out");
    int visible = 0;
    DM_CUDA_CHECK(cudaGetDeviceCount(&visible));
    if (n_devices <= 0) { n_devices = visible; devices = nullptr; }
    DM_REQUIRE(n_devices >= 1 && n_devices <= 64, DM_ERR_INVALID, "dm_multi_create: %d devices", n_devices);
    dm_multi* m = new dm_multi();
    for (int r = 0; r < n_devices; ++r) {
        const int dev = devices ? devices[r] : r;
        if (dev < 0 || dev >= visible) {
            dm_multi_destroy(m);
            DM_REQUIRE(false, DM_ERR_INVALID, "dm_multi_create: device %d is not visible (%d devices)", dev, visible);
        }
        Worker* w = new Worker();
        w->device = dev;
        m->w.push_back(w);
        w->th = std::thread([w] { w->loop(); });
    }
    int rc = DM_OK;
    for (Worker* w : m->w) {
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return w->ready; });
        if (w->rc != DM_OK && rc == DM_OK) { rc = w->rc; dm_set_error("device %d: %s", w->device, w->err.c_str()); }
    }
    if (rc != DM_OK) { dm_multi_destroy(m); return rc; }
    m->peer_ok.assign(n_devices, std::vector<char>(n_devices, 0));
    *out = m;
    return DM_OK;
}

extern "C" void dm_multi_destroy(dm_multi* m) {
    if (!m) return;
    for (Worker* w : m->w) {
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->quit = true;
        }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    if (m->nccl.CommDestroy)
        for (nccl_comm_t c : m->comms) if (c) m->nccl.CommDestroy(c);
    delete m;
}

extern "C" int dm_multi_device_count(const dm_multi* m) { return m ? (int)m->w.size() : 0; }

extern "C" int dm_multi_set_workspace_limit(dm_multi* m, size_t bytes) {
    DM_REQUIRE(m != nullptr, DM_ERR_INVALID, "dm_multi_set_workspace_limit: null");
    for (Worker* w : m->w) dm_ctx_set_workspace_limit(w->ctx, bytes);
    return DM_OK;
}

extern "C" int dm_multi_synchronize(dm_multi* m) {
    DM_REQUIRE(m != nullptr, DM_ERR_INVALID, "dm_multi_synchronize: null");
    std::vector<std::function<int()>> jobs(m->w.size());
    for (size_t r = 0; r < m->w.size(); ++r) {
        Worker* w = m->w[r];
        jobs[r] = [w]() -> int { DM_CUDA_CHECK(cudaStreamSynchronize(w->stream)); return DM_OK; };
    }
    return run_all(m, jobs);
}

// how a scene (or a batch of scenes) is spread over `n` devices
struct Share { dm_scene_params prm; size_t img_off, dmap_off, omap_off; bool any; };

// by_rows: whole tile rows per device (what a gather of contiguous row strips needs: DM_GATHER_NCCL);
// otherwise the tiles themselves are shared out, so the shares differ by at most one tile (66 tile rows
// over 8 devices are 9,9,8,... rows = 9 % imbalance; 4356 tiles over 8 are 545,545,545,545,544,...).
static int make_shares(const dm_scene_params* prm, const dm_scene_info& info, int n, bool by_rows, std::vector<Share>& sh) {
    sh.assign(n, Share());
    const int ns = prm->n_scenes > 1 ? prm->n_scenes : 1;
    const size_t plane = (size_t)info.out_h * info.out_w;
    std::vector<int32_t> lo(n), hi(n);
    if (ns > 1) {
        // a batch of pairs: whole pairs per device
        partition(ns, n, lo.data(), hi.data());
        for (int r = 0; r < n; ++r) {
            sh[r].any = hi[r] > lo[r];
            sh[r].prm = *prm;
            sh[r].prm.n_scenes = hi[r] - lo[r];
            sh[r].img_off = (size_t)lo[r] * prm->scene_h * prm->scene_w;
            sh[r].dmap_off = (size_t)lo[r] * prm->n_modes * plane;
            sh[r].omap_off = (size_t)lo[r] * plane;
        }
        return DM_OK;
    }
    long long ta, tb;
    int rc = dm_tile_range(prm, info.len0, info.len1, &ta, &tb);
    if (rc != DM_OK) return rc;
    if (by_rows) {
        DM_REQUIRE(ta % info.len1 == 0 && tb % info.len1 == 0, DM_ERR_UNSUPPORTED, "a gather of row strips needs whole tile rows");
        const int tlo = (int)(ta / info.len1), thi = (int)(tb / info.len1);
        partition(thi - tlo, n, lo.data(), hi.data());
        for (int r = 0; r < n; ++r) {
            sh[r].any = hi[r] > lo[r];
            sh[r].prm = *prm;
            sh[r].prm.tile_lo = sh[r].prm.tile_hi = 0;
            sh[r].prm.tile_row_lo = tlo + lo[r];
            sh[r].prm.tile_row_hi = tlo + hi[r];
        }
        return DM_OK;
    }
    partition((int)(tb - ta), n, lo.data(), hi.data());
    for (int r = 0; r < n; ++r) {
        sh[r].any = hi[r] > lo[r];
        sh[r].prm = *prm;
        sh[r].prm.tile_row_lo = sh[r].prm.tile_row_hi = 0;
        sh[r].prm.tile_lo = (int32_t)ta + lo[r];
        sh[r].prm.tile_hi = (int32_t)ta + hi[r];
    }
    return DM_OK;
}

// whole-scene info from the geometry plus what the devices report
static dm_scene_info merge_infos(const dm_scene_info& geometry, const std::vector<Share>& sh, const std::vector<dm_scene_info>& infos) {
    dm_scene_info total = geometry;
    total.kernel_launches = 0; total.chunk_tiles = 0; total.used_fused = 1; total.n_tiles = 0;
    for (size_t r = 0; r < sh.size(); ++r) {
        if (!sh[r].any) continue;
        total.kernel_launches += infos[r].kernel_launches;
        total.n_tiles += infos[r].n_tiles;
        if (infos[r].chunk_tiles > total.chunk_tiles) total.chunk_tiles = infos[r].chunk_tiles;
        total.used_fused = total.used_fused && infos[r].used_fused;
    }
    return total;
}

extern "C" int dm_multi_solve_scene_host(dm_multi* m, const dm_scene_params* prm, int max_devices,
                                         const uint8_t* img1_host, const uint8_t* img2_host,
                                         double* d_map_host, double* out_map_host, dm_scene_info* info_out) {
    DM_REQUIRE(m && prm && img1_host && img2_host && d_map_host && out_map_host, DM_ERR_INVALID, "dm_multi_solve_scene_host: null argument");
    dm_scene_info info;
    int rc = dm_scene_geometry(prm, &info);
    if (rc != DM_OK) return rc;
    int n = (int)m->w.size();
    if (max_devices > 0 && max_devices < n) n = max_devices;
    std::vector<Share> sh;
    if ((rc = make_shares(prm, info, n, false, sh)) != DM_OK) return rc;
    std::vector<dm_scene_info> infos(n);
    std::vector<std::function<int()>> jobs(n);
    for (int r = 0; r < n; ++r) {
        if (!sh[r].any) continue;
        Worker* w = m->w[r];
        const Share* s = &sh[r];
        dm_scene_info* inf = &infos[r];
        jobs[r] = [=]() -> int {
            return dm_solve_scene_host(w->ctx, &s->prm, img1_host + s->img_off, img2_host + s->img_off,
                                       d_map_host + s->dmap_off, out_map_host + s->omap_off, inf);
        };
    }
    if ((rc = run_all(m, jobs)) != DM_OK) return rc;
    if (info_out) *info_out = merge_infos(info, sh, infos);
    return DM_OK;
}

static int ensure_nccl(dm_multi* m) {
    if (!m->comms.empty()) return DM_OK;
    int rc = load_nccl(m->nccl);
    if (rc != DM_OK) return rc;
    const int n = (int)m->w.size();
    std::vector<int> devs(n);
    for (int r = 0; r < n; ++r) devs[r] = m->w[r]->device;
    m->comms.assign(n, nullptr);
    int prev = 0;
    cudaGetDevice(&prev);
    const int r = m->nccl.CommInitAll(m->comms.data(), n, devs.data());
    cudaSetDevice(prev);
    if (r != 0) {
        m->comms.clear();
        dm_set_error("ncclCommInitAll over %d devices failed: %s", n, m->nccl.GetErrorString(r));
        return DM_ERR_CUDA;
    }
    return DM_OK;
}

extern "C" int dm_multi_gather_strips(dm_multi* m, double* const* planes_dev, int n_planes, int out_h, int out_w,
                                      const int32_t* row_lo, const int32_t* row_hi, int root) {
    DM_REQUIRE(m && planes_dev && row_lo && row_hi, DM_ERR_INVALID, "dm_multi_gather_strips: null argument");
    const int n = (int)m->w.size();
    DM_REQUIRE(root >= 0 && root < n && n_planes >= 1 && out_h > 0 && out_w > 0, DM_ERR_INVALID, "dm_multi_gather_strips: bad arguments");
    if (n == 1) return DM_OK;
    int rc = ensure_nccl(m);
    if (rc != DM_OK) return rc;
    int prev = 0;
    cudaGetDevice(&prev);
    const size_t plane = (size_t)out_h * out_w;
    NcclApi& api = m->nccl;
    // one group: device r sends rows [row_lo[r], row_hi[r]) of every plane, the root receives them in place
    DM_NCCL_CHECK(api, api.GroupStart());
    int bad = 0;
    for (int r = 0; r < n && !bad; ++r) {
        if (r == root || row_hi[r] <= row_lo[r]) continue;
        const size_t off = (size_t)row_lo[r] * out_w, count = (size_t)(row_hi[r] - row_lo[r]) * out_w;
        for (int p = 0; p < n_planes && !bad; ++p) {
            bad = api.Send(planes_dev[r] + p * plane + off, count, NCCL_FLOAT64, root, m->comms[r], m->w[r]->stream);
            if (!bad) bad = api.Recv(planes_dev[root] + p * plane + off, count, NCCL_FLOAT64, r, m->comms[root], m->w[root]->stream);
        }
    }
    const int ge = api.GroupEnd();
    cudaSetDevice(prev);
    DM_NCCL_CHECK(api, bad);
    DM_NCCL_CHECK(api, ge);
    return DM_OK;
}

extern "C" int dm_multi_solve_scene(dm_multi* m, const dm_scene_params* prm,
                                    const uint8_t* const* img1_dev, const uint8_t* const* img2_dev,
                                    double* const* planes_dev, int root, int gather, dm_scene_info* info_out) {
    DM_REQUIRE(m && prm && img1_dev && img2_dev && planes_dev, DM_ERR_INVALID, "dm_multi_solve_scene: null argument");
    DM_REQUIRE(prm->n_scenes <= 1, DM_ERR_UNSUPPORTED, "dm_multi_solve_scene: single scenes only (a batch has no strips to gather)");
    DM_REQUIRE(gather == DM_GATHER_P2P || gather == DM_GATHER_NCCL, DM_ERR_INVALID, "dm_multi_solve_scene: gather %d", gather);
    dm_scene_info info;
    int rc = dm_scene_geometry(prm, &info);
    if (rc != DM_OK) return rc;
    const int n = (int)m->w.size();
    DM_REQUIRE(root >= 0 && root < n, DM_ERR_INVALID, "dm_multi_solve_scene: root %d of %d devices", root, n);
    std::vector<Share> sh;
    if ((rc = make_shares(prm, info, n, gather == DM_GATHER_NCCL, sh)) != DM_OK) return rc;
    const size_t plane = (size_t)info.out_h * info.out_w;
    std::vector<dm_scene_info> infos(n);
    std::vector<std::function<int()>> jobs(n);
    const int root_dev = m->w[root]->device;
    for (int r = 0; r < n; ++r) {
        if (!sh[r].any) continue;
        Worker* w = m->w[r];
        const Share* s = &sh[r];
        dm_scene_info* inf = &infos[r];
        const uint8_t* i1 = img1_dev[r]; const uint8_t* i2 = img2_dev[r];
        double* mine = planes_dev[r]; double* dst = planes_dev[root];
        const int nm = prm->n_modes;
        char* peer_flag = &m->peer_ok[r][root];
        if (r == root || gather == DM_GATHER_NCCL) {
            jobs[r] = [=]() -> int { return dm_solve_scene(w->ctx, &s->prm, i1, i2, mine, mine + nm * plane, inf); };
        } else {
            jobs[r] = [=]() -> int {
                if (!*peer_flag) {
                    int can = 0;
                    DM_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, w->device, root_dev));
                    DM_REQUIRE(can, DM_ERR_UNSUPPORTED, "DM_GATHER_P2P: device %d cannot access device %d", w->device, root_dev);
                    cudaError_t e = cudaDeviceEnablePeerAccess(root_dev, 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
                    DM_CUDA_CHECK(e);
                    *peer_flag = 1;
                }
                return dm_solve_scene_stream(w->ctx, &s->prm, i1, i2, mine, mine + nm * plane, dst, dst + nm * plane, inf);
            };
        }
    }
    if ((rc = run_all(m, jobs)) != DM_OK) return rc;
    if (gather == DM_GATHER_NCCL) {
        std::vector<int32_t> rlo(n, 0), rhi(n, 0);
        for (int r = 0; r < n; ++r) if (sh[r].any) { rlo[r] = infos[r].row_lo; rhi[r] = infos[r].row_hi; }
        if ((rc = dm_multi_gather_strips(m, planes_dev, prm->n_modes + 1, info.out_h, info.out_w, rlo.data(), rhi.data(), root)) != DM_OK) return rc;
    }
    if (info_out) *info_out = merge_infos(info, sh, infos);
    return DM_OK;
}

// ------------------------------------------------------------------ mosaics shared between processes
extern "C" int dm_ipc_alloc(size_t bytes, void** dev_ptr) {
    DM_REQUIRE(dev_ptr && bytes > 0, DM_ERR_INVALID, "dm_ipc_alloc: bad arguments");
    DM_CUDA_CHECK(cudaMalloc(dev_ptr, bytes));
    return DM_OK;
}
extern "C" int dm_ipc_free(void* dev_ptr) {
    if (dev_ptr) DM_CUDA_CHECK(cudaFree(dev_ptr));
    return DM_OK;
}
extern "C" int dm_ipc_export(void* dev_ptr, unsigned char* handle) {
    static_assert(sizeof(cudaIpcMemHandle_t) == DM_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    DM_REQUIRE(dev_ptr && handle, DM_ERR_INVALID, "dm_ipc_export: null argument");
    cudaIpcMemHandle_t h;
    DM_CUDA_CHECK(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle, &h, sizeof(h));
    return DM_OK;
}
extern "C" int dm_ipc_open(const unsigned char* handle, void** dev_ptr) {
    DM_REQUIRE(dev_ptr && handle, DM_ERR_INVALID, "dm_ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    DM_CUDA_CHECK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DM_OK;
}
extern "C" int dm_ipc_close(void* dev_ptr) {
    if (dev_ptr) DM_CUDA_CHECK(cudaIpcCloseMemHandle(dev_ptr));
    return DM_OK;
}
