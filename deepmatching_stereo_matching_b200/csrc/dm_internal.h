// Internal (non-ABI) declarations shared between the translation units of libdmstereo.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/dmstereo.h"

int dm_correlation_simt(const void* desc1, const float* stat1, const void* desc2, const float* stat2,
                        int n_tiles, int p, int kpad, int ws, int method, float* raw, cudaStream_t stream);

// tcgen05 / TMEM / TMA engine (correlation_umma.cu)
bool dm_correlation_umma_supported(int p, int kpad);
int dm_correlation_umma(const void* desc1, const float* stat1, const void* desc2, const float* stat2,
                        int n_tiles, int p, int kpad, int kreal, int method, float* raw, cudaStream_t stream);

int dm_correlation_umma_null(const void* desc1, const float* stat1, const void* desc2, const float* stat2,
                             int n_tiles, int p, int kpad, int kreal, float* raw, cudaStream_t stream);
// pooled epilogue: level 0 is max-pooled on the fly, only [n][P][P/4] + row min/max reach HBM
bool dm_correlation_umma_pool_supported(int t0, int t1, int kpad);
int dm_correlation_umma_pool(const void* desc1, const float* stat1, const void* desc2, const float* stat2,
                             int n_tiles, int t0, int t1, int kpad, int kreal, int method, int engine,
                             float* pooled, float* rowmin, float* rowmax, cudaStream_t stream);

void dm_correlation_umma_set_pair_mode(int mode);     // -1 auto (CTA pairs when possible), 0 single CTA
// number of tiles of P patches whose work units fill whole rounds of the persistent grid of this device
int dm_correlation_round_tiles(int p);
// the tiles [a,b) (row-major index of one scene) a solve covers: a tile range, a strip of tile rows, or all (capi.cu)
int dm_tile_range(const dm_scene_params* prm, int len0, int len1, long long* a, long long* b);

// upper pyramid tail + top-down pass of one tile per CTA (backtrack.cu); levels_dev[k] = level k of the fused workspace
bool dm_upper_tail_supported(int t0, int t1, int levels);
int dm_upper_tail_first_level(int t0, int t1, int levels);
int dm_upper_tail(float* const* levels_dev, int n_tiles, int t0, int t1, int levels, int32_t* match1_dev, cudaStream_t stream);

int dm_descriptors_both(const uint8_t* img1, const uint8_t* img2, int scene_h, int scene_w, int pitch,
                        const int32_t* origin_yx_dev, int n_tiles, int t0, int t1, int ws,
                        void* desc1, float* stat1, void* desc2, float* stat2, cudaStream_t st, int* launches);   // descriptors.cu
int dm_desc_kreal(int ws);      // ws * row stride of the descriptor K layout (descriptors.cu)

struct dm_ctx {
    cudaStream_t stream = nullptr;
    char* ws = nullptr;             // workspace
    size_t ws_bytes = 0;
    size_t ws_limit = (size_t)48 << 30;
    // host-API staging
    uint8_t* scene1 = nullptr; uint8_t* scene2 = nullptr; size_t scene_bytes = 0;
    double* planes = nullptr; size_t planes_bytes = 0;
    // Finished output rows streamed behind the final stage (fused path) to a second destination:
    // page-locked host arrays (dm_solve_scene_host) or the mosaic of a peer device / another process
    // (dm_solve_scene_stream; UVA pointers).  Copies run on a second stream while the last tiles are
    // still being solved.
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> band_ev;
    size_t band_used = 0;
    struct Readback {
        bool active = false;
        double* dst_d_map = nullptr; double* dst_out_map = nullptr;
        const double* d_d_map = nullptr; const double* d_out_map = nullptr;
        int n_modes = 0, out_h = 0, out_w = 0, len0 = 0, len1 = 0, s0 = 0, s1 = 0, n_scenes = 1;
        long long tiles_copied = 0;                 // the pixels owned by tiles [.., tiles_copied) (global index) are already on their way
    } rb;
    // Host scenes uploaded chunk by chunk ahead of the compute (dm_solve_scene_host): the rows the
    // tiles of chunk k+1 read are copied on a third stream while chunk k is being solved.
    cudaStream_t h2d_stream = nullptr;
    cudaEvent_t up_ev = nullptr;
    struct Upload {
        bool active = false, pending = false;
        const uint8_t* h1 = nullptr; const uint8_t* h2 = nullptr;
        size_t row_bytes = 0;
        long long rows_done = 0, row_end = 0;       // stacked scene rows [.., rows_done) are on the device (or on their way)
    } up;
    int device = 0;                                 // the CUDA device this context was created on
    // dm_sub_pix_cal_host_batch: two streams, each with its own staging slot (planes in, scores in, planes out)
    cudaStream_t sp_stream[2] = {nullptr, nullptr};
    double* sp_buf = nullptr; size_t sp_bytes = 0;
    // CUDA graphs of the upper-pyramid + top-down launch sequence (fused.cu), keyed by what they depend on
    struct UpperGraph {
        const void* ws; int nt, t0, t1, levels, kpad, filter_num, filter_win, filter_mode;   // kpad: the carve-up of the workspace (level / match offsets) depends on it
        int n_agg, n_bt, final_cur;
        cudaGraphExec_t exec;
    };
    std::vector<UpperGraph> upper_graphs;
    cudaStream_t capture_stream = nullptr;
    // timing
    bool timing = false;
    std::vector<cudaEvent_t> ev[DM_STAGE_COUNT];   // pairs (start, stop) per chunk
    int launches[DM_STAGE_COUNT] = {};
    size_t chunks_last = 0;
};

// Records a CUDA-event pair around one stage of one chunk when ctx->timing is on.
struct StageTimer {
    dm_ctx* ctx; int stage; size_t slot;
    StageTimer(dm_ctx* c, int s) : ctx(c), stage(s), slot(0) {}
    int begin(size_t chunk);
    int end();
};

// Fused solver (fused.cu): level 0 is never written to HBM.
struct dm_fused_args {
    const uint8_t* img1; const uint8_t* img2; int scene_h, scene_w, n_scenes;
    int t0, t1, ws, kpad, levels, method;
    int first_tile, n_tiles, len0, len1, s0, s1, out_h, out_w;
    int n_modes, modes[4], sub_pix;
    int filter_num, filter_win, filter_mode;     // Matching._filter on the first filter_num maps (< levels on this path)
    double* d_map; double* out_map;
};
bool dm_fused_supported(int t0, int t1, int kpad);
bool dm_fused_supported_ws(int ws);
size_t dm_fused_workspace(char* base, int n_tiles, int t0, int t1, int kpad, int levels, void* buffers_out);
int dm_fused_solve_chunk(dm_ctx* ctx, const dm_fused_args* a, int chunk_index);
// schedules the copy of the output rows completed by tiles [0, tiles_done) (global tile index) to the stream destination
int dm_readback_rows(dm_ctx* ctx, long long tiles_done);
