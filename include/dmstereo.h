/*
 * dmstereo.h -- C-ABI of libdmstereo.so, the B200 (sm_100a) implementation of the
 * DeepMatching-for-stereo hot path of Yuki-Kumon/deepmatching_stereo_matching.
 *
 * The reference has no FFI of its own: its boundary is the Python class API
 * (SURVEY.md section 8(b)).  Each entry point below names the reference function it
 * replaces (paths relative to the reference repository).  The Python mirror of that API
 * (package deepmatching_stereo_matching_b200, re-exported as misc.*) binds these symbols
 * through ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross the boundary.
 *   - "dev" pointers are CUDA device pointers on the current device; "host" pointers are
 *     ordinary (ideally pinned) host memory.
 *   - stream is a cudaStream_t passed as void* (NULL = legacy default stream).  Stage
 *     functions only enqueue work on it and return; they never synchronise.
 *   - return value: 0 on success, <0 on error (DM_ERR_*); dm_last_error() returns a
 *     thread-local human readable message.  Nothing throws.
 *   - 4-D correlation maps are row-major [n][A][B][C][D]: n tiles, (A,B) the patch grid
 *     of image 1, (C,D) the position grid in image 2.
 */
#ifndef DMSTEREO_H_
#define DMSTEREO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DM_OK                 0
#define DM_ERR_INVALID       -1   /* bad argument (shape, window size, method ...) */
#define DM_ERR_CUDA          -2   /* a CUDA runtime / driver call failed          */
#define DM_ERR_NOMEM         -3   /* workspace would exceed the configured limit   */
#define DM_ERR_UNSUPPORTED   -4   /* shape not supported by the requested kernel   */

/* cv2 enum values accepted as `method` (misc/Feature_value.py:24) */
#define DM_TM_CCOEFF          4
#define DM_TM_CCOEFF_NORMED   5

/* disparity planes (misc/Calc_difference.py:30) */
#define DM_MODE_ELEVATION     0   /* j - map[1]              */
#define DM_MODE_ELEVATION2    1   /* i - map[0]              */
#define DM_MODE_DISTANCE      2   /* ||(i,j) - map[:2]||     */

/* correlation engine selector for dm_correlation */
#define DM_CORR_AUTO          0   /* tcgen05 when the shape allows it, else SIMT   */
#define DM_CORR_SIMT          1   /* CUDA-core exact reference kernel              */
#define DM_CORR_UMMA          2   /* tcgen05.mma + TMEM + TMA (errors if unsupported) */
/* test / measurement aids, not part of the drop-in surface:
 *   3  tcgen05 MMAs + TMEM drain only (no epilogue math, nothing written): the MMA/TMA floor
 *   4  tcgen05 with the pooled epilogue of the fused solver; raw_dev then receives
 *      [n][P][P/4] pooled raw ZNCC, followed by [n][P][2] partial row minima and
 *      [n][P][2] partial row maxima (square grids only)                                  */

/* tcgen05 correlation: -1 (default) = run on CTA pairs (tcgen05.mma.cta_group::2, one 2-CTA
 * cluster per TPC) whenever a tile holds an even number of 256-patch work items, 0 = always
 * one CTA per work item.  Both produce bit-identical results; a tuning / test knob. */
int         dm_correlation_set_pair_mode(int mode);

/* a context owns the device workspace and the staging buffers of the scene solvers (dm_ctx_create, below) */
typedef struct dm_ctx dm_ctx;

int         dm_version(void);
const char* dm_last_error(void);
/* compute capability of the current device as major*10+minor, or <0 */
int         dm_device_cc(void);

/* ---------------------------------------------------------------- descriptors ------
 * Replaces Correlation_map._create_atomic_patch (misc/Correlation_map.py:51-67) plus the
 * per-patch statistics cv2.matchTemplate derives internally (misc/Feature_value.py:41).
 * For each of n tiles with top-left corner origin_yx[2*t..] in the scene, and each patch
 * centre (i,j) of the t0 x t1 grid, writes the ws*ws window as bf16 values (pixel minus
 * the patch's rounded mean, exact), zero padded to kpad, plus the window statistics:
 * stat_dev holds DM_STAT_FLOATS (6) floats per patch -- first n*P float4 {S', inv, S'/K,
 * mean} with S' the residual sum and inv = 1/sqrt(sum a'^2 - S'^2/K) (0 for a flat
 * window), then n*P floats inv again (the compact column table of the tcgen05 epilogue; the
 * rest is padding).  kpad = dm_kpad(ws).
 * side: which operand of the correlation the rows are for.  Three unused entries of every row
 * (a window row occupies ws of its 8 / 16 entries) carry the ZNCC correction -S1' S2'/K into the
 * contraction: 1 = patch image (image 1): they hold S'; 2 = search image (image 2): they hold the
 * three bf16 parts of -S'/K; 0 = zeros.  dm_correlation expects desc1 with side 1, desc2 with side 2.
 */
#define DM_STAT_FLOATS 6
int dm_kpad(int ws);
int dm_descriptors(const uint8_t* scene_dev, int scene_h, int scene_w, int pitch,
                   const int32_t* origin_yx_dev, int n_tiles, int t0, int t1, int ws, int side,
                   void* desc_bf16_dev, float* stat_dev, void* stream);

/* ---------------------------------------------------------------- correlation ------
 * Replaces Correlation_map._create_simple_initial_co_map (misc/Correlation_map.py:69-87)
 * = P x cv2.matchTemplate (misc/Feature_value.py:41), without the min-max.
 * raw[t][p][q] = ZNCC(patch p of image 1, window q of image 2) (or the un-normalised
 * TM_CCOEFF numerator), fp32, for n tiles.  p,q in [0, P) with P = t0*t1.
 */
int dm_correlation(const void* desc1_dev, const float* stat1_dev,
                   const void* desc2_dev, const float* stat2_dev,
                   int n_tiles, int p, int kpad, int ws, int method, int engine,
                   float* raw_dev, void* stream);

/* Feature_value.__call__ for arbitrary patch / image sizes (misc/Feature_value.py:39-43,
 * for_igarss/cor_map.py:33-35): float32 (ih-ph+1) x (iw-pw+1) map, min-maxed. */
int dm_feature_value(const uint8_t* patch_dev, int ph, int pw,
                     const uint8_t* image_dev, int ih, int iw,
                     int method, float* out_dev, void* stream);

/* Feature_value.min_max per row (misc/Feature_value.py:32-37) fused with the
 * rectification map**1.4 (misc/Correlation_map.py:158-159).  rows x q fp32.
 * norm_dev and rect_dev may each be NULL (skip) and may alias raw_dev. */
int dm_minmax_rectify(const float* raw_dev, long long rows, int q,
                      float* norm_dev, float* rect_dev,
                      float* rowmin_dev, float* rowmax_dev, void* stream);

/* ---------------------------------------------------------------- pyramid ----------
 * Replaces Correlation_map._aggregation + _rectification for one level transition
 * (misc/Correlation_map.py:89-130,158-159): 3x3/stride-2/pad-1 max-pool of every (C,D)
 * slice, average of the four children of each parent (no shift), then **1.4.
 * in  [n][A][B][C][D]  ->  out [n][A/2][B/2][C/2][D/2]      (A,B,C,D even)
 * rectify = 0 returns the plain average (Correlation_map._aggregation alone).
 */
int dm_aggregate(const float* in_dev, int n, int a, int b, int c, int d, int rectify,
                 float* out_dev, void* stream);

/* ---------------------------------------------------------------- backtracking -----
 * Matching._initial_move_map (misc/Matching.py:80-96): top level [n][a][b][a][b].
 * Matching._B (misc/Matching.py:98-139, filtering off): level [n][A][B][C][D] with the
 * parent matches of the (A/2,B/2) grid.  match = int32 [n][2][A][B] (row, col);
 * score [n][A][B] in the level's dtype.  is_f64 selects double input (bit-exact check
 * against the reference's float64 pyramid).
 */
int dm_backtrack_top(const void* top_dev, int is_f64, int n, int a, int b,
                     int32_t* match_dev, void* score_dev, void* stream);
int dm_backtrack_level(const void* level_dev, int is_f64, int n, int a, int b, int c, int d,
                       const int32_t* parent_match_dev,
                       int32_t* match_dev, void* score_dev, void* stream);

/* bad_matching.py:68-70 -- np.argmax(co_map[i, j, i, :]) for every patch (i, j): the best column of the
 * patch's OWN map row (first maximum, the first NaN wins).  co_map = float32 [n][t0][t1][t0][t1] (min-maxed
 * level 0, Correlation_map.co_map); arg = int32 [n][t0][t1].  rows (may be NULL) = float32 [n][t0][t1][t1]
 * receives the rows themselves, so that a host loop over co_map[i, j, i, :] needs P x T1 values, not P x P. */
int dm_row_argmax(const float* co_map_dev, int n, int t0, int t1, int32_t* arg_dev, float* rows_dev, void* stream);

/* Matching._filter (misc/Matching.py:224-255): outlier filter on the displacement field of a
 * batch of match maps, int32 [n][2][h][w] -> [n][2][h][w] (out must not alias in).  Interior
 * cells become round(mean | median of the (2e+1)^2 neighbourhood of displacements) + their
 * own coordinate, e = (window-1)/2; maps smaller than the window are copied.  The reference
 * is only defined on square maps: h != w returns DM_ERR_UNSUPPORTED. */
#define DM_FILTER_MEDIAN   0
#define DM_FILTER_AVERAGE  1
int dm_match_filter(const int32_t* match_in_dev, int n, int h, int w, int window, int mode,
                    int32_t* match_out_dev, void* stream);

/* Matching._sub_pix_cal (misc/Matching.py:165-209) + assembly of Matching.__call__'s
 * return value (misc/Matching.py:211-222): map_dev = double [n][3][T0][T1] =
 * (row (+diff), col (+diff), score).  sub_pix = 0 skips the parabola fit.
 * is_f64: 0 = float32 level, float32 arithmetic (what numpy does with a float32 co_map_list),
 *         1 = float64 level, 2 = float32 level of the library's own pyramid with the parabola
 *         evaluated in float64 (the reference holds these values in float64 arrays). */
int dm_match_map(const void* level0_dev, int is_f64, int n, int t0, int t1,
                 const int32_t* match_dev, const void* score_dev, int sub_pix,
                 double* map_dev, void* stream);

/* Calc_difference.cal_map (misc/Calc_difference.py:25-49): double (3,T0,T1) -> (T0,T1) */
int dm_cal_map(const double* map_dev, int t0, int t1, int mode, double* out_dev, void* stream);

/* sub_pix_cal (misc/sub_pix_cal.py:22-53) incl. both image_threshold clamps
 * (misc/optimize_loop.py:40-44): double (s0,s1) planes. */
int dm_sub_pix_cal(const double* arr_dev, const double* co_map_dev, int s0, int s1,
                   int direction, double ratio, double* out_dev, void* stream);

/* sub_pix_cal for a batch of planes in HOST memory (config 4 of BASELINE.json: 64 pairs, sub_pix_cal on both
 * planes): d_maps double [n][n_planes][s0][s1], co_maps double [n][s0][s1], directions[n_planes] -> out like d_maps.
 * The batch crosses the device in pieces on two streams (upload, kernels and download of neighbouring pieces
 * overlap); page-locked arrays make the copies asynchronous.  Synchronous: returns when out_host is complete. */
int dm_sub_pix_cal_host_batch(dm_ctx* ctx, const double* d_maps_host, const double* co_maps_host, int n, int n_planes,
                              int s0, int s1, const int32_t* directions, double ratio, double* out_host);

/* cv2.bilateralFilter(plane.astype('uint8'), d, sigma_color, sigma_space) -- the live branch of the
 * reference's post-process (optimize_looper.py:76-77, d = 2*exclusion+1).  uint8 (h,w) planes;
 * OpenCV's own 8-bit algorithm (BORDER_REFLECT_101, float32 weights, cvRound), bit-identical
 * to cv2 with IPP off.  dst must not alias src. */
int dm_bilateral_u8(const uint8_t* src_dev, int h, int w, int d, double sigma_color, double sigma_space,
                    uint8_t* dst_dev, void* stream);

/* ---------------------------------------------------------------- Gauss-Seidel post-process
 * The reference's sequential in-place smoothing loops (its `if 0:` branch, optimize_looper.py:55-74), in the
 * reference's own visiting order: only cells that cannot see each other are computed side by side (anti-diagonals,
 * lockstep columns, skewed diagonals).  float64, bit-identical to the reference given the same inputs.
 * All arrays are double (s0, s1) on the device; diff_dev is scratch of s0*s1 doubles; error_dev receives the
 * sequentially accumulated |old - new| the reference returns.
 *   dm_optimize_loop            misc/optimize_loop.py:15-37 (clamp to [0,10], forward sweep, "reverse" sweep), in place
 *   dm_make_weight              misc/opt_loop.py:66-85: gw (w,w), cw (s0-e, s1-e, w, w), w = 2 e + 1 (exp: CUDA's, <= 1 ulp from numpy's)
 *   dm_optimize_loop_bilateral  misc/opt_loop.py:16-63, horizon (vertical = 0) or vertical (1), in place */
int dm_optimize_loop(double* img_dev, const double* coef_dev, int s0, int s1, int exclusion, double alpha,
                     double* diff_dev, double* error_dev, void* stream);
int dm_make_weight(const double* guide_dev, int s0, int s1, int exclusion, double sigma0, double sigma1,
                   double* gw_dev, double* cw_dev, void* stream);
int dm_optimize_loop_bilateral(double* img_dev, const double* cw_dev, const double* gw_dev, const double* coef_dev,
                               int s0, int s1, int exclusion, int vertical, double* diff_dev, double* error_dev, void* stream);

/* ---------------------------------------------------------------- scene solver -----
 * Replaces ImageCutSolver._cut_and_pool/_solver/_execute_matching
 * (misc/image_cut_solver.py:95-184) for a whole scene or for a strip of tile rows.
 */
typedef struct dm_scene_params {
    int32_t scene_h, scene_w;       /* S0, S1 of both images                              */
    int32_t t0, t1;                 /* image_size (patch grid of a tile)                  */
    int32_t s0, s1;                 /* stride                                             */
    int32_t ws;                     /* window_size (odd)                                  */
    int32_t method;                 /* DM_TM_*                                            */
    int32_t n_modes;                /* number of disparity planes                         */
    int32_t modes[4];               /* DM_MODE_*                                          */
    int32_t sub_pix;                /* Matching(sub_pix=...)                              */
    int32_t tile_row_lo, tile_row_hi; /* strip of tile rows [lo,hi); hi<=0 means all      */
    int32_t fused;                  /* 1: fused tcgen05 path (no level-0 in HBM) when the
                                       shape allows it, 0: materialising path, -1: auto  */
    int32_t n_scenes;               /* batch of equally sized scene pairs stacked along rows
                                       ([n_scenes][S0][S1] inputs, [n_scenes][n_modes][S0'][S1']
                                       and [n_scenes][S0'][S1'] outputs); 0 or 1 = single pair.
                                       Not combinable with a tile-row strip.                */
    int32_t filter_num;             /* Matching(filtering=True, filtering_num=...): number of maps of
                                       the top-down pass that go through Matching._filter
                                       (misc/Matching.py:91-93,136-138); 0 = filtering off    */
    int32_t filter_cfg;             /* filter_window_size | (DM_FILTER_* << 8)               */
    int32_t tile_lo, tile_hi;       /* range of tiles [lo,hi) in row-major order (tile = gi * len1 + gj); hi <= 0
                                       means "not used".  Finer than a strip of tile rows: the shares of several
                                       devices differ by at most one TILE.  The pixels a range owns are up to three
                                       rectangles of the mosaic (a partial tile row, whole rows, a partial tile row).
                                       Not combinable with tile_row_lo/hi or a batch of scenes.                  */
} dm_scene_params;

typedef struct dm_scene_info {
    int32_t len0, len1;             /* tile grid (misc/image_cut_solver.py:62)            */
    int32_t out_h, out_w;           /* S0', S1' of the full mosaic                        */
    int32_t row_lo, row_hi;         /* output rows owned by the strip (tile range: the rows it touches) */
    int32_t n_tiles;                /* tiles in the strip (all scenes of a batch)         */
    int32_t levels;                 /* pyramid depth ("iteration")                        */
    int32_t n_map;                  /* N_map                                              */
    int32_t used_fused;             /* which path ran                                     */
    int32_t chunk_tiles;            /* tiles per batch                                    */
    int32_t kernel_launches;        /* kernels enqueued by the last solve                 */
} dm_scene_info;

int  dm_ctx_create(dm_ctx** out);                 /* on the current CUDA device          */
void dm_ctx_destroy(dm_ctx* ctx);
int  dm_ctx_set_stream(dm_ctx* ctx, void* stream);
int  dm_ctx_set_workspace_limit(dm_ctx* ctx, size_t bytes);
size_t dm_ctx_workspace_bytes(const dm_ctx* ctx);

/* geometry only (no GPU work): fills info for the given params */
int dm_scene_geometry(const dm_scene_params* prm, dm_scene_info* info);

/* geometry only: the (at most three) rectangles of the mosaic the tiles of `prm` own -- the tail of the first
 * tile row of the range, whole tile rows, the head of its last tile row; rects[k] = {row_lo, row_hi, col_lo, col_hi}.
 * (A pixel belongs to the covering tile with the largest index, misc/image_cut_solver.py:165-175.) */
int dm_owned_rectangles(const dm_scene_params* prm, int32_t* rects /* [3][4] */, int32_t* n_rects);

/* Device-resident solve.  img*_dev: uint8 [scene_h][scene_w].  d_map_dev: double
 * [n_modes][out_h][out_w], out_map_dev: double [out_h][out_w]; only rows
 * [row_lo,row_hi) are written.  Asynchronous on the ctx stream. */
int dm_solve_scene(dm_ctx* ctx, const dm_scene_params* prm,
                   const uint8_t* img1_dev, const uint8_t* img2_dev,
                   double* d_map_dev, double* out_map_dev, dm_scene_info* info);

/* Host-buffer solve: copies both scenes host->device, solves, copies the owned rows of
 * the planes device->host, and synchronises the ctx stream before returning. */
int dm_solve_scene_host(dm_ctx* ctx, const dm_scene_params* prm,
                        const uint8_t* img1_host, const uint8_t* img2_host,
                        double* d_map_host, double* out_map_host, dm_scene_info* info);

/* Device-resident solve that ALSO streams its finished output rows, band by band behind the final
 * stage, into a second pair of arrays: page-locked host memory, the mosaic of a peer device of this
 * process, or a mosaic another process exported with dm_ipc_export (any UVA pointer the device can
 * write).  This is how the strips of several devices meet in one mosaic without a collective after
 * the solve: the copy of band k rides NVLink / PCIe while band k+1 is being solved.  The ctx stream
 * is made to wait for the last copy: synchronise on it (or on an event recorded on it). */
int dm_solve_scene_stream(dm_ctx* ctx, const dm_scene_params* prm,
                          const uint8_t* img1_dev, const uint8_t* img2_dev,
                          double* d_map_dev, double* out_map_dev,
                          double* d_map_dst, double* out_map_dst, dm_scene_info* info);

/* ---------------------------------------------------------------- several devices ---
 * The tile loop of ImageCutSolver._execute_matching (misc/image_cut_solver.py:144-184) spread over
 * the devices of one process: tile rows are cut into contiguous strips (sizes differing by at most
 * one, low ranks first: 66 -> 9,9,8,...), device r solves strip r with its own dm_ctx, stream and
 * host thread.  Tiles are independent and every output row is owned by exactly one tile row, so the
 * result is bit-identical to the one-device solve.
 */
typedef struct dm_multi dm_multi;

/* devices == NULL: devices 0 .. n_devices-1; n_devices <= 0: every visible device */
int  dm_multi_create(const int* devices, int n_devices, dm_multi** out);
void dm_multi_destroy(dm_multi* m);
int  dm_multi_device_count(const dm_multi* m);
int  dm_multi_set_workspace_limit(dm_multi* m, size_t bytes_per_device);
/* contiguous tile-row strips [lo[r], hi[r]) of len0 tile rows over n parts */
int  dm_partition_tile_rows(int len0, int n, int32_t* lo, int32_t* hi);

/* Host-buffer solve over all devices: device r uploads the input rows of its strip, solves it and
 * streams its finished rows straight into the caller's (ideally page-locked) arrays.  No collective:
 * the host arrays are the meeting point.  max_devices > 0 caps the devices used (0 = all).
 * info (may be NULL) describes the whole scene; kernel_launches is summed over the devices. */
int dm_multi_solve_scene_host(dm_multi* m, const dm_scene_params* prm, int max_devices,
                              const uint8_t* img1_host, const uint8_t* img2_host,
                              double* d_map_host, double* out_map_host, dm_scene_info* info);

/* Device-resident solve over all devices + gather of the finished strips on device `root`.
 * img1_dev[r], img2_dev[r]: the scene on device r (only the input rows of strip r are read);
 * planes_dev[r]: double [n_modes + 1][out_h][out_w] on device r (disparity planes, then the score
 * plane).  After the call (asynchronous: dm_multi_synchronize) planes_dev[root] holds the whole
 * mosaic.  gather: DM_GATHER_P2P  = every device streams its finished bands into the root's planes
 *                                   over NVLink peer memory while it is still solving;
 *                  DM_GATHER_NCCL = one grouped ncclSend / ncclRecv of the owned rows after the
 *                                   solve (libnccl.so.2 is loaded on first use; ncclCommInitAll). */
#define DM_GATHER_P2P   0
#define DM_GATHER_NCCL  1
int dm_multi_solve_scene(dm_multi* m, const dm_scene_params* prm,
                         const uint8_t* const* img1_dev, const uint8_t* const* img2_dev,
                         double* const* planes_dev, int root, int gather, dm_scene_info* info);
/* the gather alone (NCCL): rows [row_lo[r], row_hi[r]) of every plane of planes_dev[r] -> planes_dev[root] */
int dm_multi_gather_strips(dm_multi* m, double* const* planes_dev, int n_planes, int out_h, int out_w,
                           const int32_t* row_lo, const int32_t* row_hi, int root);
int dm_multi_synchronize(dm_multi* m);

/* ---------------------------------------------------------------- shared mosaics ----
 * One process per device (torchrun): the root allocates the mosaic with dm_ipc_alloc, exports a
 * 64-byte handle, the other processes open it and pass the pointer to dm_solve_scene_stream. */
#define DM_IPC_HANDLE_BYTES 64
int dm_ipc_alloc(size_t bytes, void** dev_ptr);                 /* cudaMalloc on the current device */
int dm_ipc_free(void* dev_ptr);
int dm_ipc_export(void* dev_ptr, unsigned char* handle /* [DM_IPC_HANDLE_BYTES] */);
int dm_ipc_open(const unsigned char* handle, void** dev_ptr);   /* maps the exporter's memory; peer access is enabled lazily */
int dm_ipc_close(void* dev_ptr);

/* per-stage device time of the last dm_solve_scene* call in milliseconds (CUDA events on
 * the ctx stream; enabled by dm_ctx_enable_timing).  Stage ids: DM_STAGE_*. */
#define DM_STAGE_DESCRIPTORS  0
#define DM_STAGE_CORRELATION  1
#define DM_STAGE_NORMALIZE    2
#define DM_STAGE_AGGREGATE    3
#define DM_STAGE_BACKTRACK    4
#define DM_STAGE_PLANES       5
#define DM_STAGE_COUNT        6
int dm_ctx_enable_timing(dm_ctx* ctx, int on);
int dm_ctx_stage_ms(dm_ctx* ctx, float* ms_out /* [DM_STAGE_COUNT] */,
                    int* launches_out /* [DM_STAGE_COUNT] */);

#ifdef __cplusplus
}
#endif
#endif /* DMSTEREO_H_ */
