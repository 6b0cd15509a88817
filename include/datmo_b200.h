/*
 * datmo_b200 — C ABI of the B200-native DATMO optical-flow hot path.
 *
 * The reference (anvithaanchala/DATMO_using_Optical_flow) is pure Python and has
 * no FFI layer of its own: the boundary a maintainer binds is the set of stage
 * functions that Optical_flow/main.py:process_multiple_frames calls.  Each
 * entry point below names the reference function (file:line) it replaces.
 * The ctypes binding that ships with this repo is
 * datmo_using_optical_flow_b200/_lib.py; INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *  - every function returns an int status: 0 = ok, negative = error (see
 *    DATMO_E_*); nothing throws, nothing aborts.  datmo_last_error() returns a
 *    human-readable message for the last failing call on that handle.
 *  - a handle is bound to one CUDA device and one stream; calls on one handle
 *    must not overlap; different handles are independent (one per thread / rank).
 *  - "_dev" entry points take DEVICE pointers, enqueue work on the handle's
 *    stream and return without synchronising (unless stated);
 *    "_host" entry points take HOST pointers, copy in, run, copy out and
 *    synchronise before returning.
 *  - there is no CPU fallback anywhere in this library.
 *  - images are row-major (H rows, W columns), batches are contiguous [B][H][W].
 */
#ifndef DATMO_B200_H
#define DATMO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DATMO_ABI_VERSION 2

#define DATMO_OK 0
#define DATMO_E_INVALID -1   /* bad argument */
#define DATMO_E_CUDA -2      /* a CUDA call failed; see datmo_last_error */
#define DATMO_E_CAPACITY -3  /* an output buffer was too small; counts are still written */
#define DATMO_E_EMPTY -4     /* nothing to do (e.g. no point inside the ROI) */

typedef struct datmo_ctx* datmo_handle_t;

/* image element types for the Farneback inputs */
#define DATMO_U8 0
#define DATMO_F32 1
#define DATMO_F64 2   /* propagation masks only */
/* point layouts */
#define DATMO_PTS_F64_XYZ 0  /* double[n][3], what main.py passes around */
#define DATMO_PTS_F32_XYZW 1 /* float[n][4], CARLA's native layout */

/* ---- lifecycle ------------------------------------------------------------- */
/* stream: a cudaStream_t to run on (borrowed), or NULL to create a private
 * non-blocking stream. */
int datmo_create(int device, void* stream, datmo_handle_t* out);
int datmo_destroy(datmo_handle_t h);
int datmo_abi_version(void);
const char* datmo_last_error(datmo_handle_t h);
int datmo_synchronize(datmo_handle_t h);
/* bytes of device workspace currently held by the handle */
size_t datmo_workspace_bytes(datmo_handle_t h);

/* ---- per-kernel timing (bench.py's roofline leg) -------------------------------
 * When enabled, the handle brackets every launch of its tagged kernels with CUDA
 * events on its own stream.  datmo_profile_read synchronises and returns, per tag,
 * the number of launches and the summed device time in milliseconds since the last
 * reset.  Tags: see DATMO_TAG_*.  */
#define DATMO_TAG_PYRAMID 0
#define DATMO_TAG_POLYEXP 1
#define DATMO_TAG_FLOW_INIT 2
#define DATMO_TAG_FLOW_ITER 3   /* the fused updateMatrices + blur + solve kernel */
#define DATMO_TAG_VELMASK 4
#define DATMO_TAG_DBSCAN 5
#define DATMO_TAG_BEV 6
#define DATMO_TAG_RANSAC 7
#define DATMO_TAG_CLUSTER 8
#define DATMO_TAG_COUNT 9
int datmo_profile_enable(datmo_handle_t h, int on);
/* time only the tags whose bit (1u << DATMO_TAG_x) is set in mask (0 = off): the event pairs cost
 * a few microseconds per launch, so a benchmark brackets just the kernel it reports on */
int datmo_profile_tags(datmo_handle_t h, unsigned mask);
int datmo_profile_reset(datmo_handle_t h);
int datmo_profile_read(datmo_handle_t h, int64_t launches[DATMO_TAG_COUNT], double ms[DATMO_TAG_COUNT]);
/* total kernel launches issued by this handle since creation */
int64_t datmo_launch_count(datmo_handle_t h);

/* ---- Farneback dense optical flow ----------------------------------------------
 * Replaces cv2.calcOpticalFlowFarneback as called by compute_velocity_vectors,
 * Optical_flow/main.py:131-142 (parameters hard-coded at main.py:132-140:
 * pyr_scale 0.3, levels 5, winsize 15, iterations 5, poly_n 5, poly_sigma 5,
 * flags 0).  flags must be 0 (the only value on the reference path). */
typedef struct datmo_farneback_params {
    double pyr_scale;
    int levels;
    int winsize;
    int iterations;
    int poly_n;
    double poly_sigma;
    int flags;
    int variant; /* 0 = default; 1 = unfused updateMatrices / blur+solve (debug / A-B) */
} datmo_farneback_params;
void datmo_farneback_default_params(datmo_farneback_params* p);
/* number of pyramid layers actually processed and their sizes (coarsest first);
 * returns the layer count, fills up to max_layers entries of w[] and h[]. */
int datmo_farneback_layers(int H, int W, const datmo_farneback_params* p, int max_layers, int* w, int* h);
/* prev/next: [batch][H][W] of dtype; flow: float [batch][H][W][2] (dx, dy). */
int datmo_farneback_dev(datmo_handle_t h, const void* prev, const void* next, int dtype, int H, int W,
                        int batch, const datmo_farneback_params* p, float* flow);
int datmo_farneback_host(datmo_handle_t h, const void* prev, const void* next, int dtype, int H, int W,
                         int batch, const datmo_farneback_params* p, float* flow);

/* stage-level entry points (device pointers), used by the parity tests to diff
 * each step against the oracle.  R and M arrays cross the ABI as planar float [batch][5][h][w]
 * (internally R is kept as float4 + float per pixel). */
int datmo_fb_pyramid_image_dev(datmo_handle_t h, const void* img, int dtype, int H, int W, int batch,
                               int ksize, double sigma, int h_out, int w_out, float* out);
int datmo_fb_polyexp_dev(datmo_handle_t h, const float* img, int hh, int ww, int batch, int poly_n,
                         double poly_sigma, float* R);
int datmo_fb_update_matrices_dev(datmo_handle_t h, const float* R0, const float* R1, const float* flow,
                                 int hh, int ww, int batch, float* M);
int datmo_fb_blur_solve_dev(datmo_handle_t h, const float* M, int hh, int ww, int batch, int winsize,
                            float* flow);
int datmo_fb_flow_iter_dev(datmo_handle_t h, const float* R0, const float* R1, const float* flow_in,
                           int hh, int ww, int batch, int winsize, float* flow_out);
int datmo_fb_upsample_flow_dev(datmo_handle_t h, const float* flow_in, int h_in, int w_in, int batch,
                               int h_out, int w_out, double mul, float* flow_out);

/* ---- velocity grid, continuity mask, moving-cell filter ------------------------
 * Replaces the tail of compute_velocity_vectors (main.py:143-164: velocity =
 * flow * pixel size, curl), continuity_mask (main.py:224-228) and the inline
 * filter of process_multiple_frames (main.py:596-609).
 * flow: float [batch][H][W][2].  Outputs (any may be NULL), all [batch][H][W]:
 *   vx, vy      float   velocity (m per frame; the reference ignores dt)
 *   ang         float   curl of the unfiltered velocity (main.py:160)
 *   mask        uint8   continuity mask 0/1
 *   vx_f, vy_f  float   velocity * mask
 *   ang_f       float   curl of the filtered field (main.py:604-606), as float
 *   valid       uint8   sqrt(vx_f^2 + vy_f^2) > thresh, evaluated in fp64
 *   n_valid     int32 [batch]  number of valid cells per frame pair */
int datmo_velocity_mask_dev(datmo_handle_t h, const float* flow, int H, int W, int batch, double px_x,
                            double px_y, double alpha_cont, double thresh, float* vx, float* vy,
                            float* ang, uint8_t* mask, float* vx_f, float* vy_f, float* ang_f,
                            uint8_t* valid, int32_t* n_valid);

/* The filtered field in the dtype the reference holds it in (main.py:600-606: f32 * int64 mask -> float64,
 * its magnitude and the np.gradient curl of the f64 arrays) — what save_velocity_grid and the per-cell CSV
 * receive.  vx_f, vy_f: float [batch][H][W] from datmo_velocity_mask_dev; outputs double [batch][H][W],
 * any may be NULL. */
int datmo_filtered_grids_f64_dev(datmo_handle_t h, const float* vx_f, const float* vy_f, int H, int W,
                                 int batch, double* vx64, double* vy64, double* mag64, double* ang64);
/* double -> float for the velocities dbscan_clustering receives (main.py:231: f32 values held in f64);
 * *lossy (host) is set when a value is not float32-representable.  Synchronises. */
int datmo_narrow_f64_dev(datmo_handle_t h, const double* src, int64_t n, float* dst, int* lossy);

/* ---- propagation masks ---------------------------------------------------------
 * Replaces propagation_mask (main.py:166-182) and propagation_mask_with_acceleration
 * (main.py:184-221; pass ax = ay = NULL for the former).  The reference defines both
 * and its driver never calls them (main.py:596-597 is commented out); they are here
 * because its README lists them as part of the method.
 * vx, vy (ax, ay): [batch][H][W] of dtype DATMO_F32 or DATMO_F64 — the dtype the
 * reference would compute in.  Every cell's velocity is scattered to
 * (i + floor(vx dt / grid_x), j + floor(vy dt / grid_y)); among several sources of one
 * target the LAST in row-major order wins (the reference's double loop); mask =
 * |scattered - actual| <= alpha_p on both components, uint8 0/1 [batch][H][W].
 * NaN / inf displacements (the reference raises on them) do not propagate. */
int datmo_propagation_mask_dev(datmo_handle_t h, const void* vx, const void* vy, const void* ax,
                               const void* ay, int dtype, int H, int W, int batch, double dt,
                               double grid_x, double grid_y, double alpha_p, uint8_t* mask);

/* ---- DBSCAN over (row, col, vx, vy) of the valid cells --------------------------
 * Replaces dbscan_clustering, main.py:231-259 (sklearn.cluster.DBSCAN on the
 * row-major list of valid cells).  Labels are identical to sklearn's, including
 * numbering.  vx_f, vy_f: float [batch][H][W]; valid: uint8 [batch][H][W].
 * Outputs: n_valid int32[batch]; labels int32 [batch][cap]; indices int32
 * [batch][cap][2] (row, col) in row-major order of the valid cells; n_clusters
 * int32[batch] (may be NULL).  If any frame has more than cap valid cells the
 * call returns DATMO_E_CAPACITY after writing n_valid (host entry point only;
 * the device entry point truncates and the caller checks n_valid). */
int datmo_dbscan_grid_dev(datmo_handle_t h, const float* vx_f, const float* vy_f, const uint8_t* valid,
                          int H, int W, int batch, double eps, int min_samples, int cap,
                          int32_t* n_valid, int32_t* labels, int32_t* indices, int32_t* n_clusters);

/* ---- cluster summaries ------------------------------------------------------------
 * Replaces extract_cluster_data, main.py:402-434.  For each frame and each label
 * l < max_clusters writes 8 doubles: count, mean row, mean col, mean vx, mean vy,
 * cov(row,row), cov(row,col), cov(col,col) (ddof 1, NaN when count < 2);
 * the eigenvalues follow on the host from the 2x2 covariance.
 * summary: double [batch][max_clusters][8]. */
int datmo_cluster_summary_dev(datmo_handle_t h, const float* vx_f, const float* vy_f, int H, int W,
                              int batch, int cap, const int32_t* n_valid, const int32_t* labels,
                              const int32_t* indices, int max_clusters, double* summary);

/* ---- compact cell records for the host -------------------------------------------
 * (row, col) of the first min(n_valid, cap) cells of every frame packed as
 * (row << 16) | col into one uint32 — a third less to move over PCIe than the int32
 * pairs when the host reads labels and indices of every moving cell (H, W <= 65535).
 * indices: int32 [batch][cap][2] as written by datmo_dbscan_grid_dev;
 * packed: uint32 [batch][cap]. */
int datmo_pack_indices_dev(datmo_handle_t h, const int32_t* indices, const int32_t* n_valid, int cap,
                           int batch, uint32_t* packed);

/* ---- BEV rasterisation -------------------------------------------------------------
 * Replaces compute_bev_grid, main.py:98-126.  pts: n points in the given layout;
 * nx, ny = len(np.arange(lo, hi, step)) (datmo_bev_bins).  bev: uint8 [nx][ny],
 * axis 0 = x.  Bit-exact with the reference for finite inputs. */
int datmo_bev_bins(double lo, double hi, double step);
int datmo_bev_rasterize_dev(datmo_handle_t h, const void* pts, int layout, int64_t n, double res_x,
                            double res_y, double x_lo, double y_lo, int nx, int ny, double a, double b,
                            double h_max, uint8_t* bev);
int datmo_bev_rasterize_host(datmo_handle_t h, const void* pts, int layout, int64_t n, double res_x,
                             double res_y, double x_lo, double y_lo, int nx, int ny, double a, double b,
                             double h_max, uint8_t* bev);

/* ---- ROI crop and density expansion ---------------------------------------------------------
 * datmo_roi_filter_dev replaces filter_points_in_roi, main.py:30-36: keeps, in order, the points
 * with x, y, z inside the CLOSED intervals roi = {x_min, x_max, y_min, y_max, z_min, z_max}.
 * out: room for n points in the same layout; *n_out (host) receives the count.  Synchronises.
 * datmo_expand_points_dev replaces increase_point_density, main.py:38-57: `expansion`
 * consecutive copies of every point plus noise (double [n*expansion][3], or NULL to draw
 * N(0, noise_std) from seed).  pts / out: double [n][3] / [n*expansion][3]. */
int datmo_roi_filter_dev(datmo_handle_t h, const void* pts, int layout, int64_t n, const double roi[6], void* out,
                         int64_t* n_out);
int datmo_expand_points_dev(datmo_handle_t h, const double* pts, int64_t n, int expansion, double noise_std,
                            const double* noise, uint64_t seed, double* out);

/* ---- RANSAC ground plane -------------------------------------------------------------
 * Replaces flipped_pcd.segment_plane(0.5, 5, 5000), main.py:73 (Open3D).  Scores
 * num_iterations plane hypotheses (ransac_n samples each, drawn by the counter-based
 * hash documented in oracle/ransac_np.py) against all n points in fp64.
 * Outputs (device): plane double[4] = the winning hypothesis; refit double[4] = plane
 * refit on its inliers; inlier_mask uint8[n]; best int32[2] = {hypothesis index,
 * inlier count}.  hyp_planes (double[num_iterations][4]), hyp_count (int32[..]) and
 * hyp_err (double[..]) may be NULL; when given they receive every hypothesis. */
int datmo_ransac_ground_dev(datmo_handle_t h, const void* pts, int layout, int64_t n, int flip_x,
                            double distance_threshold, int ransac_n, int num_iterations, uint64_t seed,
                            double* plane, double* refit, uint8_t* inlier_mask, int32_t* best,
                            double* hyp_planes, int32_t* hyp_count, double* hyp_err);

/* ---- fused per-frame preprocessing ---------------------------------------------------
 * Replaces preprocess_pcd after the file read, main.py:65-92: flip x, RANSAC ground
 * removal, ROI crop (closed intervals, main.py:30-36), x expansion copies with Gaussian
 * noise (main.py:38-57) and rasterisation.  pts: float [n][4] (device).  noise: double
 * [n][expansion][3] (device) added to the copies of point i — or NULL to draw
 * N(0, noise_std) on the device from seed.  ground_mask: optional uint8[n] (device);
 * when given, RANSAC is skipped and these points are dropped instead.
 * n_roi (host int64*, may be NULL) receives the number of points inside the ROI;
 * returns DATMO_E_EMPTY when it is zero (the reference returns None, main.py:84-86).
 * Synchronises. */
int datmo_preprocess_dev(datmo_handle_t h, const float* pts, int64_t n, int flip_x,
                         double distance_threshold, int ransac_n, int num_iterations, uint64_t seed,
                         const uint8_t* ground_mask, const double roi[6], int expansion,
                         double noise_std, const double* noise, double res_x, double res_y, double x_lo,
                         double y_lo, int nx, int ny, double h_max, uint8_t* bev, int64_t* n_roi);

/* ---- flow -> clusters for batches of frame pairs in HOST memory ------------------------
 * The body of process_multiple_frames between the two BEVs and the EKF, main.py:577-615:
 * compute_velocity_vectors (main.py:131-164), continuity_mask (main.py:224-228), the inline
 * moving-cell filter (main.py:596-609), dbscan_clustering (main.py:231-259) and
 * extract_cluster_data (main.py:402-434), for `batch` frame pairs per submission.
 *
 * A chain is a pipelined object on top of a handle: datmo_chain_submit enqueues the
 * host-to-device copy of one batch (on a copy stream), the kernels (on the handle's stream) and a
 * gather of the ragged per-pair results into one contiguous buffer per array, and returns without
 * blocking; datmo_chain_collect waits for the batch and reads it back with ONE device-to-host copy
 * per array (labels, cells, summaries), sized from the batch's own counts.  With n_slots >= 2 the
 * copies of one batch overlap the kernels of the next.  Frames in pinned (page-locked) host memory
 * make the uploads asynchronous; pageable memory works and is staged by the driver. */
typedef struct datmo_chain_config {
    int H, W, batch;         /* frame geometry; pairs per submission */
    int dtype;               /* DATMO_U8 or DATMO_F32 frames */
    double px_x, px_y;       /* metres per pixel: range / shape, main.py:147-150 */
    double alpha_cont;       /* config.yaml masks.alpha_cont[0] */
    double thresh;           /* moving-cell threshold, 0.1 at main.py:609 */
    double eps;              /* config.yaml dbscan_params */
    int min_samples;
    int cap;                 /* moving cells kept per pair (row-major prefix; counts are exact) */
    int max_clusters;        /* summary rows per pair; 0 = no summaries */
    int want_cells;          /* 1: label + cell index of every moving cell come back; 0: counts, summaries */
    int n_slots;             /* submissions that may be in flight, 1..16 */
    datmo_farneback_params fb;
} datmo_chain_config;

/* Results of one submission.  Pointers are into pinned host memory owned by the chain and stay valid
 * until the slot is submitted again.  The cells of pair b are entries offsets[b] .. offsets[b+1]-1 of
 * labels / cells, in row-major order of the moving cells (np.nonzero order). */
typedef struct datmo_chain_result {
    const int32_t* n_valid;     /* [batch] moving cells per pair (may exceed cap) */
    const int32_t* n_clusters;  /* [batch] */
    const int64_t* offsets;     /* [batch + 1] */
    const void* labels;         /* int16 when label_bytes == 2 (every pair has < 32768 clusters), else int32; -1 = noise */
    int label_bytes;
    const uint32_t* cells;      /* (row << 16) | col */
    const double* summary;      /* [batch][summary_rows][8]: count, mean row, mean col, mean vx, mean vy, cov rr, rc, cc */
    int summary_rows;           /* min(max over pairs of n_clusters, max_clusters) */
    int truncated;              /* some pair had more than cap moving cells */
    int64_t h2d_bytes, d2h_bytes; /* bytes this submission moved over the bus */
} datmo_chain_result;

typedef struct datmo_chain* datmo_chain_t;
/* reference defaults: alpha_cont 0.2, thresh 0.1, eps 5, min_samples 3, Farneback of main.py:132-140;
 * H, W, px_x, px_y, cap must be set by the caller */
void datmo_chain_default_config(datmo_chain_config* cfg);
int datmo_chain_create(datmo_handle_t h, const datmo_chain_config* cfg, datmo_chain_t* out);
int datmo_chain_destroy(datmo_chain_t c);
const char* datmo_chain_last_error(datmo_chain_t c);
/* prev / next: [batch][H][W] of cfg.dtype in host memory; must stay untouched until the slot is collected */
int datmo_chain_submit(datmo_chain_t c, int slot, const void* prev_host, const void* next_host);
int datmo_chain_collect(datmo_chain_t c, int slot, datmo_chain_result* out);
/* One-shot form with caller-owned outputs (the chain is cached on the handle between calls with the
 * same configuration; cfg->n_slots and cfg->want_cells are ignored).  labels int32 [capacity_cells],
 * indices int32 [capacity_cells][2] (row, col) — either may be NULL; summary double
 * [batch][max_clusters][8] or NULL; offsets int64 [batch + 1].  Returns DATMO_E_CAPACITY when the
 * cells do not fit capacity_cells or a pair exceeded cfg->cap (counts and offsets are still written). */
int datmo_flow_to_clusters_host(datmo_handle_t h, const void* prev, const void* next, const datmo_chain_config* cfg,
                                int32_t* n_valid, int32_t* n_clusters, int64_t* offsets, int32_t* labels,
                                int32_t* indices, int64_t capacity_cells, double* summary);

#ifdef __cplusplus
}
#endif
#endif /* DATMO_B200_H */
