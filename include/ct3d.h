/*
 * ct3d.h -- C ABI of libct3d.so: the B200 (sm_100a) hot path of a 3DeeCellTracker-compatible
 * segment-and-track engine.
 *
 * The reference (WenChentao/3DeeCellTracker) is pure Python; it has no FFI layer.  The seams this
 * library plugs into are the reference's Python operator functions; every entry point below names the
 * reference function (file:line under CellTracker/) whose arithmetic it replaces.  The Python package
 * `3deecelltracker_b200` binds these symbols with ctypes and keeps the reference's signatures.
 *
 * Conventions
 *   - plain C types only; every data pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the caller owns every buffer, including workspaces (query with ct_*_workspace_bytes);
 *   - all work is enqueued on the CUDA stream passed as `stream` (a cudaStream_t cast to void*);
 *     no entry point synchronises the device unless documented;
 *   - return value 0 = success, non-zero = error; ct_last_error() returns a thread-local message;
 *   - the library never falls back to a CPU implementation.
 *   - volumes are C-ordered (x, y, z) arrays (z fastest), as in the reference (unet3d.py:210).
 */
#ifndef CT3D_H_
#define CT3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CT3D_ABI_VERSION 1

int ct_abi_version(void);
const char* ct_last_error(void);
/* Number of kernels this library has launched in the calling process (for bench.py "gpu_launches"). */
unsigned long long ct_launch_count(void);
/* Opt-in profiler used by bench.py: CUDA events are recorded on the launching stream around every launch of a
 * kernel family.  tag: 1 = U-Net convolutions, 2 = PR-GLS EM, 3 = FFN match, 4 = LCN, 5 = U-Net pool/upsample/
 * gather/head, 6 = watershed stage.  ct_profile_read synchronises on the recorded events and returns their summed duration. */
/* Persistent kernels of this library (the tcgen05 convolution: one CTA per SM) leave `n` SMs unclaimed, so that a
 * single-CTA kernel running concurrently on another stream (the PR-GLS EM of the previous frame, tracker.py's frame
 * pipeline) finds a free SM instead of delaying one CTA of every convolution.  The setting belongs to the CALLING HOST
 * THREAD (it applies to the launches that thread makes afterwards), so pipelines driven from different threads do not
 * disturb each other.  Default 0; returns the thread's old value. */
int ct_set_reserved_sms(int n);
int ct_profile_enable(int on);
int ct_profile_read(int tag, double* total_ms, unsigned long long* count, int reset);

/* ------------------------------------------------------------------------------------------------
 * LCN normalisation.  Replaces preprocess.py:170-188 (_normalize_image) + :136-167 (lcn_gpu) +
 * :117-133 (conv3d_keras): median subtract, clamp at 0, 27x27x1 zero-padded box mean / std,
 * (x-avg)/(std+noise_level).
 * dtype: 0 = uint16, 1 = float32, 2 = uint8.   out: float32 (x,y,z).
 * ---------------------------------------------------------------------------------------------- */
size_t ct_normalize_workspace_bytes(int x, int y, int z);
/* subtract_median = 1: _normalize_image (median subtract + clamp + LCN); 0: lcn_gpu alone (no clamp). */
int ct_normalize_image(const void* raw, int dtype, float* out, int x, int y, int z, float noise_level,
                       int filter_x, int filter_y, int subtract_median, void* ws, size_t ws_bytes,
                       void* stream);
/* The median alone (what np.median returns; mean of the two middle values for even counts).
 * Writes one double to median_out (device). */
int ct_median(const void* raw, int dtype, long long count, double* median_out, void* ws, size_t ws_bytes,
              void* stream);

/* The same median when the voxels of one volume are spread over several GPUs (config 3: a 1024x1024x96 stack cut
 * 2x2x2).  np.median (preprocess.py:181) is a GLOBAL order statistic, so the radix select runs in
 * lock-step on every rank: per pass each rank histograms one 8-bit digit of ITS voxels into `state`, the caller
 * sums the 512 histogram words at byte offset ct_select_hist_offset() of `state` over the ranks (ncclAllReduce on
 * uint32), and the scan narrows the key prefix identically everywhere.  passes: uint8 1, uint16 2, float32 4.
 *     ct_select_begin(state, total_count); for pass in 0 .. ct_select_passes(dtype)-1:
 *         ct_select_hist(...); <all-reduce the histogram>; ct_select_scan(...);
 *     ct_select_finish(state, dtype, median_out)
 * With one rank (no all-reduce) the sequence is exactly ct_median. */
size_t ct_select_state_bytes(void);
size_t ct_select_hist_offset(void);
int ct_select_passes(int dtype);
int ct_select_begin(void* state, long long total_count, void* stream);
int ct_select_hist(const void* raw, int dtype, long long local_count, void* state, int pass, void* stream);
int ct_select_scan(void* state, int dtype, int pass, void* stream);
int ct_select_finish(const void* state, int dtype, double* median_out, void* stream);
/* ct_normalize_image with the median supplied by the caller (1 double, device): LCN of a BLOCK of a larger
 * volume.  Zero padding applies at the block's own faces, so output voxels closer than 2 * (filter/2) to a face
 * that is not a face of the whole volume (the std window reads avg, which reads its own window) are not the
 * volume's values -- the caller passes a block with that margin
 * (halo exchange) and discards it.  Voxels whose whole two-level window lies inside the block are bit-identical to the
 * single-call result (per-voxel windows are summed in a fixed order, not by running sums). */
int ct_normalize_image_with_median(const void* raw, int dtype, float* out, int x, int y, int z, float noise_level,
                                   int filter_x, int filter_y, const double* median, void* ws, size_t ws_bytes,
                                   void* stream);

/* ------------------------------------------------------------------------------------------------
 * 3D U-Net.  Replaces the Keras graphs of unet3d.py:26-98 (unet3_a/b/c), the blocks :101-200 and the
 * tiled prediction unet3_prediction (unet3d.py:203-256).
 * ---------------------------------------------------------------------------------------------- */
#define CT_UNET_MAX_LEVELS 4
typedef struct CtUNetSpec {
    int in_x, in_y, in_z;              /* model input tile, e.g. 160,160,16 (unet3d.py:36) */
    int pool_x, pool_y, pool_z;        /* MaxPooling3D / UpSampling3D size (unet3d.py:35) */
    int act_relu;                      /* 0: Conv->LeakyReLU(0.3)->BN (:117-119); 1: Conv(relu)->BN (:139-140) */
    int levels;                        /* number of _downscale blocks */
    int down[CT_UNET_MAX_LEVELS][2];   /* filters of the two convs of each _downscale (:88-90) */
    int up[CT_UNET_MAX_LEVELS][2];     /* filters of the two convs of each _upscale, deepest first (:91-93) */
    int out[2];                        /* output_m2 / output_m1 filters (:94-95) */
} CtUNetSpec;

typedef struct CtUNet CtUNet;

/* weights_host: float32 arrays concatenated in Keras Model.get_weights() order: per Conv3D+BN block
 * kernel(3,3,3,Cin,Cout), bias, gamma, beta, moving_mean, moving_var; head kernel(1,1,1,C,1), bias. */
size_t ct_unet_weight_count(const CtUNetSpec* spec);
int ct_unet_create(const CtUNetSpec* spec, const float* weights_host, size_t n_floats, CtUNet** out);
void ct_unet_destroy(CtUNet* net);
/* engine: 0 = auto, 1 = CUDA-core fp32 direct convolution, 2 = tcgen05 implicit GEMM (fp16 hi/lo operand split, fp32
 * accumulate; the x-stacked kernel for Cout 8/16, the 27-tap kernel for Cout 32/64; first block Cin = 1 on CUDA cores
 * fused with the tile gather), 3 = tcgen05 27-tap kernel only (no stacked kernel, no fused first block),
 * 4 = x-stacked kernel wherever it supports the layer (today identical to 2). */
int ct_unet_set_engine(CtUNet* net, int engine);
double ct_unet_flops_per_tile(const CtUNet* net);

size_t ct_unet_workspace_bytes(const CtUNet* net, int tiles_per_batch);
/* Keras `model.predict(tiles)`: tiles (B, x, y, z) float32 -> prob (B, x, y, z) float32. */
int ct_unet_predict_tiles(const CtUNet* net, const float* tiles, float* prob, int batch,
                          void* ws, size_t ws_bytes, int tiles_per_batch, void* stream);
/* One Conv3D(3, 'same') + LeakyReLU/ReLU + BatchNormalization block (unet3d.py:101-141) of the network, on
 * Keras channels-last tensors: in (B, x, y, z, Cin) float32 -> out (B, x, y, z, Cout) float32.  `layer` indexes the
 * network's conv blocks in graph order; engine 1..4 as in ct_unet_set_engine, or 5..7 = the tcgen05 kernels on
 * the split-fp16 activation buffers they use inside the network (5: source and destination, 6: destination only,
 * 7: source only; Cin resp. Cout % 8 == 0).  Any x, y; the tcgen05 engine needs z % 8 == 0. */
size_t ct_unet_conv_block_workspace_bytes(const CtUNet* net, int layer, int batch, int x, int y, int z);
int ct_unet_conv_block(const CtUNet* net, int layer, int engine, const float* in, float* out, int batch,
                       int x, int y, int z, void* ws, size_t ws_bytes, void* stream);
/* Number of tiles unet3_prediction visits for a volume (unet3d.py:226-228,259-279). */
int ct_unet_tile_count(const CtUNet* net, int x, int y, int z, const int shrink[3], int counts_out[3]);
/* unet3_prediction over tiles [tile_begin, tile_end) of the (i,j,k) row-major tile grid: reflect
 * pre-pad, per-tile zero 'same' padding, centre crop, scatter into prob (x,y,z).  Voxels of prob not
 * covered by the tile range are left untouched (multi-GPU tile sharding). */
int ct_unet3_prediction(const CtUNet* net, const float* vol_norm, float* prob, int x, int y, int z,
                        const int shrink[3], int tile_begin, int tile_end,
                        void* ws, size_t ws_bytes, int tiles_per_batch, void* stream);
/* unet3_prediction for one rank of a spatially decomposed volume (config 3).  The rank holds the normalised voxels
 * of the box in_lo .. in_lo+in_dim (global coordinates) and runs the tiles tile_lo <= (i,j,k) < tile_hi of the
 * volume's tile grid; every voxel those tiles read after reflect padding against the WHOLE volume (x,y,z) must lie
 * inside the box (checked, error otherwise).  Centre windows are written into prob_block, which covers the box
 * out_lo .. out_lo+out_dim; voxels outside it are dropped.  Same tiles, same kernels: the union over ranks is
 * bit-identical to ct_unet3_prediction on one GPU. */
int ct_unet3_prediction_block(const CtUNet* net, const float* vol_block, const int in_lo[3], const int in_dim[3],
                              float* prob_block, const int out_lo[3], const int out_dim[3], int x, int y, int z,
                              const int shrink[3], const int tile_lo[3], const int tile_hi[3],
                              void* ws, size_t ws_bytes, int tiles_per_batch, void* stream);

/* ------------------------------------------------------------------------------------------------
 * FFN match.  Replaces ffn.py:225-265 (FFN.call), ffn.py:268-327 (initial_matching_ffn) and
 * track.py:117-178 (initial_matching_quick).
 * ---------------------------------------------------------------------------------------------- */
typedef struct CtFFN CtFFN;
/* weights_host: W1(61,512), bn1 gamma,beta,mean,var, W2(1024,512), bn2 gamma,beta,mean,var, W3(512,1), b3(1). */
size_t ct_ffn_weight_count(void);
int ct_ffn_create(const float* weights_host, size_t n_floats, CtFFN** out);
void ct_ffn_destroy(CtFFN* ffn);
/* k-NN features (ffn.py:288-304): pts (n,3) float64 -> feat (n, 3k+1) float32.  n >= k+1 required. */
int ct_knn_features(const double* pts, int n, int k, float* feat, void* stream);
size_t ct_ffn_match_workspace_bytes(int n_ref, int n_tgt);
/* corr (M,N) float32 = FFN([feat(ref n) | feat(tgt m)]) for every pair. */
int ct_ffn_match(const CtFFN* ffn, const double* ref, int n_ref, const double* tgt, int n_tgt, int k,
                 float* corr, void* ws, size_t ws_bytes, void* stream);
/* Keras FFN.predict on arbitrary rows: x (rows,122) float32 -> out (rows) float32. */
size_t ct_ffn_predict_workspace_bytes(int rows);
int ct_ffn_predict(const CtFFN* ffn, const float* x, int rows, float* out, void* ws, size_t ws_bytes,
                   void* stream);

/* ------------------------------------------------------------------------------------------------
 * Greedy prior + PR-GLS / CPD EM.  Replaces track.py:11-114 (pr_gls_quick), trackerlite.py:242-259
 * (simple_match), :262-358 (prgls_quick, prgls_with_two_ref), :361-382, :409-417, and
 * tracker.py:1269-1289 (_predict_one_rep).
 * ---------------------------------------------------------------------------------------------- */
#define CT_PRGLS_TRACK 0     /* track.py flavour: gamma0 0.1, vol 1e8, sigma^2 >= 1, fixed iteration count */
#define CT_PRGLS_LITE 1      /* trackerlite.py flavour: gamma0 0.05, vol 1, increments, convergence test */

typedef struct CtPrglsParams {
    int mode;                /* CT_PRGLS_TRACK | CT_PRGLS_LITE */
    int max_iteration;       /* loop runs iterations 1 .. max_iteration-1 */
    double beta;
    double lambda;
    double vol;              /* 1e8 (track.py:11) or 1 (trackerlite.py:376) */
    double threshold;        /* greedy threshold: 0.5 (track.py:63) or 0.1 (trackerlite.py:242) */
} CtPrglsParams;

typedef struct CtPrglsProblem {
    const double* ref;       /* (N,3)  X / prts_ref_nx3 */
    const double* tgt;       /* (M,3)  Y / ptrs_tgt_mx3 */
    const void* corr;        /* (M,N)  FFN output (float32 or float64), or a ready prior if prior_given */
    const double* tracked;   /* (L,3)  LITE only: tracked_ref_lx3 (may be NULL when L == 0) */
    double* post;            /* (M,N)  out: P / posterior_mxn (required; also used as scratch) */
    double* ref_out;         /* (N,3)  out: T_X (TRACK) / predicted_coord_ref_nx3 (LITE) */
    double* coef;            /* (3,N)  out: C of the last iteration */
    double* tracked_out;     /* (L,3)  out, LITE only */
    int* iterations;         /* out: number of EM iterations executed (1 int), may be NULL */
    int n_ref, n_tgt, n_tracked;
    int corr_is_f64;         /* dtype of corr: 0 float32, 1 float64 */
    int prior_given;         /* 1: corr already holds the prior (skip the greedy step) */
} CtPrglsProblem;

/* Greedy prior alone.  mode TRACK: rows default 1/N, matched rows 0.1/(N-1) & 0.9 (track.py:58-70);
 * mode LITE: everything 0.1/(N-1), matched 0.9, values rounded to corr's dtype (trackerlite.py:256).
 * prior (M,N) float64; pairs (min(M,N),2) int32 (tgt m, ref n) in pick order, n_pairs 1 int (may be NULL). */
size_t ct_greedy_workspace_bytes(int n_ref, int n_tgt);
int ct_greedy_prior(const void* corr, int corr_is_f64, int n_tgt, int n_ref, int mode, double threshold,
                    double* prior, int* pairs, int* n_pairs, void* ws, size_t ws_bytes, void* stream);

/* Workspace for one problem of the given size; a batch needs the sum over its problems. */
size_t ct_prgls_workspace_bytes(int n_ref, int n_tgt, int n_tracked);
/* Runs `batch` independent EM problems (one persistent CTA each, no host round trips).
 * problems_host: array of descriptors in HOST memory (copied to the device inside the workspace). */
int ct_prgls(const CtPrglsParams* params, const CtPrglsProblem* problems_host, int batch,
             void* ws, size_t ws_bytes, void* stream);

/* _predict_one_rep: post(L,3) = pre(L,3) + (C G)^T, G[n,l] = exp(-|pre_l - inter_n|^2 / 2 beta^2). */
int ct_predict_one_rep(const double* pre, int n_tracked, const double* inter, int n_ref, double beta,
                       const double* coef, double* post, void* stream);

/* trim_mean(stack (E,L,3), proportion, axis=0) -> (L,3)  (tracker.py:1507, trackerlite.py:123). */
int ct_trim_mean(const double* stack, int e, int count, double proportion, double* out, void* stream);

/* One volume of the replay chain in one call: n_rep x _predict_one_rep (tracker.py:1269-1289, repetition i with the host
 * arrays inter[i] (n_ref[i],3), beta[i], coef[i] (3,n_ref[i]) of DEVICE pointers) followed by the single-mode trimmed mean
 * of tracker.py:1503-1507 over a stack of one.  pre, out: (n_tracked,3); scratch: 2 x n_tracked x 3 doubles. */
int ct_replay_fit(const double* pre, int n_tracked, int n_rep, const double* const* inter, const int* n_ref,
                  const double* beta, const double* const* coef, double proportion, double* scratch, double* out,
                  void* stream);

/* ------------------------------------------------------------------------------------------------
 * Watershed + centroid stage between the two hot paths.  Replaces Tracker._watershed (tracker.py:671-684) =
 * watershed_2d (watershed.py:16-52: per z slice threshold 0.5, distance_transform_edt, gaussian_filter(2),
 * peak_local_max(min_distance 7), label, watershed, find_boundaries outer) + watershed_3d (watershed.py:55-101:
 * EDT with sampling (1, 1, z_xy_ratio), gaussian_filter((2, 2, 0.3)), peak_local_max(min_distance 3,
 * exclude_border 0), label, watershed, min_size / cell_num, remove_small_objects) + relabel_sequential, and the
 * centre of mass of every cell (tracker.py:646-648).
 *   prob        (x,y,z) float32 probability map (output of ct_unet3_prediction)
 *   method      0 = "min_size" (cell_num is derived), 1 = "cell_num" (min_size is derived)
 *   gauss_w_*   HOST arrays: one-sided Gaussian weights w[j], j = 0..8 (sigma 2) and j = 0..1 (sigma 0.3), computed by
 *               the caller the way scipy.ndimage does (NumPy exp), so that the smoothing is bit-identical to SciPy's
 *   labels      (x,y,z) int32 label image, cells numbered 1..n in raster order of their seed (segmentation_auto)
 *   centres     (2,max_cells,3) float64: [0] voxel units (l_center_coordinates), [1] real units, z * z_xy_ratio
 *               (r_coordinates_segment, tracker.py:648); rows >= n_cells are not written
 *   scalars_out 4 x int32 on the DEVICE: n_cells, min_size, cell_num, background voxel count
 * Label image and centres are bit-identical to the CPU path (integer / exactly rounded fp64 work throughout). */
size_t ct_watershed_workspace_bytes(int x, int y, int z, int max_cells);
int ct_watershed_segment(const float* prob, int x, int y, int z, double z_xy_ratio, int method, int min_size,
                         int cell_num, const double* gauss_w_xy9_host, const double* gauss_w_z2_host, int32_t* labels,
                         double* centres, int max_cells, int32_t* scalars_out, void* ws, size_t ws_bytes,
                         void* stream);

/* skimage.measure.label(label image, connectivity = 3) as used by Tracker._relabel_separated_cells
 * (tracker.py:1073-1081): connected components of EQUAL non-zero value, full connectivity, numbered 1..n in raster
 * order of their first voxel.  n_out: 1 x int32 on the device. */
size_t ct_label_components_workspace_bytes(int x, int y, int z);
int ct_label_components(const int32_t* image, int x, int y, int z, int32_t* labels, int32_t* n_out, void* ws,
                        size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Accurate correction of the tracked positions.  Replaces Tracker._accurate_correction (tracker.py:1177-1191) =
 * up to max_rep (REP_NUM_CORRECTION = 20) x [_correction_once_interp (:1310-1350): labels of volume 1 moved by the
 * integer displacements (_transform_cells_quick :1352-1389), overlaps and boundary cells removed, centre of mass of
 * (probability + raw / 65536) per cell on the planes that exist in the raw stack, displacement update] with the stop
 * rule of _evaluate_correction (:1401-1413).  All repetitions are enqueued at once; a device flag turns the ones after
 * convergence into no-ops, the host never waits inside the loop.
 *   Cells of volume 1 on the interpolated grid (cal_subregions, tracker.py:1093-1110; get_subregions, track.py:501-533):
 *   vox4 (n_vox,4) int16 = x, y, z, 0 of every labelled voxel, grouped by cell; start (L+1) offsets; region_min /
 *   region_width (L,3); pad_host = per-axis maximum of region_width (HOST); xi, yi, zi = interpolated volume extents
 *   (= x, y, z * z_scaling).
 *   prob (x,y,z) float32; raw (x,y,z) of raw_dtype (0 = uint16, 1 = float32, 2 = uint8);
 *   r_tracked_t0, r_disp_prev (history.r_displacements[-1]), r_tracked_prev (history.r_tracked_coordinates[-1]),
 *   r_pred (FFN + PR-GLS prediction): (L,3) float64; on_boundary (L) int32.
 *   outputs: r_disp_out (L,3) float64, i_disp_out (L,3) int32, reps_out 2 x int32 (repetitions run, converged flag).
 * ---------------------------------------------------------------------------------------------- */
size_t ct_correction_workspace_bytes(int x, int y, int z, int n_cells, int n_vox);
int ct_accurate_correction(const int16_t* vox4, const int32_t* start, const int32_t* region_min,
                           const int32_t* region_width, int n_cells, int n_vox, const int32_t* pad_host,
                           int xi, int yi, int zi, int z_scaling, const float* prob, const void* raw,
                           int raw_dtype, int x, int y, int z, double z_xy_ratio, const double* r_tracked_t0,
                           const double* r_disp_prev, const double* r_tracked_prev, const double* r_pred,
                           const int32_t* on_boundary, int max_rep, double* r_disp_out, int32_t* i_disp_out,
                           int32_t* reps_out, void* ws, size_t ws_bytes, void* stream);

/* Tracked label image of a volume.  Replaces Tracker._transform_motion_to_image (tracker.py:1391-1399): the cells of
 * volume 1 stamped at i_disp on the planes of the raw stack, overlaps and boundary cells removed, then
 * recalculate_cell_boundaries (watershed.py:111-151: per z slice, watershed of the distance transform of the overlap
 * region seeded by the remaining labels).  labels_out (x,y,z) int32. */
size_t ct_tracked_labels_workspace_bytes(int x, int y, int z, int n_cells, int n_vox);
int ct_tracked_labels(const int16_t* vox4, const int32_t* start, const int32_t* region_min,
                      const int32_t* region_width, int n_cells, int n_vox, const int32_t* pad_host, int xi, int yi,
                      int zi, int z_scaling, const int32_t* i_disp, const int32_t* on_boundary, int x, int y, int z,
                      int32_t* labels_out, void* ws, size_t ws_bytes, void* stream);

/* recalculate_cell_boundaries (watershed.py:111-151) on given images: segmentation (x,y,z) int32 (MODIFIED: voxels
 * where overlaps > 1 are zeroed, as the reference does through its view), overlaps (x,y,z) int32. */
size_t ct_recalculate_cell_boundaries_workspace_bytes(int x, int y, int z);
int ct_recalculate_cell_boundaries(int32_t* segmentation, const int32_t* overlaps, int x, int y, int z,
                                   int32_t* labels_out, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CT3D_H_ */
