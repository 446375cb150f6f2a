/*
 * pandora_b200.h -- C-ABI of the B200-native (sm_100a) implementation of Pandora's dense
 * cost-volume hot path (Census / SAD / SSD / ZNCC -> CBCA -> SGM -> WTA).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch/pybind types.  The
 * reference's own native boundary for this path is four pybind11 modules functions
 * (all paths relative to /root/reference/src/pandora):
 *
 *   compute_matching_costs   matching_cost/cpp/includes/census.hpp:44-51       -> pb200_census_cost_volume[_host]
 *   reverse_cost_volume      matching_cost/cpp/includes/matching_cost.hpp:39-46 -> pb200_reverse_cost_volume[_host]
 *   reverse_disp_range       matching_cost/cpp/includes/matching_cost.hpp:48-54 -> pb200_reverse_disp_range[_host]
 *   cross_support            aggregation/cpp/includes/aggregation.hpp:47-53     -> pb200_cross_support[_host]
 *   cbca                     aggregation/cpp/includes/aggregation.hpp:55-65     -> pb200_cbca_aggregate / pb200_cbca_host
 *
 * and, for the steps the reference runs in numpy or in the un-vendored libSGM plugin:
 *
 *   SadSsd.compute_cost_volume   matching_cost/sad_ssd.py:75-207        -> pb200_sad_ssd_cost_volume
 *   Zncc.compute_cost_volume     matching_cost/zncc.py:114-241          -> pb200_zncc_cost_volume
 *   MedianFilter.median_filter   filter/median.py:134-179 (size 3)      -> pb200_median3
 *   optimize_cv (libSGM plugin)  optimization/optimization.py:104-123   -> pb200_sgm
 *   WinnerTakesAll.to_disp       disparity/disparity.py:400-480         -> pb200_wta
 *   mask_invalid_variable_disparity_range / mask_border  criteria.py:291-353 -> pb200_validity_mask
 *
 * Conventions
 *   - every cost volume is float32, C-contiguous (row, col, disp): disparity is the fastest axis,
 *     NaN = not computable (matching_cost/matching_cost.py:394-397).
 *   - "d_" pointers are DEVICE pointers owned by the caller; "_host" entry points take HOST
 *     pointers, do the H2D/D2H copies themselves and synchronise before returning.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *     asynchronous on that stream and re-entrant across streams.
 *   - return value: PB200_OK or a negative error code; pb200_last_error() gives a thread-local
 *     message.  No exception crosses this boundary.  There is NO CPU fallback: without a CUDA
 *     device every compute entry point returns PB200_ERR_CUDA.
 *   - disparities are integers (subpix == 1): disparity index k <-> dmin + k.
 */
#ifndef PANDORA_B200_H
#define PANDORA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PB200_API __attribute__((visibility("default")))
#else
#define PB200_API
#endif

#define PB200_OK 0
#define PB200_ERR_BAD_ARG (-1)     /* NULL pointer, non-positive size, D out of range ...            */
#define PB200_ERR_UNSUPPORTED (-2) /* window not in {3,5,7,9,11,13}, D above the kernel's maximum ... */
#define PB200_ERR_CUDA (-3)        /* CUDA runtime error (message in pb200_last_error)               */
#define PB200_ERR_WORKSPACE (-4)   /* workspace too small                                             */

#define PB200_SGM_MAX_DISP 512     /* one warp holds a whole disparity vector: D <= 32 lanes * 16     */

/* ---- library ---------------------------------------------------------------------------------- */
PB200_API int pb200_version(void);                 /* 100 * major + minor                                        */
PB200_API const char *pb200_last_error(void);      /* thread-local, never NULL                                   */
PB200_API int pb200_device_count(void);            /* number of visible CUDA devices (0 when none / no driver)   */
PB200_API uint64_t pb200_kernel_launches(void);    /* kernels launched by this library since load (this process) */

/* Kernel-selection options (process-wide; -1 = the library's own choice, the default).  Several stages have more than one
 * kernel computing the same bits; the parity tests pin one or the other through these names:
 *   "sgm.no_wave" 1 = four-launch packed schedule;  "sgm.no_byte_tier" 1 = 16-bit cost storage;
 *   "sgm.wave_kernel" 1 = one column per warp, 2 = two columns per warp;  "census.direct" 0 / 1;  "census.tile" floats;
 *   "cbca.pipe" 1 = staged kernel;  "cbca.bands" n;  "reverse.gather" 1;  "fuse_census_sgm" 0 = pb200_disparity_host
 *   keeps the Census and SGM steps apart.
 * No option changes a result; none is read from the environment. */
PB200_API int pb200_set_option(const char *name, int value);
PB200_API int pb200_get_option(const char *name);
/* Which kernel family served the LAST call of a stage on this thread: stage in {"sgm", "cbca", "census", "reverse", "sad"}.
 * sgm: 1 float kernels, 2 packed four-launch schedule, 3 / 4 two-column wavefront (4 = Census costs computed inside),
 * 5 / 6 one-column wavefront (6 = Census inside); cbca: 10 register kernel (*detail = compile-time D or 0), 11 pipelined,
 * 12 staged; census: 20 TMA-tiled, 21 direct, 22 sub-pixel; reverse: 30 tiled, 31 gather; sad: 40 tap-ordered, 41 running sums.
 * The parity tests assert it so that a silent fall-back to a slower kernel family cannot pass unnoticed. */
PB200_API int pb200_last_path(const char *stage, int *detail);

/* ---- matching cost ---------------------------------------------------------------------------- */
/* bytes of device scratch pb200_census_cost_volume needs (two planar census-descriptor images). */
PB200_API size_t pb200_census_workspace_bytes(int H, int W, int window);

/* Census cost volume (replaces compute_matching_costs, census.cpp:97-180 + the NaN pre-fill of census.py:138).
 * d_cv[y,x,k] = popcount(cL[y,x] ^ cR[y,x+dmin+k]) or NaN.  window in {3,5,7,9,11,13}.
 * Optional fused winner-takes-all: when d_disp != NULL the kernel also writes
 * d_disp[y,x] = dmin + argmin_k (first minimum) or invalid_disparity when every k is NaN
 * (same rule as pb200_wta), and, when d_all_nan != NULL, 1/0 into d_all_nan[y,x]. */
PB200_API int pb200_census_cost_volume(const float *d_left, const float *d_right, int H, int W, int window, int dmin, int D,
                             float *d_cv, void *d_workspace, size_t workspace_bytes, float *d_disp,
                             float invalid_disparity, uint8_t *d_all_nan, void *stream);

/* Same, restricted to the rows [row_begin, row_end) of the volume (and of d_disp / d_all_nan).  Only the image rows
 * [row_begin - window/2, row_end + window/2) need to be resident when the call runs: used to overlap the host-to-
 * device copy of the images with the fill, row band by row band. */
PB200_API int pb200_census_cost_volume_rows(const float *d_left, const float *d_right, int H, int W, int window, int dmin, int D,
                                  float *d_cv, void *d_workspace, size_t workspace_bytes, float *d_disp,
                                  float invalid_disparity, uint8_t *d_all_nan, int row_begin, int row_end, void *stream);

/* Sub-pixel Census (subpix 2 / 4): compute_matching_costs with the LIST of shifted right images (census.cpp:128-155; the images
 * come from img_tools.shift_right_img, img_tools.py:713-752).  d_rights: HOST array of n_right device pointers; image 0 is the
 * right image (H, W), image i > 0 its copy resampled at column offset i / n_right, (H, W - 1), contiguous.  Cell k of the
 * (H, W, n_disp) volume = Hamming cost between the left descriptor at x and the descriptor of image k % n_right at column
 * x + k / n_right + dmin, NaN where either window leaves its image.  n_right == 1 is the integer case. */
PB200_API size_t pb200_census_subpix_workspace_bytes(int H, int W, int window, int n_right);
PB200_API int pb200_census_cost_volume_subpix(const float *d_left, const float *const *d_rights, int n_right, int H, int W, int window,
                                    int dmin, int n_disp, float *d_cv, void *d_workspace, size_t workspace_bytes, void *stream);

/* Census transform only: the planar descriptors of rows [row_begin, row_end) of both images into d_workspace
 * (census_transform, matching_cost/cpp/src/census.cpp:45-95).  Feeds pb200_census_sgm(descriptors_ready = 1). */
PB200_API int pb200_census_descriptors_rows(const float *d_left, const float *d_right, int H, int W, int window, void *d_workspace,
                                  size_t workspace_bytes, int row_begin, int row_end, void *stream);

/* Fused matching-cost + optimisation steps for a pipeline with nothing in between (state_machine.py:292-364 followed
 * by :404-419): Census (census.cpp:45-180) -> 8-path SGM (P1, P2, invalid value = window^2 + P2 + 1) [-> WTA], the
 * Hamming costs going from the descriptors straight into the first SGM pass -- the float32 Census volume is never
 * written.  d_cv_out receives the SGM volume (bit-identical to pb200_census_cost_volume + pb200_sgm), d_disp /
 * d_all_nan the optional fused WTA.  *ran = 1 when the fused kernels were launched, 0 when the configuration is not
 * eligible (window not in {3, 5}, D not in {64, 128, 256}, non-integer or large penalties, image wider than one
 * co-resident wave): nothing has been computed then and the caller runs the two steps separately.
 * Workspaces: pb200_census_workspace_bytes / pb200_sgm_workspace_bytes. */
/* Workspace and descriptor pre-pass of pb200_census_sgm.  The fused kernels read the census descriptors in a layout of
 * their own (the skewed wavefront takes the right image's descriptors as four word-shifted, padded copies so that every
 * pixel's D-wide window is a pair of aligned 16-byte copies): pb200_census_sgm_workspace_bytes sizes d_census_workspace for
 * it (>= pb200_census_workspace_bytes), pb200_census_sgm_descriptors runs only the transforms (*eligible = 0 and nothing
 * launched when the configuration is not eligible) so that pb200_census_sgm(..., descriptors_ready = 1) with the same
 * arguments can follow -- used to time the transforms apart from the SGM passes. */
PB200_API size_t pb200_census_sgm_workspace_bytes(int H, int W, int window, int dmin, int D);
PB200_API int pb200_census_sgm_descriptors(const float *d_left, const float *d_right, int H, int W, int window, int dmin, int D, float p1,
                                 float p2, void *d_census_workspace, size_t census_workspace_bytes, int *eligible, void *stream);
PB200_API int pb200_census_sgm(const float *d_left, const float *d_right, int H, int W, int window, int dmin, int D, float p1, float p2,
                     int overcounting, float *d_cv_out, void *d_census_workspace, size_t census_workspace_bytes,
                     void *d_sgm_workspace, size_t sgm_workspace_bytes, float *d_disp, float invalid_disparity,
                     uint8_t *d_all_nan, int descriptors_ready, int *ran, void *stream);
/* A batch of `nimg` pairs through one wave per pass (two-column wavefront kernels): d_left / d_right (nimg, H, W), d_cv_out
 * (nimg, H, W, D), d_disp / d_all_nan (nimg, H, W); the results are those of nimg pb200_census_sgm calls, bit for bit, and the
 * fill and drain of the wave across the SMs are paid once per batch.  Census workspace: nimg * pb200_census_sgm_workspace_bytes.
 * *ran = 0: nothing was computed (shape / parameters not eligible): call pb200_census_sgm per pair. */
PB200_API int pb200_census_sgm_batch(const float *d_left, const float *d_right, int nimg, int H, int W, int window, int dmin, int D, float p1,
                           float p2, int overcounting, float *d_cv_out, void *d_census_workspace, size_t census_workspace_bytes,
                           void *d_sgm_workspace, size_t sgm_workspace_bytes, float *d_disp, float invalid_disparity,
                           uint8_t *d_all_nan, int *ran, void *stream);

/* ---- the fused stage column-tiled over several GPUs (one process per GPU) ---------------------------------------------------
 * The skewed wavefront (pandora_b200/csrc/sgm_wave1.cu) walks SHEARED columns c = (image column + row) mod Wg, in which every
 * SGM dependency points to the right; GPU `tile` of `ntiles` owns the sheared columns [tile * Wt, (tile + 1) * Wt), Wt = Wg /
 * ntiles, and hands the path states of its last column to the next GPU through a LINK buffer in that GPU's memory (NVLink peer
 * stores issued by the kernel itself; credits flow back the same way).  There is no host-side step and no collective between
 * the tiles: all GPUs run one wave.  The result is bit-identical to pb200_census_sgm on one GPU.
 *   - d_left / d_right: nimg WHOLE images (nimg, H, Wg) on every GPU (a tile's pixels drift Wt + nimg * H - 1 image columns to
 *     the left; only the descriptors of those columns are computed).  The images of a batch follow each other in ONE wave -- the
 *     first row of image i + 1 enters the pipeline behind the last row of image i, every path starting afresh -- so the time a
 *     wave needs to cross all GPUs is paid once per pass and batch instead of once per pass and image (stream throughput).
 *   - d_cv_tile (nimg, H, Wt, D), d_disp_tile / d_all_nan_tile (nimg, H, Wt): the tiles in SHEARED layout: element (i, y, c)
 *     belongs to image column (tile * Wt + c - (i * H + y)) mod Wg of image i.
 *   - link_local: this GPU's link buffer (pb200_tile_link_bytes, zero-initialised once, e.g. by pb200_ipc_alloc); link_left /
 *     link_right: the link buffers of tiles (tile - 1) and (tile + 1) mod ntiles mapped into this process (pb200_ipc_open);
 *     with ntiles == 1 all three are the same buffer.
 *   - epoch (1 .. 65535): the same on all GPUs for one call, different for consecutive calls (the links are never cleared;
 *     every word carries the epoch); prev_epoch / prev_rows: epoch and nimg * H of the call that used the links before (0: none).
 *   - passes: 1 = descriptors + the top-down pass, 2 = the bottom-up pass (+ WTA), 3 = both (the usual call).
 *   - census workspace: nimg * pb200_census_sgm_workspace_bytes(H, Wg, ...).
 * pb200_ipc_*: device memory that other processes of the node can map (cudaIpc; the 64-byte handle travels through any host
 * channel, e.g. torch.distributed). */
PB200_API size_t pb200_tile_link_bytes(int D);
PB200_API int pb200_census_sgm_tile(const float *d_left, const float *d_right, int nimg, int H, int Wg, int window, int dmin, int D, float p1,
                          float p2, int overcounting, int tile, int ntiles, float *d_cv_tile, void *d_census_workspace,
                          size_t census_workspace_bytes, void *d_sgm_workspace, size_t sgm_workspace_bytes, float *d_disp_tile,
                          float invalid_disparity, uint8_t *d_all_nan_tile, void *link_local, void *link_left, void *link_right,
                          unsigned epoch, unsigned prev_epoch, unsigned prev_rows, int passes, void *stream);
PB200_API int pb200_ipc_alloc(size_t bytes, void **d_ptr, void *handle64);
PB200_API int pb200_ipc_open(const void *handle64, void **d_ptr);
PB200_API int pb200_ipc_close(void *d_ptr);
PB200_API int pb200_ipc_free(void *d_ptr);

/* SAD (squared == 0) / SSD (squared != 0) cost volume, window odd >= 1 (sad_ssd.py:180-206). */
PB200_API int pb200_sad_ssd_cost_volume(const float *d_left, const float *d_right, int H, int W, int window, int dmin, int D,
                              int squared, float *d_cv, void *stream);

/* ZNCC cost volume, window odd >= 1, float64 window statistics like img_tools.py:834-952. */
PB200_API size_t pb200_zncc_workspace_bytes(int H, int W);
PB200_API int pb200_zncc_cost_volume(const float *d_left, const float *d_right, int H, int W, int window, int dmin, int D,
                           float *d_cv, void *d_workspace, size_t workspace_bytes, void *stream);

/* right(i,j,k) = left(i, j+k+min_disp, D-1-k) or NaN (matching_cost.cpp:26-57). */
PB200_API int pb200_reverse_cost_volume(const float *d_left_cv, int H, int W, int D, int min_disp, float *d_right_cv, void *stream);

/* Right disparity grids from the left ones (matching_cost.cpp:59-131, called by state_machine.py:673-675): for every left
 * pixel (row, col) with non-NaN bounds and every d in [int(min), int(max)], right pixel col + d sees -d; right_min / right_max
 * are the extrema of what a right pixel sees, NaN where it sees nothing.  All four grids are float32 (H, W). */
PB200_API int pb200_reverse_disp_range(const float *d_left_min, const float *d_left_max, int H, int W, float *d_right_min,
                             float *d_right_max, void *stream);

/* ---- aggregation (CBCA) ----------------------------------------------------------------------- */
/* 3x3 NaN-aware median; border ring and NaN pixels unchanged (filter/median.py:134-179). */
PB200_API int pb200_median3(const float *d_in, int H, int W, float *d_out, void *stream);

/* Cross-support arms (aggregation.cpp:224-321) of the (H, W) view starting at d_img with row pitch
 * `pitch` (elements).  nan_as_inf != 0 treats NaN pixels as +inf (cbca.py:233).  d_cross is
 * (H, W, 4) int16 in the order (left, right, up, bottom). */
PB200_API int pb200_cross_support(const float *d_img, int H, int W, int pitch, int len_arms, float intensity, int nan_as_inf,
                        int16_t *d_cross, void *stream);

/* Cross-based aggregation of a whole volume (cbca.py:127-177 + aggregation.cpp:28-221), all
 * disparities in one launch.  The volume is the full (H, W, D) one; `offset` = half window: only
 * the interior [offset, H-offset) x [offset, W-offset) is aggregated (supports are (H-2o, W-2o, 4)),
 * the border ring is copied unchanged.  d_cv_out may not alias d_cv_in. */
PB200_API int pb200_cbca_aggregate(const float *d_cv_in, float *d_cv_out, int H, int W, int D, int dmin, int offset,
                         const int16_t *d_cross_left, const int16_t *d_cross_right, int len_arms, void *stream);

/* ---- optimisation (SGM) ----------------------------------------------------------------------- */
PB200_API size_t pb200_sgm_workspace_bytes(int H, int W, int D);

/* 8-path semi-global matching on a min-type cost volume (negate max-type volumes around the call).
 * NaN -> invalid_value inside, NaN restored in the output.  overcounting != 0 subtracts 7*C.
 * d_cv_out may not alias d_cv_in.  Optional fused WTA like pb200_census_cost_volume.
 * Path-state hand-over for row-tiled multi-GPU runs: when d_halo_in_top / d_halo_in_bottom are not
 * NULL they hold the three downward (S, SE, SW) / upward (N, NE, NW) path states (3, W, D) of the
 * row just above / below this tile; d_halo_out_* receive this tile's last / first row states.
 * `dir_mask` selects the directions to run, bit r = r-th direction of E, W, S, SE, SW, N, NE, NW (always
 * executed in that order).  `init_final` bit 0: the first direction of this call initialises d_cv_out
 * (otherwise it accumulates into it); bit 1: the last direction of this call finalises it (NaN restore,
 * overcounting, fused WTA).  A single-GPU call uses dir_mask = 0xFF, init_final = 3; a tiled run splits
 * the directions over several calls (pandora_b200/tiling.py).  Halo planes are indexed by the direction's
 * rank inside its group: (S, SE, SW) for the top/bottom-out pair, (N, NE, NW) for the other.
 * `init_final` bit 2 (value 4), split calls only: the caller allows PACKED intermediates -- between the calls
 * d_cv_out and the halo buffers may hold the 16-bit representation of the exact integer fast path instead of
 * float32 (same sizes, opaque contents).  Call sequence, all on the same workspace:
 *   1. dir_mask 0x03, init_final 1|4      packed E + W (verifies the data; raises the int flag at
 *                                         d_workspace + pb200_sgm_flag_offset(W, D) when it does not qualify)
 *   2. (ranks that exchange halos max-reduce the flag among themselves)
 *   3. dir_mask 0x03, init_final 1|4|8    float E + W, executed only where the flag is raised
 *   4. dir_mask 0x1C and 0xE0 (any order), init_final 4, the last one 2|4: packed sweep, or float sweep when flagged.
 * When the shape / penalties are not eligible for the packed path the same sequence runs the float kernels. */
PB200_API size_t pb200_sgm_flag_offset(int W, int D);
PB200_API int pb200_sgm(const float *d_cv_in, float *d_cv_out, int H, int W, int D, float p1, float p2, float invalid_value,
              int overcounting, int dir_mask, int init_final, const float *d_halo_in_top, const float *d_halo_in_bottom,
              float *d_halo_out_bottom, float *d_halo_out_top, float *d_disp, int dmin, float invalid_disparity,
              uint8_t *d_all_nan, void *d_workspace, size_t workspace_bytes, void *stream);

/* use_confidence of the SGM step (docs/source/userguide/plugins/plugin_libsgm.rst:38-47): out(p, d) = cv(p, d) * confidence(p),
 * NaN costs stay NaN; d_out may be d_cv.  The scaled volume then goes through pb200_sgm (float costs: the float kernels). */
PB200_API int pb200_scale_volume(const float *d_cv, const float *d_confidence, int H, int W, int D, float *d_out, void *stream);

/* min_cost_paths of the SGM step (plugin_libsgm.rst:411-413): the 8-path optimisation like pb200_sgm (all directions, float
 * kernels, one launch per direction) plus the map "optimization_plugin_libsgm_nb_of_directions" (float32 (H, W)): how many
 * of the 8 directions have the minimum of their own path cost at the disparity where the summed cost is minimal (first
 * minima; 0 for a pixel without any valid cost). */
PB200_API size_t pb200_sgm_paths_workspace_bytes(int H, int W);
PB200_API int pb200_sgm_min_cost_paths(const float *d_cv_in, float *d_cv_out, int H, int W, int D, float p1, float p2, float invalid_value,
                             int overcounting, float *d_nb_of_directions, void *d_workspace, size_t workspace_bytes, void *stream);

/* ---- disparity (WTA) -------------------------------------------------------------------------- */
/* d_disp[y,x] = dmin + argmin_k cv (is_max: argmax), NaN never wins, first index on ties, all-NaN ->
 * invalid_disparity (disparity.py:434-455, 483-553).  d_all_nan (optional) gets 1 for all-NaN pixels. */
PB200_API int pb200_wta(const float *d_cv, int H, int W, int D, int dmin, int is_max, float invalid_disparity, float *d_disp,
              uint8_t *d_all_nan, void *stream);

/* criteria.validity_mask without input masks (criteria.py:106-147): bit 2 (incomplete range) / bit 1
 * (range missing) per column from [dmin, dmax] and the half-window `offset`.  d_mask (H, W) uint16. */
PB200_API int pb200_validity_mask_init(uint16_t *d_mask, int H, int W, int dmin, int dmax, int offset, void *stream);

/* Validity-mask updates, d_mask (H, W) uint16 in place.  wta_invalidate == 0: the cv_masked side effects
 * (criteria.py:291-353): pixels flagged in d_all_nan get bit 1 when not already set, then the
 * `offset`-wide border ring is overwritten with 1.  wta_invalidate != 0: the WTA rule only
 * (disparity.py:470-474): all-NaN pixels without an invalid bit := PANDORA_MSK_PIXEL_INVALID. */
PB200_API int pb200_validity_mask(uint16_t *d_mask, const uint8_t *d_all_nan, int H, int W, int offset, int wta_invalidate,
                        void *stream);

/* ---- input masks and per-pixel disparity grids ---------------------------- */
/* One byte of flags per pixel from an image mask `msk` (int16, (H, W)): bit 0 = a no_data pixel lies inside the
 * window x window neighbourhood (binary_dilation_msk, criteria.py:36-63), bit 1 = neither valid_pixels nor no_data
 * (masks_dilatation, matching_cost/matching_cost.py:527-553), bit 2 = msk != valid_pixels (criteria.py:171). */
PB200_API int pb200_mask_flags(const int16_t *d_msk, int H, int W, int valid_pixels, int no_data, int window, uint8_t *d_flags,
                     void *stream);

/* criteria.validity_mask with image masks: adds to a mask initialised by pb200_validity_mask_init the bits of
 * allocate_left_mask (criteria.py:178-213; d_flags_left may be NULL), allocate_right_mask (criteria.py:216-288;
 * d_flags_right may be NULL) and, when a right mask and the (H, W) float32 disparity grids are given,
 * mask_partially_missing_variable_ranges (criteria.py:161-175, cpp/src/criteria.cpp:27-110; needs grid_min <= grid_max). */
PB200_API int pb200_validity_mask_masks(uint16_t *d_mask, int H, int W, int dmin, int dmax, int offset, const uint8_t *d_flags_left,
                              const uint8_t *d_flags_right, const float *d_grid_min, const float *d_grid_max, void *stream);

/* The masking part of cv_masked (matching_cost/matching_cost.py:815-856, subpix 1, step 1), in place: cell (y, x, k)
 * becomes NaN when the left pixel or the right pixel x + dmin + k carries flag bit 0 or 1 (only where that column is
 * inside the image), or when dmin + k lies outside [grid_min(y,x), grid_max(y,x)].  Flags / grids may be NULL.
 * d_all_nan (optional) receives 1 for pixels whose whole vector is NaN afterwards -- the input of
 * pb200_validity_mask(..., wta_invalidate = 0), which finishes cv_masked (criteria.py:291-353). */
PB200_API int pb200_cv_masked(float *d_cv, int H, int W, int D, int dmin, const uint8_t *d_flags_left, const uint8_t *d_flags_right,
                    const float *d_grid_min, const float *d_grid_max, uint8_t *d_all_nan, void *stream);

/* ---- fast cross-checking -------------------------------------------------- */
/* Right disparity map straight from the LEFT volume: WTA over right(i, j, k) = left(i, j + k + min_disp_right, D-1-k)
 * without materialising the right volume (replaces reverse_cost_volume, matching_cost.cpp:26-57, followed by
 * WinnerTakesAll.to_disp as run by state_machine.py:436-448).  d_disp[i, j] = min_disp_right + argmin_k (argmax when
 * is_max), first index on ties, invalid_disparity when every k is NaN (d_all_nan, optional, flags those).
 * Needs D % 4 == 0, D <= 992 and a 16-byte aligned volume; otherwise PB200_ERR_UNSUPPORTED (use
 * pb200_reverse_cost_volume + pb200_wta). */
PB200_API int pb200_wta_right(const float *d_left_cv, int H, int W, int D, int min_disp_right, int is_max, float invalid_disparity,
                    float *d_disp, uint8_t *d_all_nan, void *stream);

/* CrossCheckingAccurate.disparity_checking (validation/validation.py:226-371): for every valid left pixel whose match
 * rint(x + disp_left) lies inside the right map, distance = abs(disp_right + disp_left) goes to d_conf (NaN elsewhere;
 * d_conf may be NULL) and, when distance > threshold, the pixel gets PANDORA_MSK_PIXEL_MISMATCH if some d in
 * [dmin, dmax] has rint(disp_right(x + d)) == -d, else PANDORA_MSK_PIXEL_OCCLUSION.  offset > 0 re-applies
 * mask_border (criteria.py:325-353).  d_mask_left is updated in place. */
PB200_API int pb200_cross_checking(const float *d_disp_left, uint16_t *d_mask_left, const float *d_disp_right, int H, int W,
                         float threshold, int dmin, int dmax, int offset, float *d_conf, void *stream);

/* ---- disparity filter ------------------------------------------------------------------------------ */
/* MedianFilter.filter_disparity with filter_size 3 (filter/median.py:96-179): invalid pixels (validity mask &
 * PANDORA_MSK_PIXEL_INVALID) are ignored by the 3x3 NaN-median and left untouched, every finite valid pixel is replaced
 * by the median of its valid neighbourhood (border ring unchanged).  d_scratch: 2 * H * W floats. */
PB200_API int pb200_filter_median3(float *d_disp, const uint16_t *d_mask, int H, int W, float *d_scratch, void *stream);

/* ---- sub-pixel refinement -------------------------------------------------- */
/* loop_refinement (approximate == 0) / loop_approximate_refinement (approximate != 0), refinement/cpp/src/
 * refinement.cpp:29-181, with method 0 = vfit (vfit.cpp:28-55) or 1 = quadratic (quadratic.cpp:28-49).
 * d_min / d_max = first / last disparity coordinate of the volume, subpix = cv.attrs["subpixel"].
 * approximate == 2: loop_refinement of a RIGHT disparity map on the reversed volume right(i, j, k) =
 * left(i, j + k + d_min, D-1-k) read straight from the LEFT volume d_cv (the right refinement of the
 * cross_checking_fast mode, state_machine.py:488-490, without the right volume); d_min / d_max are then the right
 * coordinates (-dmax_left, -dmin_left) and subpix must be 1.
 * d_disp (H, W) float32 and d_mask (H, W) uint16 are updated in place; d_itp_coeff (H, W) gets the interpolated cost. */
PB200_API int pb200_refinement(const float *d_cv, int H, int W, int D, double d_min, double d_max, int subpix, int is_max, int method,
                     int approximate, float *d_disp, uint16_t *d_mask, float *d_itp_coeff, void *stream);

/* ---- cost-volume confidence ------------------------------------------------ */
PB200_API size_t pb200_confidence_workspace_bytes(int H, int W, int n_etas);

/* Ambiguity (cost_volume_confidence/cpp/src/ambiguity.cpp:28-142) and risk (risk.cpp:28-197) in one pass over the
 * volume (plus one pass for the global extrema, cost_volume_confidence_tools.cpp:40-87).  `etas` is a HOST array
 * (np.arange(eta_min, eta_max, eta_step): non-negative, non-decreasing; ambiguity compares in float32, risk in
 * float64 like the reference).  d_grids: (2, H, W) int32 per-pixel [disp_min, disp_max] or NULL (whole range);
 * d_disparity_range: (D) float32 device array.  is_max negates the volume on the fly (ambiguity.py:135-137).
 * Outputs, each optional: d_ambiguity (H, W); d_sampled_ambiguity (H, W, n_etas); the four risk maps (together);
 * sampled risks (H, W, n_etas, together).  d_sampled_ambiguity_in: the sampled ambiguity risk.cpp takes as input,
 * NULL = the one computed by this call (what Risk.confidence_prediction does, risk.py:141-153). */
PB200_API int pb200_confidence(const float *d_cv, int H, int W, int D, int is_max, const double *etas, int n_etas,
                     const int32_t *d_grids, const float *d_disparity_range, float *d_ambiguity, float *d_sampled_ambiguity,
                     const float *d_sampled_ambiguity_in, float *d_risk_max, float *d_risk_min, float *d_disp_sup,
                     float *d_disp_inf, float *d_sampled_risk_max, float *d_sampled_risk_min, void *d_workspace,
                     size_t workspace_bytes, void *stream);

/* ---- host-buffer entry points (what a reference-side binding calls; synchronous) --------------- */
/* compute_matching_costs(img_left, [img_right], cv, disps, w, w): dmin = lround(disps[0]) (census.cpp:109). */
PB200_API int pb200_census_cost_volume_host(const float *left, const float *right, int H, int W, int window,
                                  const float *disps, int D, float *cv);
/* compute_matching_costs(img_left, imgs_right_shift, cv, disps, w, h) with the whole list of shifted right images
 * (census.hpp:44-51): rights[0] is (H, W), rights[i > 0] are (H, W - 1). */
PB200_API int pb200_census_cost_volume_multi_host(const float *left, const float *const *rights, int n_right, int H, int W, int window,
                                        const float *disps, int n_disp, float *cv);
PB200_API int pb200_reverse_cost_volume_host(const float *left_cv, int H, int W, int D, int min_disp, float *right_cv);
/* reverse_disp_range(left_min, left_max) -> (right_min, right_max) (matching_cost.hpp:48-54). */
PB200_API int pb200_reverse_disp_range_host(const float *left_min, const float *left_max, int H, int W, float *right_min, float *right_max);
PB200_API int pb200_cross_support_host(const float *image, int H, int W, int len_arms, float intensity, int16_t *cross);
/* one aggregation_cpp.cbca call: (H, W) float32 slice, supports (H, W, 4), n valid columns
 * range_col[i] -> range_col_right[i]; outputs step4 and sum4 (H, W) like aggregation.cpp:323-355. */
PB200_API int pb200_cbca_host(const float *input, int H, int W, const int16_t *cross_left, const int16_t *cross_right,
                    const int64_t *range_col, const int64_t *range_col_right, int n, float *step4, float *sum4);
/* whole pipeline on host images: matching cost (method 0 census, 1 sad, 2 ssd, 3 zncc) ->
 * optional CBCA (cbca_distance > 0) -> optional SGM (sgm_p2 > 0) -> WTA.  disp_map (H, W) float32 and
 * validity_mask (H, W) uint16 (may be NULL) are written; cv_out (may be NULL) receives the final volume. */
PB200_API int pb200_disparity_host(const float *left, const float *right, int H, int W, int method, int window, int dmin,
                         int dmax, int cbca_distance, float cbca_intensity, float sgm_p1, float sgm_p2,
                         int sgm_overcounting, float invalid_disparity, float *disp_map, uint16_t *validity_mask,
                         float *cv_out);

#ifdef __cplusplus
}
#endif
#endif /* PANDORA_B200_H */
