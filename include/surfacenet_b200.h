/*
 * surfacenet_b200 -- C ABI of the B200-native SurfaceNet per-cube inference hot path.
 *
 * Every entry point replaces one Python-level interface of the reference (mjiUST/SurfaceNet,
 * paths relative to the reference root); the reference has no FFI of its own (it is pure
 * Python driving Theano/cuDNN), so the "binding a maintainer would add" is the ctypes stub shown
 * in INTEGRATION.md and implemented in surfacenet_b200/_lib.py.
 *
 * Conventions
 *   - plain C: pointers + sizes only.  `*_dev` pointers are CUDA device pointers on the current
 *     device, `*_host` pointers are host memory (pinned memory makes the copies asynchronous).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Device entry
 *     points only enqueue work on `stream`; they never synchronise unless stated.
 *   - return value: SN_OK or a negative SN_ERR_* code; sn_last_error() gives the message of the
 *     last failure on the calling thread.  The Python layer maps SN_ERR_INVALID -> ValueError
 *     (the reference raises ValueError for bad shapes: utils/rayPooling.py:201-202,
 *     utils/camera.py:163-170) and everything else -> RuntimeError.
 *   - volumes are C-order (x, y, z) with z fastest, exactly the reference's meshgrid('ij')
 *     flattening (utils/CVC.py:15-20).
 */
#ifndef SURFACENET_B200_H
#define SURFACENET_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SN_OK            0
#define SN_ERR_INVALID  (-1)   /* bad argument / shape                       -> ValueError   */
#define SN_ERR_CUDA     (-2)   /* CUDA runtime error                         -> RuntimeError */
#define SN_ERR_DOMAIN   (-3)   /* input outside the supported numeric domain -> ValueError   */
#define SN_ERR_NOMEM    (-4)   /* workspace too small                        -> RuntimeError */

#define SN_MODE_FP32     0     /* CUDA-core fp32 convolutions (exact reference precision)        */
#define SN_MODE_TC_EXACT 1     /* tcgen05, fp16 hi+lo split of both operands (3 products, 2 MMAs): <= 1e-4 */
#define SN_MODE_TC_FAST  2     /* tcgen05, single-pass fp16 operands (does NOT meet 1e-4 parity) */

#define SN_ACT_RELU      0
#define SN_ACT_SIGMOID   1
#define SN_ACT_NONE      2

const char* sn_last_error(void);
int         sn_version(void);
/* number of kernels this library has launched since load / since the last reset (bench: gpu_launches) */
int64_t     sn_launch_count(void);
void        sn_launch_count_reset(void);
/* which convolution kernels ran since the last sn_launch_count_reset: counts[0] = CUDA-core fp32 units (mode fp32), counts[1] = direct
 * tcgen05 units (conv_tc.cu), counts[2] = Winograd F(2,3) tcgen05 units (conv_wg.cu).  Lets a caller / test verify that the path it asked
 * for is the path that ran (there is no silent fallback between them). */
void        sn_conv_path_counts(int64_t counts[3]);

/* Per-launch timing of the convolution units with CUDA events on the launching stream (the reference
 * has only commented-out time.time() probes: utils/rayPooling.py:69-138).  enable(1) clears and starts
 * recording, enable(0) stops; collect() synchronises the recorded events and returns, per unit of
 * surfacenet_b200/weights.py:UNITS (n_units must be 22), the summed device milliseconds and launch count. */
void sn_profile_enable(int on);
int  sn_profile_collect(double* ms_per_unit, int64_t* launches_per_unit, int n_units);

/* ------------------------------------------------------------------------------------------------
 * utils/camera.py:123-184  perspectiveProj(projection_M, xyz_3D, return_int_hw, return_depth)
 *   P_dev   (n_mats,3,4) f64;  xyz_dev (n_pts,3) f64
 *   h_out_dev, w_out_dev, depth_out_dev : (n_mats,n_pts) f64 (depth_out_dev may be NULL).
 *   round_to_int != 0 -> h, w are rint()-ed (half to even) as `.round()` does at camera.py:179;
 *   the caller casts to int64.
 */
int sn_perspective_proj(const double* P_dev, int n_mats, const double* xyz_dev, int64_t n_pts, int round_to_int,
                        double* h_out_dev, double* w_out_dev, double* depth_out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * utils/CVC.py:6-53,56-104  __colorize_cube__ / gen_coloredCubes   (+ :108-111 mean subtraction)
 *   images_dev   packed uint8 RGB images, view v is (H_v, W_v, 3) row-major at images_dev+img_offset[v]
 *   img_offset_dev (V) i64, img_hw_dev (V,2) i32 = (H_v, W_v);  views never referenced may have H=W=0
 *   P_dev        (V,3,4) f64 camera matrices, indexed by view position (main_reconstruct.py:48)
 *   xyz_dev      (B,3) f32 cube min corners;  resol_dev (B) f32
 *   views_dev    (B, 2*n_vp) i32 = selected_viewPairs[b].flatten()   (CVC.py:82)
 *   X_out_dev    (B*n_vp, 6, D,D,D) f32, channel order [A.R,A.G,A.B,B.R,B.G,B.B] (CVC.py:47,67,104), may be NULL
 *   mean6_dev    NULL -> raw colours (gen_coloredCubes);  (6) f32 -> X - mean (preprocess_augmentation, CVC.py:110-111)
 *   idx_w_out_dev, idx_h_out_dev (B, 2*n_vp, D^3) i32 and in_scope_out_dev (same shape) u8: the
 *                voxel -> pixel index map of CVC.py:39-45 (all three NULL or all three set; for tests)
 */
int sn_cvc_gather(const uint8_t* images_dev, const int64_t* img_offset_dev, const int32_t* img_hw_dev, int n_views,
                  const double* P_dev, const float* xyz_dev, const float* resol_dev, const int32_t* views_dev,
                  int n_cubes, int n_vp, int D, const float* mean6_dev, float* X_out_dev,
                  int32_t* idx_w_out_dev, int32_t* idx_h_out_dev, uint8_t* in_scope_out_dev, void* stream);

/* elementwise X[n,c,...] -= mean[c]  (utils/CVC.py:110-111 on an existing tensor) */
int sn_sub_channel_mean(float* X_dev, int64_t n, int channels, int64_t spatial, const float* mean_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * nets/SurfaceNet.py:385-402  SurfaceNet_inference(...) -> (viewPair_relativeImpt_fn, nViewPair_SurfaceNet_fn)
 *
 * sn_net_create  = build + lasagne.layers.set_all_param_values(...) (SurfaceNet.py:397-400): takes the
 *   flat list of 105 float32 host arrays in get_all_param_values order (layout: SURVEY.md App. B,
 *   surfacenet_b200/weights.py); element counts are validated against the architecture.
 */
typedef struct sn_net sn_net;
int  sn_net_create(const float* const* arrays_host, const int64_t* sizes, int n_arrays, sn_net** out);
void sn_net_destroy(sn_net* net);

/* bytes of device workspace sn_net_forward needs for (n_pair_cubes, D, mode) */
int64_t sn_net_workspace_bytes(const sn_net* net, int n_pair_cubes, int D, int mode);

/* nViewPair_SurfaceNet_fn(X[, w])  (SurfaceNet.py:343-383; main_reconstruct.py:145-146)
 *   X_dev        (n_cubes*n_vp, 6, D,D,D) f32, mean already subtracted
 *   w_dev        (n_cubes, n_vp) f32 or NULL (NULL only when n_vp == 1: SurfaceNet.py:354-357)
 *   fused_out_dev   (n_cubes, 1, D,D,D) f32   = sum_v (w/sum w) p   (nets/layers.py:325-336)
 *   unfused_out_dev (n_cubes, n_vp, D,D,D) f32 (may be NULL)
 */
int sn_net_forward(const sn_net* net, const float* X_dev, int n_cubes, int n_vp, int D, const float* w_dev,
                   float* fused_out_dev, float* unfused_out_dev, void* workspace_dev, int64_t workspace_bytes,
                   int mode, void* stream);

/* viewPair_relativeImpt_fn(features, n_samples_perGroup)  (nets/SurfaceNet.py:84-100,337)
 *   features_dev (n_rows, 258) f32 -> out_dev (n_rows / n_per_group, n_per_group) f32 softmax weights */
int sn_net_relative_importance(const sn_net* net, const float* features_dev, int64_t n_rows, int n_per_group,
                               float* out_dev, void* stream);

/* single layers, exposed for per-kernel parity tests and calibration (NCDHW fp32):
 *   conv + BatchNorm(inference) + activation: nets/SurfaceNet.py:33-74 units; `unit` indexes
 *   surfacenet_b200/weights.py:UNITS.  in (n,C_in,S,S,S) -> out (n,C_out,S,S,S); mode = SN_MODE_*. */
int sn_net_layer_conv(const sn_net* net, int unit, const float* in_dev, int n, int S, float* out_dev, int mode, void* stream);
int sn_maxpool2(const float* in_dev, int n, int C, int S, float* out_dev, void* stream);      /* SurfaceNet.py:37,46 */
/* nets/layers.py:376-390: zero-stuff by f, k^3 fixed conv (W from the parameter list), 'same'.
 *   in (n,C,S,S,S) -> written into out (n, C_total, fS,fS,fS) at channel offset c_off. */
int sn_net_layer_upsample(const sn_net* net, int unit, const float* in_dev, int n, int C, int S, float* out_dev,
                          int C_total, int c_off, void* stream);
/* nets/layers.py:321-339 ChannelPool_weightedAverage:  p (n_cubes,n_vp,vol) , w (n_cubes,n_vp) -> (n_cubes,vol) */
int sn_fuse_weighted_average(const float* p_dev, const float* w_dev, int n_cubes, int n_vp, int64_t vol, float* out_dev,
                             void* stream);

/* ------------------------------------------------------------------------------------------------
 * utils/rayPooling.py:143-260  rayPooling_1cube_numpy, batched over cubes as
 * utils/sparseCubes.py:57-62 calls it.
 *   pred_dev     (B, D,D,D) float16 (pred_is_f16 != 0) or float32
 *   has_thresh/thresh: selection `pred > thresh`, thresh already rounded to pred's dtype (numpy
 *                compares a python float against a float16 array in float16); has_thresh == 0 is
 *                prediction_thresh=None (all voxels).  Selected predictions must be > 0 (SN_ERR_DOMAIN otherwise).
 *   viewpairs_dev (B, n_vp, 2) i32; P_dev (V,3,4) f64; xyz_dev (B,3) f32; resol_dev (B) f32
 *   votes_out_dev (B, D,D,D) u8
 *   workspace: sn_raypool_workspace_bytes(B, n_vp, D)
 *   This call synchronises `stream` once at the end to read the domain-error flag.
 */
int64_t sn_raypool_workspace_bytes(int n_cubes, int n_vp, int D);
int sn_raypool_votes(const void* pred_dev, int pred_is_f16, int has_thresh, float thresh, const int32_t* viewpairs_dev,
                     const double* P_dev, int n_views, const float* xyz_dev, const float* resol_dev, int n_cubes, int n_vp,
                     int D, uint8_t* votes_out_dev, void* workspace_dev, int64_t workspace_bytes, void* stream);

/* float32 -> float16 cast of the fused prediction (utils/sparseCubes.py:115) */
int sn_cast_f32_to_f16(const float* in_dev, int64_t n, void* out_f16_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * main_reconstruct.py:132-162 -- the per-batch hot loop body as ONE call, everything resident on
 * the device: CVC gather (+mean) -> SurfaceNet -> view-pair fusion -> float16 cast -> ray-pool votes.
 *   outputs: fused_out_dev (B,1,D,D,D) f32, unfused_out_dev (B,n_vp,D,D,D) f32 or NULL,
 *            pred16_out_dev (B,D,D,D) f16, votes_out_dev (B,D,D,D) u8 or NULL (skip ray pooling)
 * In the default mode at D = 16 / 32 / 64 the gather writes the first unit's operand directly (no fp32 CVC tensor in
 * HBM); the bits are those of sn_cvc_gather followed by sn_net_forward.
 */
int64_t sn_infer_batch_workspace_bytes(const sn_net* net, int n_cubes, int n_vp, int D, int mode);
int sn_infer_batch(const sn_net* net, const uint8_t* images_dev, const int64_t* img_offset_dev, const int32_t* img_hw_dev,
                   int n_views, const double* P_dev, const float* xyz_dev, const float* resol_dev,
                   const int32_t* viewpairs_dev, const float* w_dev, int n_cubes, int n_vp, int D, float min_prob_f16,
                   float* fused_out_dev, float* unfused_out_dev, void* pred16_out_dev, uint8_t* votes_out_dev,
                   void* workspace_dev, int64_t workspace_bytes, int mode, void* stream);

/* same, HOST buffers for the per-batch arguments and results (images / cameras / weights stay
 * resident on the device, as they are per-scene constants: main_reconstruct.py:49-51,70-72).
 * Copies in: xyz, resol, viewpairs, w.  Copies out: fused f32, pred16, votes (each may be NULL); with votes requested the
 * two probability volumes leave on an internal side stream while ray pooling still runs on `stream`.
 * Synchronises `stream` (which waits for the side stream) before returning. */
int sn_infer_batch_host(const sn_net* net, const uint8_t* images_dev, const int64_t* img_offset_dev, const int32_t* img_hw_dev,
                        int n_views, const double* P_dev, const float* xyz_host, const float* resol_host,
                        const int32_t* viewpairs_host, const float* w_host, int n_cubes, int n_vp, int D, float min_prob_f16,
                        float* fused_out_host, void* pred16_out_host, uint8_t* votes_out_host,
                        void* workspace_dev, int64_t workspace_bytes, int mode, void* stream);

/* ------------------------------------------------------------------------------------------------
 * "Next" rows of the scope table: the immediate consumers of the dense outputs.
 *
 * utils/utils.py:8-42  generate_voxelLevelWeighted_coloredCubes(viewPair_coloredCubes, viewPair_surf_predictions, weight4viewPair)
 *   cvc_dev      (n_cubes*n_vp, 6, vol) f32 colours; mean6_dev != NULL adds the mean back first, in fp32, exactly as
 *                main_reconstruct.py:150 does to the mean-subtracted tensor
 *   unfused_dev  (n_cubes, n_vp, vol) f32;  w_dev (n_cubes, n_vp) f32 (may be NULL when n_vp == 1)
 *   rgb_out_dev  (n_cubes, 3, vol) u8  -- bit-exact: every fp32 operation is rounded individually in numpy's order
 */
int sn_color_fusion(const float* cvc_dev, const float* mean6_dev, const float* unfused_dev, const float* w_dev, int n_cubes,
                    int n_vp, int64_t vol, uint8_t* rgb_out_dev, void* stream);

/* utils/sparseCubes.py:9-77  dense2sparse(..., enable_centerCrop=True, cube_Dcenter, enable_rayPooling) on device-resident
 * dense volumes: centre crop [(D-Dc)/2, (D-Dc)/2+Dc)^3, keep `pred > min_prob` (rayPool_thresh == 0, the hot-loop call
 * site main_reconstruct.py:156) or `votes >= rayPool_thresh` (> 0), ordered compaction in np.where order.
 *   pred16_dev (n_cubes,D,D,D) f16; rgb_dev (n_cubes,3,D,D,D) u8 or NULL; votes_dev (n_cubes,D,D,D) u8 or NULL
 *   cube_count_dev (n_cubes) i32 kept voxels per cube; cube_offset_dev (n_cubes+1) i32 start of each cube in the flat
 *   outputs, last entry = total;  ijk_out (cap,3) u8 crop coordinates, pred_out (cap) f16, rgb_out (cap,3) u8, votes_out (cap) u8
 *   capacity = entries the flat outputs can hold (n_cubes*Dc^3 always suffices); extra voxels are dropped, the counts stay exact.
 */
int64_t sn_dense2sparse_workspace_bytes(int n_cubes, int D, int Dcenter);
int sn_dense2sparse(const void* pred16_dev, const uint8_t* rgb_dev, const uint8_t* votes_dev, int n_cubes, int D, int Dcenter,
                    float min_prob_f16, int rayPool_thresh, int32_t* cube_count_dev, int32_t* cube_offset_dev,
                    uint8_t* ijk_out_dev, void* pred_out_dev, uint8_t* rgb_out_dev, uint8_t* votes_out_dev, int64_t capacity,
                    void* workspace_dev, int64_t workspace_bytes, void* stream);

/* main_reconstruct.py:134-162 including the colour fusion (150-152) and the sparsification (154-162): like sn_infer_batch,
 * but the results leave the device as the compacted per-cube lists sparseCubes.append_dense_2sparseList builds. */
int64_t sn_infer_batch_sparse_workspace_bytes(const sn_net* net, int n_cubes, int n_vp, int D, int Dcenter, int mode);
int sn_infer_batch_sparse(const sn_net* net, const uint8_t* images_dev, const int64_t* img_offset_dev, const int32_t* img_hw_dev,
                          int n_views, const double* P_dev, const float* xyz_dev, const float* resol_dev,
                          const int32_t* viewpairs_dev, const float* w_dev, int n_cubes, int n_vp, int D, int Dcenter,
                          float min_prob_f16, int rayPool_thresh, int32_t* cube_count_dev, int32_t* cube_offset_dev,
                          uint8_t* ijk_out_dev, void* pred_out_dev, uint8_t* rgb_out_dev, uint8_t* votes_out_dev, int64_t capacity,
                          void* workspace_dev, int64_t workspace_bytes, int mode, void* stream);

/* ------------------------------------------------------------------------------------------------
 * "Next" row N4: thresholding, cross-cube denoising and the adaptive-threshold refinement on the
 * scene's sparse cubes, device resident.  Layout: ONE flat voxel array for the whole scene; cube n owns entries
 * [cube_offset[n], cube_offset[n+1]) (the order of the reference's per-cube lists, i.e. of the NPZ arrays of
 * utils/sparseCubes.py:330-366, whose cube_1st_vxlIndx_np is exactly cube_offset).
 *   cube_ijk_dev (C,3) i32 cube grid index (param 'ijk');  cube_offset_dev (C+1) i64
 *   ijk_dev (N,3) u8 voxel index inside the cube;  pred16_dev (N) f16;  votes_dev (N) u8;  masks (N) u8 (0/1)
 *   grid_extent G: every masked voxel coordinate must be < G (<= 256); workspace: sn_sparse_post_workspace_bytes(C, N, G)
 * sn_sparse_denoise and sn_sparse_adapthresh synchronise `stream` once at the end to read the error flags.
 */
int64_t sn_sparse_post_workspace_bytes(int n_cubes, int64_t n_vox, int grid_extent);

/* utils/sparseCubes.py:205-243  filter_voxels: mask = [mask &] (pred >= prob_thresh) [& (votes >= rayPool_thresh)].
 *   has_prob != 0: threshold = thresh_per_cube_dev[cube] (C doubles) or thresh_scalar when that pointer is NULL, rounded to
 *   float16 for the comparison (numpy compares a float16 array with a python float in float16);
 *   votes_dev != NULL and rayPool_thresh >= 0: votes >= rayPool_thresh;  and_into != 0: and with the incoming mask. */
int sn_sparse_filter_voxels(const void* pred16_dev, const uint8_t* votes_dev, const int64_t* cube_offset_dev, int n_cubes,
                            int64_t n_vox, const double* thresh_per_cube_dev, double thresh_scalar, int has_prob,
                            int rayPool_thresh, int and_into, uint8_t* mask_inout_dev, void* workspace_dev,
                            int64_t workspace_bytes, void* stream);

/* utils/denoising.py:145-184  denoise_crossCubes(cube_ijk_np, vxl_ijk_list, vxl_mask_list, D_cube)  (neighbor_dist = 3) and its
 * parts __cluster_inCube__ (8-62) / __mark_overlappingLabels__ (67-140) for neighbor_dist 1, 2, 3:
 *   keep_out_dev (N) u8    = np.in1d(labels, overlappingLabels): masked voxels whose cluster shares a voxel with one of the
 *                            26 neighbouring cubes (neighbour voxel + (D_cube/2)*shift == voxel), may be NULL
 *   labels_out_dev (N) u32 = scipy.ndimage.label numbering (raster order of each cluster's first voxel), 0 = masked out; may be NULL
 *   n_labels_out_dev (C) i32 clusters per cube; may be NULL */
int sn_sparse_denoise(const int32_t* cube_ijk_dev, const int64_t* cube_offset_dev, const uint8_t* ijk_dev, const uint8_t* mask_dev,
                      int n_cubes, int64_t n_vox, int grid_extent, int D_cube, int neighbor_dist, uint8_t* keep_out_dev,
                      uint32_t* labels_out_dev, int32_t* n_labels_out_dev, void* workspace_dev, int64_t workspace_bytes, void* stream);

/* utils/adapthresh.py:126-174  n_iter iterations of the per-cube threshold refinement, no host round trip:
 *   per dict cube (non-empty under init_mask_dev) and threshold perturbation [0.1, 0, -0.1]: cost = sum over the 6 face
 *   neighbours of XOR(current half, neighbour half) - beta * AND (when both halves hold >= 6 voxels), accumulated in float16
 *   as numpy 1.13 does; thresh += perturbation[argmin], clamped to max_probThresh; mask &= pred >= thresh.
 *   thresh_inout_dev (C) f64, mask_inout_dev (N) u8 (start: init mask copy), argmin_out_dev (n_iter, C) i32 or NULL (-1 = not a dict cube) */
int sn_sparse_adapthresh(const int32_t* cube_ijk_dev, const int64_t* cube_offset_dev, const uint8_t* ijk_dev, const void* pred16_dev,
                         const uint8_t* init_mask_dev, int n_cubes, int64_t n_vox, int grid_extent, int D_cube, double max_probThresh,
                         double beta, int n_iter, double* thresh_inout_dev, uint8_t* mask_inout_dev, int32_t* argmin_out_dev,
                         void* workspace_dev, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * "Next" row N3: the producers of the selected view pairs and their weights.
 *
 * utils/camera.py:275-309  viewPairAngles_wrt_pts(cameraTs, pts_xyz) -> (n_pts, n_pairs) angles; everything float32
 *   (is_f64 == 0) or float64 (the reference computes in the promoted dtype of its inputs); viewpairs_dev (n_pairs,2) i32. */
int sn_viewpair_angles(const void* cameraTs_dev, const void* pts_dev, int n_views, int64_t n_pts, const int32_t* viewpairs_dev,
                       int n_pairs, int is_f64, void* out_dev, void* stream);
/* utils/viewPairSelection.py:70-74  rows [e[c,v1,:], e[c,v2,:], dissimilarity[c,q], angle[c,q]] as float32:
 *   emb_dev (n_cubes, n_views, E) f32, dissim_dev (n_cubes, n_pairs) f32, theta_dev (n_cubes, n_pairs) f32|f64
 *   -> out_dev (n_cubes*n_pairs, 2E+2) f32, the input of sn_net_relative_importance */
int sn_viewpair_features(const float* emb_dev, const int32_t* viewpairs_dev, const float* dissim_dev, const void* theta_dev,
                         int theta_is_f64, int64_t n_cubes, int n_views, int n_pairs, int D_embedding, float* out_dev, void* stream);
/* utils/viewPairSelection.py:36  w.argsort(axis=1)[:, -N:]: per row the indices of the N largest values, ascending by value
 *   (equal values in index order; numpy leaves ties unspecified).  w_dev (n_rows, n) f64, n <= 8192 -> idx_out_dev (n_rows, N) i32 */
int sn_topn_rows(const double* w_dev, int64_t n_rows, int n, int N, int32_t* idx_out_dev, void* stream);
/* utils/earlyRejection.py:82-93  selectFromSimilarity: ((d < 0.5) & (d > 0.1)).sum(axis=1) >= N -> out_dev (n_cubes) u8 */
int sn_select_from_similarity(const float* dissim_dev, int64_t n_cubes, int n_pairs, int N, uint8_t* out_dev, void* stream);
/* utils/image.py:92-200 cropImgPatches(pyramidRate=1, cubeCenter_hw=...) fused with preprocess_patches (image.py:9-48):
 *   image_dev (H,W,3) u8 RGB; centres (n) f64 -> out_dev (n, 3, patch, patch) f32 in BGR order minus mean_bgr_dev (3) */
int sn_crop_patches(const uint8_t* image_dev, int H, int W, const double* center_h_dev, const double* center_w_dev, int64_t n,
                    int patch, const float* mean_bgr_dev, float* out_dev, void* stream);

/* nets/similarityNet.py:234-246  similarityNet_inference(model_file, imgPatch_hw_size) -> (patch2embedding_fn, embeddingPair2simil_fn)
 *   sn_simnet_create: the 30 arrays of lasagne.layers.get_all_param_values([embedding, similarity]): 13 x (conv W (Cout,Cin,3,3), b),
 *   dense W (5888,128), b (128), similarity W (1,1), b (1).  patch must be 64 (params.py:93). */
typedef struct sn_simnet sn_simnet;
int  sn_simnet_create(const float* const* arrays_host, const int64_t* sizes, int n_arrays, int patch, sn_simnet** out);
void sn_simnet_destroy(sn_simnet* net);
int64_t sn_simnet_workspace_bytes(const sn_simnet* net, int64_t n_patches);
/* patch2embedding_fn: patches_dev (n, 3, 64, 64) f32 (BGR, mean subtracted) -> emb_out_dev (n, 128) f32   similarityNet.py:23-56 */
int sn_simnet_patch2embedding(const sn_simnet* net, const float* patches_dev, int64_t n_patches, float* emb_out_dev,
                              void* workspace_dev, int64_t workspace_bytes, void* stream);
/* embeddingPair2simil_fn: rows (2m, 2m+1) of embedding_pairs_dev (2*n_pairs, E) f32 -> out_dev (n_pairs) f32   similarityNet.py:66-77 */
int sn_simnet_embeddingpair2simil(const sn_simnet* net, const float* embedding_pairs_dev, int64_t n_pairs, int D_embedding,
                                  float* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SURFACENET_B200_H */
