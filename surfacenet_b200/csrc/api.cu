// The per-batch hot loop of main_reconstruct.py:132-162 as one device-resident call.
#include "net.cuh"
#include <mutex>

using namespace sn;

static int64_t infer_ws_layout(const sn_net* h, int n_cubes, int n_vp, int D, int mode, int64_t* x_bytes, int64_t* net_bytes, int64_t* rp_bytes) {
    const int64_t V = (int64_t)D * D * D;
    *x_bytes = align_up((int64_t)n_cubes * n_vp * 6 * V * 4, 256);
    *net_bytes = sn_net_workspace_bytes(h, n_cubes * n_vp, D, mode);
    *rp_bytes = sn_raypool_workspace_bytes(n_cubes, n_vp, D);
    if (*net_bytes < 0 || *rp_bytes < 0) return -1;
    return *x_bytes + std::max(*net_bytes, *rp_bytes) + 256;
}

extern "C" int64_t sn_infer_batch_workspace_bytes(const sn_net* h, int n_cubes, int n_vp, int D, int mode) {
    if (!h || n_cubes < 0 || n_vp < 1) return -1;
    int64_t a, b, c;
    return infer_ws_layout(h, n_cubes, n_vp, D, mode, &a, &b, &c);
}

static int infer_enqueue(const sn_net* h, const uint8_t* images_dev, const int64_t* img_offset_dev, const int32_t* img_hw_dev,
                         int n_views, const double* P_dev, const float* xyz_dev, const float* resol_dev,
                         const int32_t* viewpairs_dev, const float* w_dev, int n_cubes, int n_vp, int D, float min_prob_f16,
                         float* fused_out_dev, float* unfused_out_dev, void* pred16_out_dev, uint8_t* votes_out_dev,
                         void* workspace_dev, int64_t workspace_bytes, int mode, void* stream, int32_t** rp_flags,
                         cudaEvent_t pred_ready = nullptr, bool keep_X = false) {
    SN_CHECK_ARG(h && images_dev && img_offset_dev && img_hw_dev && P_dev && xyz_dev && resol_dev && viewpairs_dev && fused_out_dev,
                 "sn_infer_batch: NULL argument");
    SN_CHECK_ARG(!votes_out_dev || pred16_out_dev, "sn_infer_batch: votes need the float16 prediction buffer");
    *rp_flags = nullptr;
    if (n_cubes == 0) return SN_OK;
    int64_t xb, nb, rb;
    const int64_t need = infer_ws_layout(h, n_cubes, n_vp, D, mode, &xb, &nb, &rb);
    if (need < 0) return SN_ERR_INVALID;
    if (!workspace_dev || workspace_bytes < need) { set_error("sn_infer_batch: workspace %lld B < %lld B", (long long)workspace_bytes, (long long)need); return SN_ERR_NOMEM; }
    float* X = (float*)workspace_dev;
    void* ws2 = (char*)workspace_dev + xb;
    const int64_t ws2_bytes = workspace_bytes - xb;
    const int64_t V = (int64_t)D * D * D;
    int rc;
    if (!keep_X && tc_gathers_directly(h->net, D, mode)) {
        // main_reconstruct.py:134-146 in one chain: the network's first pass colours its own operand (conv_wg.cu:cvc_wino_kernel), no fp32 X
        SN_CHECK_ARG(n_views >= 1 && (int64_t)n_cubes * n_vp <= (1 << 28), "sn_infer_batch: bad sizes (n_cubes=%d n_vp=%d n_views=%d)", n_cubes, n_vp, n_views);
        const CvcSource src{images_dev, img_offset_dev, img_hw_dev, n_views, P_dev, xyz_dev, resol_dev, viewpairs_dev, n_vp, h->net.mean6};
        rc = net_forward(h, nullptr, &src, n_cubes, n_vp, D, w_dev, fused_out_dev, unfused_out_dev, ws2, ws2_bytes, mode, stream);
        if (rc != SN_OK) return rc;
    } else {
        // main_reconstruct.py:134-143: CVC.gen_coloredCubes + preprocess_augmentation (mean subtraction only)
        rc = sn_cvc_gather(images_dev, img_offset_dev, img_hw_dev, n_views, P_dev, xyz_dev, resol_dev, viewpairs_dev, n_cubes, n_vp, D,
                           h->net.mean6, X, nullptr, nullptr, nullptr, stream);
        if (rc != SN_OK) return rc;
        // main_reconstruct.py:145-146: nViewPair_SurfaceNet_fn(X[, w])
        rc = sn_net_forward(h, X, n_cubes, n_vp, D, w_dev, fused_out_dev, unfused_out_dev, ws2, ws2_bytes, mode, stream);
        if (rc != SN_OK) return rc;
    }
    if (pred16_out_dev) {
        // utils/sparseCubes.py:115: prediction_sub.astype(np.float16)
        rc = sn_cast_f32_to_f16(fused_out_dev, (int64_t)n_cubes * V, pred16_out_dev, stream);
        if (rc != SN_OK) return rc;
    }
    if (pred_ready) SN_CUDA(cudaEventRecord(pred_ready, (cudaStream_t)stream));     // the probabilities are final: ray pooling only reads them
    if (votes_out_dev) {
        // utils/sparseCubes.py:57-62: rayPooling_1cube_numpy(..., prediction_thresh=min_prob) per cube
        rc = raypool_enqueue(pred16_out_dev, 1, 1, min_prob_f16, viewpairs_dev, P_dev, n_views, xyz_dev, resol_dev, n_cubes, n_vp, D,
                             votes_out_dev, ws2, ws2_bytes, stream, rp_flags);
        if (rc != SN_OK) return rc;
    }
    return SN_OK;
}

extern "C" int sn_infer_batch(const sn_net* h, const uint8_t* images_dev, const int64_t* img_offset_dev, const int32_t* img_hw_dev,
                              int n_views, const double* P_dev, const float* xyz_dev, const float* resol_dev,
                              const int32_t* viewpairs_dev, const float* w_dev, int n_cubes, int n_vp, int D, float min_prob_f16,
                              float* fused_out_dev, float* unfused_out_dev, void* pred16_out_dev, uint8_t* votes_out_dev,
                              void* workspace_dev, int64_t workspace_bytes, int mode, void* stream) {
    int32_t* flags = nullptr;
    int rc = infer_enqueue(h, images_dev, img_offset_dev, img_hw_dev, n_views, P_dev, xyz_dev, resol_dev, viewpairs_dev, w_dev,
                           n_cubes, n_vp, D, min_prob_f16, fused_out_dev, unfused_out_dev, pred16_out_dev, votes_out_dev,
                           workspace_dev, workspace_bytes, mode, stream, &flags);
    if (rc != SN_OK) return rc;
    return flags ? raypool_check(flags, stream) : SN_OK;
}

// D2H side stream of the host-buffer entry: the probability volumes (6 of the 7 result bytes per voxel) leave the device while ray pooling
// still runs on the caller's stream.  One per device, created on first use.
struct HostCopyLane { cudaStream_t st = nullptr; cudaEvent_t ready = nullptr, done = nullptr; };
static int host_copy_lane(HostCopyLane** out) {
    static HostCopyLane lanes[64];
    static std::mutex mu;                                   // first use may come from several host threads (ctypes releases the GIL)
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    SN_CUDA(cudaGetDevice(&dev));
    SN_CHECK_ARG(dev >= 0 && dev < 64, "sn_infer_batch_host: device ordinal %d", dev);
    HostCopyLane& l = lanes[dev];
    if (!l.st) {
        SN_CUDA(cudaEventCreateWithFlags(&l.ready, cudaEventDisableTiming));
        SN_CUDA(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
        SN_CUDA(cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking));
    }
    *out = &l;
    return SN_OK;
}

extern "C" int sn_infer_batch_host(const sn_net* h, const uint8_t* images_dev, const int64_t* img_offset_dev, const int32_t* img_hw_dev,
                                   int n_views, const double* P_dev, const float* xyz_host, const float* resol_host,
                                   const int32_t* viewpairs_host, const float* w_host, int n_cubes, int n_vp, int D, float min_prob_f16,
                                   float* fused_out_host, void* pred16_out_host, uint8_t* votes_out_host,
                                   void* workspace_dev, int64_t workspace_bytes, int mode, void* stream) {
    SN_CHECK_ARG(n_cubes >= 0 && n_vp >= 1 && D >= 4, "sn_infer_batch_host: bad sizes");
    if (n_cubes == 0) return SN_OK;
    SN_CHECK_ARG(xyz_host && resol_host && viewpairs_host, "sn_infer_batch_host: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t V = (int64_t)D * D * D;
    // carve the staging buffers off the tail of the workspace
    const int64_t inner = sn_infer_batch_workspace_bytes(h, n_cubes, n_vp, D, mode);
    if (inner < 0) return SN_ERR_INVALID;
    const int64_t stage = align_up(n_cubes * 12, 256) + align_up(n_cubes * 4, 256) + align_up((int64_t)n_cubes * n_vp * 8, 256) +
                          align_up((int64_t)n_cubes * n_vp * 4, 256) + align_up(n_cubes * V * 4, 256) + align_up(n_cubes * V * 2, 256) +
                          align_up(n_cubes * V, 256);
    if (!workspace_dev || workspace_bytes < inner + stage) { set_error("sn_infer_batch_host: workspace %lld B < %lld B", (long long)workspace_bytes, (long long)(inner + stage)); return SN_ERR_NOMEM; }
    Arena ar((char*)workspace_dev + inner, workspace_bytes - inner);
    float* xyz_d = ar.take<float>(n_cubes * 3);
    float* resol_d = ar.take<float>(n_cubes);
    int32_t* vp_d = ar.take<int32_t>((int64_t)n_cubes * n_vp * 2);
    float* w_d = ar.take<float>((int64_t)n_cubes * n_vp);
    float* fused_d = ar.take<float>(n_cubes * V);
    __half* p16_d = ar.take<__half>(n_cubes * V);
    uint8_t* votes_d = ar.take<uint8_t>(n_cubes * V);
    SN_CUDA(cudaMemcpyAsync(xyz_d, xyz_host, n_cubes * 12, cudaMemcpyHostToDevice, st));
    SN_CUDA(cudaMemcpyAsync(resol_d, resol_host, n_cubes * 4, cudaMemcpyHostToDevice, st));
    SN_CUDA(cudaMemcpyAsync(vp_d, viewpairs_host, (int64_t)n_cubes * n_vp * 8, cudaMemcpyHostToDevice, st));
    if (w_host) SN_CUDA(cudaMemcpyAsync(w_d, w_host, (int64_t)n_cubes * n_vp * 4, cudaMemcpyHostToDevice, st));
    int32_t* flags = nullptr;
    const bool want16 = pred16_out_host || votes_out_host;
    // with ray pooling requested, the probability copies run on the side stream under the ray-pool kernels (the staging buffers sit
    // behind the inner workspace, which ray pooling reuses, so nothing they read is overwritten)
    HostCopyLane* lane = nullptr;
    if (votes_out_host && (fused_out_host || pred16_out_host)) { int lrc = host_copy_lane(&lane); if (lrc != SN_OK) return lrc; }
    int rc = infer_enqueue(h, images_dev, img_offset_dev, img_hw_dev, n_views, P_dev, xyz_d, resol_d, vp_d, w_host ? w_d : nullptr,
                           n_cubes, n_vp, D, min_prob_f16, fused_d, nullptr, want16 ? p16_d : nullptr, votes_out_host ? votes_d : nullptr,
                           workspace_dev, inner, mode, stream, &flags, lane ? lane->ready : nullptr);
    if (rc != SN_OK) return rc;
    const cudaStream_t cst = lane ? lane->st : st;
    if (lane) SN_CUDA(cudaStreamWaitEvent(lane->st, lane->ready, 0));
    if (fused_out_host) SN_CUDA(cudaMemcpyAsync(fused_out_host, fused_d, n_cubes * V * 4, cudaMemcpyDeviceToHost, cst));
    if (pred16_out_host) SN_CUDA(cudaMemcpyAsync(pred16_out_host, p16_d, n_cubes * V * 2, cudaMemcpyDeviceToHost, cst));
    if (lane) {
        SN_CUDA(cudaEventRecord(lane->done, lane->st));
        SN_CUDA(cudaStreamWaitEvent(st, lane->done, 0));                        // the caller's stream (synchronised below) covers both
    }
    if (votes_out_host) SN_CUDA(cudaMemcpyAsync(votes_out_host, votes_d, n_cubes * V, cudaMemcpyDeviceToHost, st));
    if (flags) return raypool_check(flags, stream);
    SN_CUDA(cudaStreamSynchronize(st));
    return SN_OK;
}

// ---- main_reconstruct.py:134-162 with colour fusion and sparsification ------------------------------------------------
static int64_t sparse_ws_layout(const sn_net* h, int n_cubes, int n_vp, int D, int Dc, int mode, int64_t off[6]) {
    const int64_t V = (int64_t)D * D * D;
    const int64_t inner = sn_infer_batch_workspace_bytes(h, n_cubes, n_vp, D, mode);
    const int64_t d2s = sn_dense2sparse_workspace_bytes(n_cubes, D, Dc);
    if (inner < 0 || d2s < 0) return -1;
    int64_t o = align_up(inner, 256);
    off[0] = o; o += align_up((int64_t)n_cubes * n_vp * V * 4, 256);     // unfused
    off[1] = o; o += align_up(n_cubes * V * 4, 256);                    // fused
    off[2] = o; o += align_up(n_cubes * V * 2, 256);                    // pred16
    off[3] = o; o += align_up(n_cubes * V, 256);                        // votes
    off[4] = o; o += align_up(n_cubes * V * 3, 256);                    // rgb
    off[5] = o; o += align_up(d2s, 256);
    return o;
}

extern "C" int64_t sn_infer_batch_sparse_workspace_bytes(const sn_net* h, int n_cubes, int n_vp, int D, int Dcenter, int mode) {
    if (!h || n_cubes < 0 || n_vp < 1) return -1;
    int64_t off[6];
    return sparse_ws_layout(h, n_cubes, n_vp, D, Dcenter, mode, off);
}

extern "C" int sn_infer_batch_sparse(const sn_net* h, const uint8_t* images_dev, const int64_t* img_offset_dev, const int32_t* img_hw_dev,
                                     int n_views, const double* P_dev, const float* xyz_dev, const float* resol_dev,
                                     const int32_t* viewpairs_dev, const float* w_dev, int n_cubes, int n_vp, int D, int Dcenter,
                                     float min_prob_f16, int rayPool_thresh, int32_t* cube_count_dev, int32_t* cube_offset_dev,
                                     uint8_t* ijk_out_dev, void* pred_out_dev, uint8_t* rgb_out_dev, uint8_t* votes_out_dev, int64_t capacity,
                                     void* workspace_dev, int64_t workspace_bytes, int mode, void* stream) {
    SN_CHECK_ARG(h && n_cubes >= 0 && n_vp >= 1, "sn_infer_batch_sparse: bad arguments");
    if (n_cubes == 0) return SN_OK;
    int64_t off[6];
    const int64_t need = sparse_ws_layout(h, n_cubes, n_vp, D, Dcenter, mode, off);
    if (need < 0) return SN_ERR_INVALID;
    if (!workspace_dev || workspace_bytes < need) { set_error("sn_infer_batch_sparse: workspace %lld B < %lld B", (long long)workspace_bytes, (long long)need); return SN_ERR_NOMEM; }
    char* ws = (char*)workspace_dev;
    float* unf = (float*)(ws + off[0]); float* fused = (float*)(ws + off[1]);
    void* p16 = ws + off[2]; uint8_t* votes = (uint8_t*)(ws + off[3]); uint8_t* rgb = (uint8_t*)(ws + off[4]);
    const int64_t V = (int64_t)D * D * D;
    int32_t* flags = nullptr;
    int rc = infer_enqueue(h, images_dev, img_offset_dev, img_hw_dev, n_views, P_dev, xyz_dev, resol_dev, viewpairs_dev, w_dev, n_cubes, n_vp, D,
                           min_prob_f16, fused, n_vp > 1 ? unf : nullptr, p16, votes, workspace_dev, off[0], mode, stream, &flags, nullptr,
                           /*keep_X: the colours below are read from it*/ true);
    if (rc != SN_OK) return rc;
    // main_reconstruct.py:150-152: colours = mean-subtracted CVC (head of the workspace) + mean, fused with w * unfused p
    const float* X = (const float*)workspace_dev;
    rc = sn_color_fusion(X, h->net.mean6, n_vp > 1 ? unf : fused, w_dev, n_cubes, n_vp, V, rgb, stream);
    if (rc != SN_OK) return rc;
    // main_reconstruct.py:154-162 -> sparseCubes.dense2sparse
    rc = sn_dense2sparse(p16, rgb, votes, n_cubes, D, Dcenter, min_prob_f16, rayPool_thresh, cube_count_dev, cube_offset_dev, ijk_out_dev,
                         pred_out_dev, rgb_out_dev, votes_out_dev, capacity, ws + off[5], workspace_bytes - off[5], stream);
    if (rc != SN_OK) return rc;
    return flags ? raypool_check(flags, stream) : SN_OK;
}
