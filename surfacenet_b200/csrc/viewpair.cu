// "Next" row N3 of the scope table, selection side: the producers of the view pairs and weights the hot
// path consumes.
//   utils/camera.py:275-309          viewPairAngles_wrt_pts                         -> sn_viewpair_angles
//   utils/viewPairSelection.py:74    feature rows [e_v1 | e_v2 | dissimilarity | angle]  -> sn_viewpair_features
//   utils/viewPairSelection.py:8-41  __argmaxN_viewPairs__ (argsort()[:, -N:])      -> sn_topn_rows
//   utils/earlyRejection.py:82-93    selectFromSimilarity                           -> sn_select_from_similarity
//   utils/image.py:9-48,92-200       cropImgPatches(pyramidRate=1) + preprocess_patches -> sn_crop_patches
// Small gather / sort kernels (HBM or latency bound); the similarityNet itself is csrc/simnet.cu.
#include "common.cuh"
#include <algorithm>
#include <math.h>

namespace sn {

// unit vectors point->camera, cosine of every camera pair, clip, arccos; evaluated in T (the reference computes in the
// promoted dtype of its inputs) with individually rounded operations in numpy's order (no FMA contraction).
template <typename T> struct VpOps;
template <> struct VpOps<float> {
    static __device__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ float sqrt_(float a) { return __fsqrt_rn(a); }
    static __device__ float acos_(float a) { return acosf(a); }
};
template <> struct VpOps<double> {
    static __device__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ double sqrt_(double a) { return __dsqrt_rn(a); }
    static __device__ double acos_(double a) { return acos(a); }
};

template <typename T>
__global__ void vp_angles_kernel(const T* __restrict__ camT, const T* __restrict__ pts, const int32_t* __restrict__ pairs, int64_t n_pts,
                                 int n_pairs, T* __restrict__ out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_pts * n_pairs) return;
    const int64_t p = g / n_pairs;
    const int q = (int)(g - p * n_pairs);
    typedef VpOps<T> O;
    T u[2][3];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int v = pairs[2 * q + s];
        T d[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) d[a] = O::add(pts[3 * p + a], -camT[3 * v + a]);                       // camera.py:300
        const T nrm = O::sqrt_(O::add(O::add(O::mul(d[0], d[0]), O::mul(d[1], d[1])), O::mul(d[2], d[2])));   // np.linalg.norm
#pragma unroll
        for (int a = 0; a < 3; ++a) u[s][a] = O::div(d[a], nrm);                                             // :301
    }
    T c = O::add(O::add(O::mul(u[0][0], u[1][0]), O::mul(u[0][1], u[1][1])), O::mul(u[0][2], u[1][2]));       // :307
    c = c < (T)-1 ? (T)-1 : (c > (T)1 ? (T)1 : c);                                                            // np.clip; NaN passes through
    out[g] = O::acos_(c);                                                                                     // :298
}

// features[(c*n_pairs + q), :] = [e[c, v1, :], e[c, v2, :], d[c, q], theta[c, q]] as float32          viewPairSelection.py:70-74
__global__ void vp_features_kernel(const float* __restrict__ emb, const int32_t* __restrict__ pairs, const float* __restrict__ dis,
                                   const void* __restrict__ theta, int theta_is_f64, int64_t n_cubes, int n_views, int n_pairs, int E,
                                   float* __restrict__ out) {
    const int F = 2 * E + 2;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_cubes * n_pairs * F) return;
    const int f = (int)(g % F);
    const int64_t row = g / F;
    const int q = (int)(row % n_pairs);
    const int64_t c = row / n_pairs;
    float v;
    if (f < E) v = emb[(c * n_views + pairs[2 * q]) * E + f];
    else if (f < 2 * E) v = emb[(c * n_views + pairs[2 * q + 1]) * E + (f - E)];
    else if (f == 2 * E) v = dis[c * n_pairs + q];
    else v = theta_is_f64 ? (float)((const double*)theta)[c * n_pairs + q] : ((const float*)theta)[c * n_pairs + q];
    out[g] = v;
}

// per row: indices of the N largest values in ascending value order = argsort(row)[-N:]; equal values keep index order
// (numpy's default sort leaves ties unspecified).  Bitonic sort of (value, index) in shared memory, n <= 8192.
constexpr int TOPN_MAX = 8192;
__global__ void __launch_bounds__(1024)
topn_rows_kernel(const double* __restrict__ w, int n, int n_pad, int N, int32_t* __restrict__ idx_out) {
    extern __shared__ unsigned char smem_raw[];
    double* key = (double*)smem_raw;
    int* idx = (int*)(key + n_pad);
    const double* row = w + (int64_t)blockIdx.x * n;
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
        key[i] = i < n ? row[i] : INFINITY;               // padding (idx >= n) sorts after every real entry, NaN included
        idx[i] = i;
    }
    __syncthreads();
    for (int k = 2; k <= n_pad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const double a = key[i], b = key[l];
                    const int ia = idx[i], ib = idx[l];
                    // order: real values ascending (ties by index) < NaN entries (by index, numpy sorts NaN last) < padding
                    const bool pa = ia >= n, pb = ib >= n, na = a != a, nb = b != b;
                    const bool a_gt_b = (pa != pb) ? pa : ((na != nb) ? na : ((na || a == b) ? ia > ib : a > b));
                    if (a_gt_b == up) { key[i] = b; key[l] = a; idx[i] = ib; idx[l] = ia; }
                }
            }
            __syncthreads();
        }
    for (int t = threadIdx.x; t < N; t += blockDim.x) idx_out[(int64_t)blockIdx.x * N + t] = idx[n - N + t];
}

// selectionBool = ((d < 0.5) & (d > 0.1)).sum(axis=1) >= N                                  earlyRejection.py:90-92
__global__ void select_similarity_kernel(const float* __restrict__ d, int64_t n_cubes, int n_pairs, int N, uint8_t* __restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (c >= n_cubes) return;
    const int lane = threadIdx.x & 31;
    int cnt = 0;
    for (int q = lane; q < n_pairs; q += 32) {
        const float x = d[c * n_pairs + q];
        cnt += (x < 0.5f && x > 0.1f) ? 1 : 0;                // float32 array vs python float: compared in float32
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
    if (lane == 0) out[c] = cnt >= N ? 1 : 0;
}

// cropImgPatches(pyramidRate=1): patch pixel (r, s) = img[clip(int(center_h) - P/2 + r, 0, H-1), clip(int(center_w) - P/2 + s, 0, W-1)]
// (image.py:181-192; the order-2 zoom by 1.0 is the identity), then preprocess_patches: (h,w,c) -> (c,h,w), RGB -> BGR, - mean_BGR
__global__ void crop_patches_kernel(const uint8_t* __restrict__ img, int H, int W, const double* __restrict__ ch, const double* __restrict__ cw,
                                    int64_t n, int P, const float* __restrict__ mean_bgr, float* __restrict__ out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n * 3 * P * P) return;
    const int s = (int)(g % P), r = (int)((g / P) % P), c = (int)((g / ((int64_t)P * P)) % 3);
    const int64_t p = g / ((int64_t)3 * P * P);
    const int h0 = (int)ch[p] - P / 2, w0 = (int)cw[p] - P / 2;           // .astype(np.int): truncation toward zero
    const int h = min(max(h0 + r, 0), H - 1), w = min(max(w0 + s, 0), W - 1);
    const float v = (float)img[((int64_t)h * W + w) * 3 + (2 - c)];        // BGR channel c = RGB channel 2 - c
    out[g] = __fadd_rn(v, -mean_bgr[c]);
}

}  // namespace sn

using namespace sn;

extern "C" int sn_viewpair_angles(const void* cameraTs_dev, const void* pts_dev, int n_views, int64_t n_pts, const int32_t* viewpairs_dev,
                                  int n_pairs, int is_f64, void* out_dev, void* stream) {
    SN_CHECK_ARG(n_views >= 0 && n_pts >= 0 && n_pairs >= 0, "sn_viewpair_angles: negative size");
    if (n_pts == 0 || n_pairs == 0) return SN_OK;
    SN_CHECK_ARG(cameraTs_dev && pts_dev && viewpairs_dev && out_dev, "sn_viewpair_angles: NULL argument");
    const int64_t total = n_pts * n_pairs;
    const unsigned blocks = (unsigned)cdiv(total, 256);
    if (is_f64) vp_angles_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((const double*)cameraTs_dev, (const double*)pts_dev, viewpairs_dev, n_pts, n_pairs, (double*)out_dev);
    else vp_angles_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)cameraTs_dev, (const float*)pts_dev, viewpairs_dev, n_pts, n_pairs, (float*)out_dev);
    SN_LAUNCHED();
    return SN_OK;
}

extern "C" int sn_viewpair_features(const float* emb_dev, const int32_t* viewpairs_dev, const float* dissim_dev, const void* theta_dev,
                                    int theta_is_f64, int64_t n_cubes, int n_views, int n_pairs, int D_embedding, float* out_dev, void* stream) {
    SN_CHECK_ARG(n_cubes >= 0 && n_views > 0 && n_pairs >= 0 && D_embedding > 0, "sn_viewpair_features: bad sizes");
    if (n_cubes == 0 || n_pairs == 0) return SN_OK;
    SN_CHECK_ARG(emb_dev && viewpairs_dev && dissim_dev && theta_dev && out_dev, "sn_viewpair_features: NULL argument");
    const int64_t total = n_cubes * n_pairs * (2 * D_embedding + 2);
    vp_features_kernel<<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(emb_dev, viewpairs_dev, dissim_dev, theta_dev, theta_is_f64,
                                                                                     n_cubes, n_views, n_pairs, D_embedding, out_dev);
    SN_LAUNCHED();
    return SN_OK;
}

extern "C" int sn_topn_rows(const double* w_dev, int64_t n_rows, int n, int N, int32_t* idx_out_dev, void* stream) {
    SN_CHECK_ARG(n_rows >= 0 && n >= 1 && n <= TOPN_MAX && N >= 0 && N <= n, "sn_topn_rows: need 0 <= N <= n <= %d (n=%d, N=%d)", TOPN_MAX, n, N);
    if (n_rows == 0 || N == 0) return SN_OK;
    SN_CHECK_ARG(w_dev && idx_out_dev, "sn_topn_rows: NULL argument");
    int n_pad = 2;
    while (n_pad < n) n_pad <<= 1;
    const size_t smem = (size_t)n_pad * 12;
    SN_CUDA(cudaFuncSetAttribute(topn_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TOPN_MAX * 12)));
    topn_rows_kernel<<<(unsigned)n_rows, std::min(1024, n_pad), smem, (cudaStream_t)stream>>>(w_dev, n, n_pad, N, idx_out_dev);
    SN_LAUNCHED();
    return SN_OK;
}

extern "C" int sn_select_from_similarity(const float* dissim_dev, int64_t n_cubes, int n_pairs, int N, uint8_t* out_dev, void* stream) {
    SN_CHECK_ARG(n_cubes >= 0 && n_pairs >= 0, "sn_select_from_similarity: negative size");
    if (n_cubes == 0) return SN_OK;
    SN_CHECK_ARG(dissim_dev && out_dev, "sn_select_from_similarity: NULL argument");
    select_similarity_kernel<<<(unsigned)cdiv(n_cubes, 8), 256, 0, (cudaStream_t)stream>>>(dissim_dev, n_cubes, n_pairs, N, out_dev);
    SN_LAUNCHED();
    return SN_OK;
}

extern "C" int sn_crop_patches(const uint8_t* image_dev, int H, int W, const double* center_h_dev, const double* center_w_dev, int64_t n,
                               int patch, const float* mean_bgr_dev, float* out_dev, void* stream) {
    SN_CHECK_ARG(H > 0 && W > 0 && n >= 0 && patch > 0 && patch % 2 == 0, "sn_crop_patches: bad sizes");
    if (n == 0) return SN_OK;
    SN_CHECK_ARG(image_dev && center_h_dev && center_w_dev && mean_bgr_dev && out_dev, "sn_crop_patches: NULL argument");
    const int64_t total = n * 3 * patch * patch;
    crop_patches_kernel<<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(image_dev, H, W, center_h_dev, center_w_dev, n, patch,
                                                                                     mean_bgr_dev, out_dev);
    SN_LAUNCHED();
    return SN_OK;
}
