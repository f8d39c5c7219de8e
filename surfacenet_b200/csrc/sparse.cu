// "Next" rows N1 / N2 of the scope table: the immediate consumers of the hot path's dense outputs.
//   N2  utils/utils.py:8-42       generate_voxelLevelWeighted_coloredCubes  -> per-voxel fused RGB (uint8)
//   N1  utils/sparseCubes.py:9-77 dense2sparse: centre crop + threshold + ORDERED compaction (np.where order)
// Both are HBM-bound elementwise / stream-compaction kernels; they exist so that only the kept voxels
// (a few % of the dense volume) have to cross PCIe instead of 7 bytes per dense voxel.
#include "common.cuh"
#include <algorithm>

namespace sn {

// ---------------------------------------------------------------------------------------------------------
// Colour fusion.  Every operation is an individually rounded fp32 operation in the reference's order
// (numpy never contracts to FMA), because the result is truncated to uint8 (utils.py:42):
//   vw_v  = w[b,v] * p[b,v,x]                                   utils.py:32
//   s     = vw_0 + vw_1 + ...   (sequential over v)             utils.py:33
//   nw_v  = vw_v / s
//   mc_vc = (colA_vc + colB_vc) / 2                             utils.py:37 (np.mean over the view axis)
//   out_c = nw_0*mc_0c + nw_1*mc_1c + ...                       utils.py:40
// colours: cvc[(b*n_vp + v), 6, x] (+ mean6[c] when add_mean: main_reconstruct.py:150 `_CVCs2_sub += mean`)
__global__ void color_fusion_kernel(const float* __restrict__ cvc, const float* __restrict__ mean6, const float* __restrict__ p,
                                    const float* __restrict__ w, int n_vp, long long vol, long long total, uint8_t* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const long long b = i / vol, x = i - b * vol;
        const float* wb = w + b * n_vp;
        const float* pb = p + b * n_vp * vol + x;
        float s = 0.f;
        for (int v = 0; v < n_vp; ++v) {
            const float vw = __fmul_rn(w ? __ldg(wb + v) : 1.f, __ldg(pb + (long long)v * vol));
            s = (v == 0) ? vw : __fadd_rn(s, vw);
        }
        float acc[3] = {0.f, 0.f, 0.f};
        for (int v = 0; v < n_vp; ++v) {
            const float nw = __fdiv_rn(__fmul_rn(w ? __ldg(wb + v) : 1.f, __ldg(pb + (long long)v * vol)), s);
            const float* c6 = cvc + ((b * n_vp + v) * 6) * vol + x;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float a = __ldg(c6 + (long long)c * vol), bb = __ldg(c6 + (long long)(3 + c) * vol);
                if (mean6) { a = __fadd_rn(a, __ldg(mean6 + c)); bb = __fadd_rn(bb, __ldg(mean6 + 3 + c)); }
                const float mc = __fdiv_rn(__fadd_rn(a, bb), 2.f);
                const float term = __fmul_rn(nw, mc);
                acc[c] = (v == 0) ? term : __fadd_rn(acc[c], term);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = acc[c];                                   // astype(np.uint8): C truncation; NaN / out of range -> 0
            out[(b * 3 + c) * vol + x] = (v >= 0.f && v < 256.f) ? (uint8_t)v : (uint8_t)0;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// dense2sparse.  kept(b, l) for the crop-linear index l = (i*Dc + j)*Dc + k, source voxel (i+c0, j+c0, k+c0):
//   rayPool_thresh > 0 (with ray pooling): votes >= rayPool_thresh          sparseCubes.py:64
//   otherwise:                             pred (float16) > min_prob        sparseCubes.py:66
constexpr int D2S_THREADS = 1024;

__device__ __forceinline__ bool d2s_keep(const __half* pred, const uint8_t* votes, long long src, float min_prob, int rp_thresh) {
    if (rp_thresh > 0 && votes) return (int)votes[src] >= rp_thresh;
    return __half2float(pred[src]) > min_prob;
}

__device__ __forceinline__ long long d2s_src(int l, int Dc, int c0, int D) {
    const int k = l % Dc, j = (l / Dc) % Dc, i = l / (Dc * Dc);
    return ((long long)(i + c0) * D + (j + c0)) * D + (k + c0);
}

// pass 1: kept voxels per (cube, chunk of 1024 crop voxels)
__global__ void __launch_bounds__(D2S_THREADS)
d2s_count_kernel(const __half* __restrict__ pred, const uint8_t* __restrict__ votes, int D, int Dc, int c0, float min_prob, int rp_thresh,
                 int chunks, int32_t* __restrict__ chunk_count) {
    const int b = blockIdx.y, l = blockIdx.x * D2S_THREADS + threadIdx.x;
    const long long vol = (long long)D * D * D;
    bool keep = false;
    if (l < Dc * Dc * Dc) keep = d2s_keep(pred + b * vol, votes ? votes + b * vol : nullptr, d2s_src(l, Dc, c0, D), min_prob, rp_thresh);
    const int n = __syncthreads_count(keep);
    if (threadIdx.x == 0) chunk_count[b * chunks + blockIdx.x] = n;
}

// pass 2 (one block): exclusive scan of the chunk counts in (cube, chunk) order; per-cube totals and start offsets
__global__ void __launch_bounds__(1024)
d2s_scan_kernel(const int32_t* __restrict__ chunk_count, int n_cubes, int chunks, int32_t* __restrict__ chunk_off,
                int32_t* __restrict__ cube_count, int32_t* __restrict__ cube_off) {
    __shared__ int s_part[1024];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int total = n_cubes * chunks;
    for (int base = 0; base < total; base += 1024) {
        const int idx = base + threadIdx.x;
        const int v = idx < total ? chunk_count[idx] : 0;
        s_part[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {                      // Hillis-Steele inclusive scan
            const int t = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0;
            __syncthreads();
            s_part[threadIdx.x] += t;
            __syncthreads();
        }
        if (idx < total) chunk_off[idx] = s_base + s_part[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_base += s_part[1023];
        __syncthreads();
    }
    for (int b = threadIdx.x; b < n_cubes; b += 1024) {
        const int first = chunk_off[b * chunks];
        const int end = (b + 1 < n_cubes) ? chunk_off[(b + 1) * chunks] : s_base;
        cube_off[b] = first; cube_count[b] = end - first;
    }
    if (threadIdx.x == 0) cube_off[n_cubes] = s_base;                  // grand total
}

// pass 3: ordered write (np.where order = ascending crop-linear index)
__global__ void __launch_bounds__(D2S_THREADS)
d2s_write_kernel(const __half* __restrict__ pred, const uint8_t* __restrict__ rgb, const uint8_t* __restrict__ votes, int D, int Dc, int c0,
                 float min_prob, int rp_thresh, int chunks, const int32_t* __restrict__ chunk_off, int64_t capacity,
                 uint8_t* __restrict__ ijk_out, __half* __restrict__ pred_out, uint8_t* __restrict__ rgb_out, uint8_t* __restrict__ votes_out) {
    __shared__ int s_warp[32];
    const int b = blockIdx.y, l = blockIdx.x * D2S_THREADS + threadIdx.x;
    const long long vol = (long long)D * D * D;
    bool keep = false;
    long long src = 0;
    if (l < Dc * Dc * Dc) { src = d2s_src(l, Dc, c0, D); keep = d2s_keep(pred + b * vol, votes ? votes + b * vol : nullptr, src, min_prob, rp_thresh); }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) s_warp[wid] = __popc(m);
    __syncthreads();
    if (wid == 0) {
        int v = s_warp[lane];
        for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, off); if (lane >= off) v += t; }
        s_warp[lane] = v - s_warp[lane];                                  // exclusive
    }
    __syncthreads();
    if (!keep) return;
    const long long o = (long long)chunk_off[b * chunks + blockIdx.x] + s_warp[wid] + __popc(m & ((1u << lane) - 1));
    if (o >= capacity) return;
    const int k = l % Dc, j = (l / Dc) % Dc, i = l / (Dc * Dc);
    ijk_out[3 * o] = (uint8_t)i; ijk_out[3 * o + 1] = (uint8_t)j; ijk_out[3 * o + 2] = (uint8_t)k;      // sparseCubes.py:71 (crop coordinates)
    pred_out[o] = pred[b * vol + src];
    if (rgb_out) {
        rgb_out[3 * o] = rgb[(b * 3 + 0) * vol + src]; rgb_out[3 * o + 1] = rgb[(b * 3 + 1) * vol + src]; rgb_out[3 * o + 2] = rgb[(b * 3 + 2) * vol + src];
    }
    if (votes_out) votes_out[o] = votes[b * vol + src];
}

}  // namespace sn

using namespace sn;

extern "C" int sn_color_fusion(const float* cvc_dev, const float* mean6_dev, const float* unfused_dev, const float* w_dev, int n_cubes,
                               int n_vp, int64_t vol, uint8_t* rgb_out_dev, void* stream) {
    SN_CHECK_ARG(n_cubes >= 0 && n_vp >= 1 && vol >= 1, "sn_color_fusion: bad sizes");
    if (n_cubes == 0) return SN_OK;
    SN_CHECK_ARG(cvc_dev && unfused_dev && rgb_out_dev && (w_dev || n_vp == 1), "sn_color_fusion: NULL argument");
    const long long total = (long long)n_cubes * vol;
    const int blocks = (int)std::min<long long>(cdiv(total, 256), 148 * 32);
    color_fusion_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(cvc_dev, mean6_dev, unfused_dev, w_dev, n_vp, vol, total, rgb_out_dev);
    SN_LAUNCHED();
    return SN_OK;
}

extern "C" int64_t sn_dense2sparse_workspace_bytes(int n_cubes, int D, int Dcenter) {
    if (n_cubes < 0 || D < 1 || Dcenter < 1 || Dcenter > D) return -1;
    const int64_t chunks = cdiv((int64_t)Dcenter * Dcenter * Dcenter, D2S_THREADS);
    return 2 * align_up(n_cubes * chunks * 4, 256) + 256;
}

extern "C" int sn_dense2sparse(const void* pred16_dev, const uint8_t* rgb_dev, const uint8_t* votes_dev, int n_cubes, int D, int Dcenter,
                               float min_prob_f16, int rayPool_thresh, int32_t* cube_count_dev, int32_t* cube_offset_dev,
                               uint8_t* ijk_out_dev, void* pred_out_dev, uint8_t* rgb_out_dev, uint8_t* votes_out_dev, int64_t capacity,
                               void* workspace_dev, int64_t workspace_bytes, void* stream) {
    SN_CHECK_ARG(n_cubes >= 0 && D >= 1 && Dcenter >= 1 && Dcenter <= D && D <= 256, "sn_dense2sparse: bad sizes (n_cubes=%d D=%d Dcenter=%d)", n_cubes, D, Dcenter);
    SN_CHECK_ARG(rayPool_thresh >= 0 && (rayPool_thresh == 0 || votes_dev), "sn_dense2sparse: rayPool_thresh > 0 needs the votes");
    SN_CHECK_ARG(!rgb_out_dev || rgb_dev, "sn_dense2sparse: rgb output without rgb input");
    SN_CHECK_ARG(!votes_out_dev || votes_dev, "sn_dense2sparse: votes output without votes input");
    if (n_cubes == 0) return SN_OK;
    SN_CHECK_ARG(pred16_dev && cube_count_dev && cube_offset_dev && ijk_out_dev && pred_out_dev && capacity >= 0, "sn_dense2sparse: NULL argument");
    SN_CHECK_ARG(n_cubes <= 65535, "sn_dense2sparse: at most 65535 cubes per call");
    const int64_t need = sn_dense2sparse_workspace_bytes(n_cubes, D, Dcenter);
    if (!workspace_dev || workspace_bytes < need) { set_error("sn_dense2sparse: workspace %lld B < %lld B", (long long)workspace_bytes, (long long)need); return SN_ERR_NOMEM; }
    const int chunks = (int)cdiv((int64_t)Dcenter * Dcenter * Dcenter, D2S_THREADS);
    const int c0 = (D - Dcenter) / 2;                                   // sparseCubes.py:51
    Arena ar(workspace_dev, workspace_bytes);
    int32_t* chunk_count = ar.take<int32_t>((int64_t)n_cubes * chunks);
    int32_t* chunk_off = ar.take<int32_t>((int64_t)n_cubes * chunks);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)chunks, (unsigned)n_cubes);
    d2s_count_kernel<<<grid, D2S_THREADS, 0, st>>>((const __half*)pred16_dev, votes_dev, D, Dcenter, c0, min_prob_f16, rayPool_thresh, chunks, chunk_count);
    SN_LAUNCHED();
    d2s_scan_kernel<<<1, 1024, 0, st>>>(chunk_count, n_cubes, chunks, chunk_off, cube_count_dev, cube_offset_dev);
    SN_LAUNCHED();
    d2s_write_kernel<<<grid, D2S_THREADS, 0, st>>>((const __half*)pred16_dev, rgb_dev, votes_dev, D, Dcenter, c0, min_prob_f16, rayPool_thresh, chunks,
                                                  chunk_off, capacity, ijk_out_dev, (__half*)pred_out_dev, rgb_out_dev, votes_out_dev);
    SN_LAUNCHED();
    return SN_OK;
}
