// "Next" row N3 of the scope table: similarityNet (nets/similarityNet.py:23-77), the 2-D VGG-16 patch
// embedding used by early rejection (utils/earlyRejection.py:6-55) and the embedding-pair (dis)similarity (:58-80).
//   patch (3,64,64) BGR - mean -> 13 x [3x3 conv pad 1 + bias + ReLU] with 2x2 max-pools (VGG-16 widths 64..512)
//   -> concat[flatten(pool5), centre 2x2 crops of pool1..pool4] (5888) -> L2 normalise -> Dense 128        similarityNet.py:28-56
//   pair: ||e1 - e2||_2 -> Dense(1) -> sigmoid                                                              similarityNet.py:72-77
// fp32 on the CUDA cores (FMA-bound direct convolution, 4 pixels x 32 channels per thread, weights broadcast from shared memory); this net runs once per
// (cube, view) before the hot loop and is not on the north-star path, so it gets the simple exact-precision kernel.
#include "common.cuh"
#include <algorithm>
#include <vector>
#include <math.h>

struct sn_simnet {
    int patch;
    float* conv_w[13];      // [Cout/32][Cin][9][32]
    float* conv_b[13];
    float* dense_w;         // (5888, 128)
    float* dense_b;         // (128)
    float sim_w, sim_b;     // Dense(1) of the similarity head
};

namespace sn {

static const int SIM_CIN[13] = {3, 64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512};
static const int SIM_COUT[13] = {64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512};
static const int SIM_POOL_AFTER[13] = {0, 1, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1};
constexpr int SIM_COT = 32, SIM_PX = 4, SIM_CI = 8, SIM_THREADS = 128, SIM_E = 128, SIM_CONCAT = 5888;

// out[n,co,y,x] = relu(b[co] + sum_{ci,ky,kx} in[n,ci,y+ky-1,x+kx-1] * W[co,ci,ky,kx])   (cross-correlation: Conv2DDNNLayer, flip_filters=False)
// One thread = SIM_PX consecutive pixels of a row x SIM_COT output channels (128 accumulators): every weight float4 fetched
// from shared memory (a broadcast) feeds 4 pixels x 4 channels = 16 FMAs and the 3 x 6 input window of the 4 pixels is loaded
// once per input channel, so the FMA pipe -- not the shared-memory pipe -- is the limiter.
__global__ void __launch_bounds__(SIM_THREADS)
conv2d3x3_kernel(const float* __restrict__ in, const float* __restrict__ wt, const float* __restrict__ bias, float* __restrict__ out,
                 int64_t n_img, int Cin, int Cout, int H, int W) {
    __shared__ __align__(16) float s_w[SIM_CI][9][SIM_COT];
    const int Wq = W / SIM_PX;                                       // W is a multiple of 4 for every VGG map (64 ... 4)
    const int64_t quad = (int64_t)blockIdx.x * SIM_THREADS + threadIdx.x;
    const int64_t total = n_img * H * Wq;
    const bool live = quad < total;
    const int x0 = live ? (int)(quad % Wq) * SIM_PX : 0, y = live ? (int)((quad / Wq) % H) : 0;
    const int64_t n = live ? quad / ((int64_t)Wq * H) : 0;
    const int cog = blockIdx.y;
    float acc[SIM_PX][SIM_COT];
#pragma unroll
    for (int p = 0; p < SIM_PX; ++p)
#pragma unroll
        for (int c = 0; c < SIM_COT; ++c) acc[p][c] = 0.f;
    const float* in_n = in + n * Cin * (int64_t)H * W;
    const float* wt_g = wt + (int64_t)cog * Cin * 9 * SIM_COT;
    bool okr[3], okl, okr_edge;                                      // rows y-1..y+1 inside?  columns x0-1 / x0+4 inside?
#pragma unroll
    for (int r = 0; r < 3; ++r) okr[r] = live && (y + r - 1) >= 0 && (y + r - 1) < H;
    okl = x0 > 0; okr_edge = x0 + SIM_PX < W;
    for (int ci0 = 0; ci0 < Cin; ci0 += SIM_CI) {
        const int nci = min(SIM_CI, Cin - ci0);
        __syncthreads();
        for (int i = threadIdx.x; i < nci * 9 * SIM_COT; i += SIM_THREADS) (&s_w[0][0][0])[i] = wt_g[(int64_t)ci0 * 9 * SIM_COT + i];
        __syncthreads();
        for (int ci = 0; ci < nci; ++ci) {
            const float* p = in_n + (int64_t)(ci0 + ci) * H * W + (int64_t)y * W + x0;
            float xv[3][SIM_PX + 2];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float* pr = p + (r - 1) * W;
                float4 mid = make_float4(0.f, 0.f, 0.f, 0.f);
                float lft = 0.f, rgt = 0.f;
                if (okr[r]) {
                    mid = __ldg(reinterpret_cast<const float4*>(pr));
                    if (okl) lft = __ldg(pr - 1);
                    if (okr_edge) rgt = __ldg(pr + SIM_PX);
                }
                xv[r][0] = lft; xv[r][1] = mid.x; xv[r][2] = mid.y; xv[r][3] = mid.z; xv[r][4] = mid.w; xv[r][5] = rgt;
            }
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float4* w4 = reinterpret_cast<const float4*>(&s_w[ci][t][0]);
#pragma unroll
                for (int c = 0; c < SIM_COT / 4; ++c) {
                    const float4 w = w4[c];
#pragma unroll
                    for (int q = 0; q < SIM_PX; ++q) {
                        const float xq = xv[t / 3][q + t % 3];
                        acc[q][4 * c + 0] = fmaf(xq, w.x, acc[q][4 * c + 0]);
                        acc[q][4 * c + 1] = fmaf(xq, w.y, acc[q][4 * c + 1]);
                        acc[q][4 * c + 2] = fmaf(xq, w.z, acc[q][4 * c + 2]);
                        acc[q][4 * c + 3] = fmaf(xq, w.w, acc[q][4 * c + 3]);
                    }
                }
            }
        }
    }
    if (!live) return;
    float* o = out + (n * Cout + (int64_t)cog * SIM_COT) * H * W + (int64_t)y * W + x0;
#pragma unroll
    for (int c = 0; c < SIM_COT; ++c) {
        const float b = __ldg(bias + cog * SIM_COT + c);
        *reinterpret_cast<float4*>(o + (int64_t)c * H * W) = make_float4(fmaxf(acc[0][c] + b, 0.f), fmaxf(acc[1][c] + b, 0.f),
                                                                         fmaxf(acc[2][c] + b, 0.f), fmaxf(acc[3][c] + b, 0.f));
    }
}

__global__ void maxpool2d_kernel(const float* __restrict__ in, int64_t total_out, int H, int W, float* __restrict__ out) {
    const int Ho = H / 2, Wo = W / 2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_out) return;
    const int x = (int)(i % Wo), y = (int)((i / Wo) % Ho);
    const int64_t nc = i / ((int64_t)Wo * Ho);
    const float* p = in + nc * (int64_t)H * W + (int64_t)(2 * y) * W + 2 * x;
    out[i] = fmaxf(fmaxf(p[0], p[1]), fmaxf(p[W], p[W + 1]));
}

// concat (similarityNet.py:46-52) + L2NormLayer (nets/layers.py:34-38) + Dense(128, no nonlinearity); one block per patch
__global__ void __launch_bounds__(SIM_E)
simnet_head_kernel(const float* __restrict__ pool5, const float* __restrict__ pool1, const float* __restrict__ pool2, const float* __restrict__ pool3,
                   const float* __restrict__ pool4, int S1, const float* __restrict__ Wd, const float* __restrict__ bd, float* __restrict__ emb) {
    __shared__ float v[SIM_CONCAT];
    __shared__ float red[SIM_E / 32];
    const int64_t n = blockIdx.x;
    const int S5 = S1 / 16;
    const int n5 = 512 * S5 * S5;                         // flatten(pool5): (c, h, w) order
    for (int i = threadIdx.x; i < n5; i += SIM_E) v[i] = pool5[n * n5 + i];
    int base = n5;
    const float* src[4] = {pool1, pool2, pool3, pool4};
    const int ch[4] = {64, 128, 256, 512};
    for (int l = 0; l < 4; ++l) {                          // CropFeatureMapCenterLayer(cropCenter_r=1): rows/cols [S/2-1, S/2+1)
        const int S = S1 >> l, c0 = S / 2 - 1;
        for (int i = threadIdx.x; i < ch[l] * 4; i += SIM_E) {
            const int c = i / 4, r = (i / 2) % 2, s = i % 2;
            v[base + i] = src[l][((n * ch[l] + c) * S + (c0 + r)) * (int64_t)S + (c0 + s)];
        }
        base += ch[l] * 4;
    }
    __syncthreads();
    float ss = 0.f;
    for (int i = threadIdx.x; i < base; i += SIM_E) ss = fmaf(v[i], v[i], ss);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < SIM_E / 32; ++i) tot += red[i];
    const float nrm = sqrtf(tot);
    __syncthreads();
    for (int i = threadIdx.x; i < base; i += SIM_E) v[i] = v[i] / nrm;        // incoming / input_L2[:, None]
    __syncthreads();
    float acc = 0.f;
    const int j = threadIdx.x;
    for (int k = 0; k < base; ++k) acc = fmaf(v[k], __ldg(Wd + (int64_t)k * SIM_E + j), acc);
    emb[n * SIM_E + j] = acc + bd[j];
}

// rows (2m, 2m+1) of the (2M, E) input form pair m: sigmoid(w * ||e1 - e2||_2 + b)          similarityNet.py:72-77, layers.py:131-136
__global__ void pair_simil_kernel(const float* __restrict__ e, int64_t n_pairs, int E, float w, float b, float* __restrict__ out) {
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (m >= n_pairs) return;
    const int lane = threadIdx.x & 31;
    float s = 0.f;
    for (int k = lane; k < E; k += 32) { const float d = fabsf(e[(2 * m) * E + k] - e[(2 * m + 1) * E + k]); s = fmaf(d, d, s); }
#pragma unroll
    for (int t = 16; t > 0; t >>= 1) s += __shfl_xor_sync(0xffffffffu, s, t);
    if (lane == 0) out[m] = 1.f / (1.f + expf(-(sqrtf(s) * w + b)));
}

}  // namespace sn

using namespace sn;

extern "C" int sn_simnet_create(const float* const* arrays_host, const int64_t* sizes, int n_arrays, int patch, sn_simnet** out) {
    SN_CHECK_ARG(arrays_host && sizes && out, "sn_simnet_create: NULL argument");
    SN_CHECK_ARG(n_arrays == 30, "sn_simnet_create: expected 30 parameter arrays (13 x conv W,b + dense W,b + similarity W,b), got %d", n_arrays);
    SN_CHECK_ARG(patch == 64, "sn_simnet_create: the patch size must be 64 (params.py:93), got %d", patch);
    for (int l = 0; l < 13; ++l) {
        SN_CHECK_ARG(sizes[2 * l] == (int64_t)SIM_COUT[l] * SIM_CIN[l] * 9 && sizes[2 * l + 1] == SIM_COUT[l],
                     "sn_simnet_create: conv layer %d has %lld / %lld elements", l, (long long)sizes[2 * l], (long long)sizes[2 * l + 1]);
    }
    SN_CHECK_ARG(sizes[26] == (int64_t)SIM_CONCAT * SIM_E && sizes[27] == SIM_E && sizes[28] == 1 && sizes[29] == 1,
                 "sn_simnet_create: dense / similarity parameter sizes do not match the architecture");
    sn_simnet* h = new sn_simnet();
    h->patch = patch;
    for (int l = 0; l < 13; ++l) {
        const int Cin = SIM_CIN[l], Cout = SIM_COUT[l];
        std::vector<float> t((size_t)Cout * Cin * 9);
        const float* W = arrays_host[2 * l];                                   // (Cout, Cin, 3, 3)
        for (int co = 0; co < Cout; ++co)
            for (int ci = 0; ci < Cin; ++ci)
                for (int k = 0; k < 9; ++k)
                    t[(((size_t)(co / SIM_COT) * Cin + ci) * 9 + k) * SIM_COT + co % SIM_COT] = W[((size_t)co * Cin + ci) * 9 + k];
        SN_CUDA(cudaMalloc(&h->conv_w[l], t.size() * 4));
        SN_CUDA(cudaMemcpy(h->conv_w[l], t.data(), t.size() * 4, cudaMemcpyHostToDevice));
        SN_CUDA(cudaMalloc(&h->conv_b[l], (size_t)Cout * 4));
        SN_CUDA(cudaMemcpy(h->conv_b[l], arrays_host[2 * l + 1], (size_t)Cout * 4, cudaMemcpyHostToDevice));
    }
    SN_CUDA(cudaMalloc(&h->dense_w, (size_t)SIM_CONCAT * SIM_E * 4));
    SN_CUDA(cudaMemcpy(h->dense_w, arrays_host[26], (size_t)SIM_CONCAT * SIM_E * 4, cudaMemcpyHostToDevice));
    SN_CUDA(cudaMalloc(&h->dense_b, SIM_E * 4));
    SN_CUDA(cudaMemcpy(h->dense_b, arrays_host[27], SIM_E * 4, cudaMemcpyHostToDevice));
    h->sim_w = arrays_host[28][0];
    h->sim_b = arrays_host[29][0];
    *out = h;
    return SN_OK;
}

extern "C" void sn_simnet_destroy(sn_simnet* h) {
    if (!h) return;
    for (int l = 0; l < 13; ++l) { cudaFree(h->conv_w[l]); cudaFree(h->conv_b[l]); }
    cudaFree(h->dense_w); cudaFree(h->dense_b);
    delete h;
}

// two ping-pong activation buffers of the largest layer + the five pooled maps kept for the concat
static int64_t simnet_layout(int64_t n, int P, int64_t* act_elems, int64_t pool_elems[5]) {
    *act_elems = n * 64 * P * P;
    const int ch[5] = {64, 128, 256, 512, 512};
    int64_t tot = 2 * align_up(*act_elems * 4, 256);
    for (int l = 0; l < 5; ++l) { const int S = P >> (l + 1); pool_elems[l] = n * ch[l] * S * S; tot += align_up(pool_elems[l] * 4, 256); }
    return tot + 256;
}

extern "C" int64_t sn_simnet_workspace_bytes(const sn_simnet* h, int64_t n_patches) {
    if (!h || n_patches < 0) return -1;
    int64_t a, p[5];
    return simnet_layout(std::max<int64_t>(n_patches, 1), h->patch, &a, p);
}

extern "C" int sn_simnet_patch2embedding(const sn_simnet* h, const float* patches_dev, int64_t n_patches, float* emb_out_dev,
                                         void* workspace_dev, int64_t workspace_bytes, void* stream) {
    SN_CHECK_ARG(h && n_patches >= 0, "sn_simnet_patch2embedding: bad arguments");
    if (n_patches == 0) return SN_OK;
    SN_CHECK_ARG(patches_dev && emb_out_dev, "sn_simnet_patch2embedding: NULL argument");
    int64_t act, pe[5];
    const int64_t need = simnet_layout(n_patches, h->patch, &act, pe);
    if (!workspace_dev || workspace_bytes < need) { set_error("sn_simnet_patch2embedding: workspace %lld B < %lld B", (long long)workspace_bytes, (long long)need); return SN_ERR_NOMEM; }
    cudaStream_t st = (cudaStream_t)stream;
    Arena a(workspace_dev, workspace_bytes);
    float* buf[2] = {a.take<float>(act), a.take<float>(act)};
    float* pool[5];
    for (int l = 0; l < 5; ++l) pool[l] = a.take<float>(pe[l]);
    const float* cur = patches_dev;
    int S = h->patch, which = 0, np = 0;
    for (int l = 0; l < 13; ++l) {
        const int Cin = SIM_CIN[l], Cout = SIM_COUT[l];
        float* dst = buf[which];
        const int64_t quads = n_patches * S * (S / SIM_PX);
        conv2d3x3_kernel<<<dim3((unsigned)cdiv(quads, SIM_THREADS), Cout / SIM_COT), SIM_THREADS, 0, st>>>(cur, h->conv_w[l], h->conv_b[l], dst,
                                                                                                       n_patches, Cin, Cout, S, S);
        SN_LAUNCHED();
        cur = dst; which ^= 1;
        if (SIM_POOL_AFTER[l]) {
            const int64_t tot = n_patches * Cout * (S / 2) * (S / 2);
            maxpool2d_kernel<<<(unsigned)cdiv(tot, 256), 256, 0, st>>>(cur, tot, S, S, pool[np]);
            SN_LAUNCHED();
            cur = pool[np++]; S /= 2;
        }
    }
    simnet_head_kernel<<<(unsigned)n_patches, SIM_E, 0, st>>>(pool[4], pool[0], pool[1], pool[2], pool[3], h->patch / 2, h->dense_w, h->dense_b,
                                                             emb_out_dev);
    SN_LAUNCHED();
    return SN_OK;
}

extern "C" int sn_simnet_embeddingpair2simil(const sn_simnet* h, const float* embedding_pairs_dev, int64_t n_pairs, int D_embedding,
                                             float* out_dev, void* stream) {
    SN_CHECK_ARG(h && n_pairs >= 0 && D_embedding > 0, "sn_simnet_embeddingpair2simil: bad arguments");
    if (n_pairs == 0) return SN_OK;
    SN_CHECK_ARG(embedding_pairs_dev && out_dev, "sn_simnet_embeddingpair2simil: NULL argument");
    pair_simil_kernel<<<(unsigned)cdiv(n_pairs, 8), 256, 0, (cudaStream_t)stream>>>(embedding_pairs_dev, n_pairs, D_embedding, h->sim_w, h->sim_b, out_dev);
    SN_LAUNCHED();
    return SN_OK;
}
