// fp32 CUDA-core kernels of the SurfaceNet forward (NCDHW, reference precision):
//   conv(1^3 / 3^3, dilation 1|2, 'same') + folded BatchNorm + ReLU/sigmoid   nets/SurfaceNet.py:33-74
//   2^3 max-pool                                                               nets/SurfaceNet.py:37,46
//   zero-stuff + fixed k^3 up-sampling conv, written into the concat buffer    nets/layers.py:376-390, SurfaceNet.py:71
//   view-pair weighted average                                                 nets/layers.py:321-339
//   view-pair relative-importance MLP + group softmax                          nets/SurfaceNet.py:84-100
// This is SN_MODE_FP32: the exact-precision device path and the on-GPU cross-check for the
// tensor-core path (conv_tc.cu).
#include "net.cuh"

namespace sn {

// ---------------------------------------------------------------------------------------------
// direct convolution: block = 4x8x8 output voxels (256 threads, 1 voxel each) x 16 output channels
constexpr int CV_TD = 4, CV_TH = 8, CV_TW = 8, CV_THREADS = 256;

__device__ __forceinline__ float apply_act(float y, int act) {
    if (act == SN_ACT_RELU) return fmaxf(y, 0.f);
    if (act == SN_ACT_SIGMOID) return 1.f / (1.f + expf(-y));
    return y;
}

template <int K>
__global__ void __launch_bounds__(CV_THREADS)
conv3d_fp32_kernel(const float* __restrict__ in, const float* __restrict__ wt, const float* __restrict__ scale,
                   const float* __restrict__ shift, float* __restrict__ out, int Cin, int Cin_pad, int Cout, int S, int dil,
                   int act, int C_total, int c_off) {
    constexpr int K3 = K * K * K;
    const int pad = dil * (K / 2);
    const int ED = CV_TD + 2 * pad, EH = CV_TH + 2 * pad, EW = CV_TW + 2 * pad;
    const int tile_vox = ED * EH * EW;
    extern __shared__ float smem[];
    float* s_in = smem;                                   // [CV_CI][ED][EH][EW]
    float* s_w = smem + CV_CI * tile_vox;                 // [CV_CI][K3][CV_COT]

    const int tiles_w = (S + CV_TW - 1) / CV_TW, tiles_h = (S + CV_TH - 1) / CV_TH, tiles_d = (S + CV_TD - 1) / CV_TD;
    int tile = blockIdx.x;
    const int tw0 = (tile % tiles_w) * CV_TW; tile /= tiles_w;
    const int th0 = (tile % tiles_h) * CV_TH; tile /= tiles_h;
    const int td0 = (tile % tiles_d) * CV_TD; tile /= tiles_d;
    const int n = tile;
    const int cog = blockIdx.y;
    const int tid = threadIdx.x;
    const int lw = tid % CV_TW, lh = (tid / CV_TW) % CV_TH, ld = tid / (CV_TW * CV_TH);
    const int64_t vol = (int64_t)S * S * S;

    float acc[CV_COT];
#pragma unroll
    for (int c = 0; c < CV_COT; ++c) acc[c] = 0.f;

    const float* in_n = in + (int64_t)n * Cin * vol;
    const float* wt_g = wt + (int64_t)cog * Cin_pad * K3 * CV_COT;

    for (int ci0 = 0; ci0 < Cin_pad; ci0 += CV_CI) {
        __syncthreads();
        // input halo tile, zero outside the volume ('same' padding) and for padded channels
        for (int idx = tid; idx < CV_CI * tile_vox; idx += CV_THREADS) {
            const int ci = idx / tile_vox;
            int r = idx - ci * tile_vox;
            const int ew = r % EW; r /= EW;
            const int eh = r % EH; const int ed = r / EH;
            const int gd = td0 + ed - pad, gh = th0 + eh - pad, gw = tw0 + ew - pad;
            float v = 0.f;
            if (ci0 + ci < Cin && gd >= 0 && gd < S && gh >= 0 && gh < S && gw >= 0 && gw < S)
                v = __ldg(in_n + (int64_t)(ci0 + ci) * vol + ((int64_t)gd * S + gh) * S + gw);
            s_in[idx] = v;
        }
        {
            const float4* src = reinterpret_cast<const float4*>(wt_g + (int64_t)ci0 * K3 * CV_COT);
            float4* dst = reinterpret_cast<float4*>(s_w);
            for (int idx = tid; idx < CV_CI * K3 * CV_COT / 4; idx += CV_THREADS) dst[idx] = __ldg(src + idx);
        }
        __syncthreads();
#pragma unroll 1
        for (int ci = 0; ci < CV_CI; ++ci) {
            const float* si = s_in + ci * tile_vox + ((ld * EH) + lh) * EW + lw;
            const float* sw = s_w + ci * K3 * CV_COT;
#pragma unroll
            for (int kd = 0; kd < K; ++kd)
#pragma unroll
                for (int kh = 0; kh < K; ++kh)
#pragma unroll
                    for (int kw = 0; kw < K; ++kw) {
                        const float x = si[((kd * dil) * EH + kh * dil) * EW + kw * dil];
                        const float4* w4 = reinterpret_cast<const float4*>(sw + ((kd * K + kh) * K + kw) * CV_COT);
#pragma unroll
                        for (int q = 0; q < CV_COT / 4; ++q) {
                            const float4 w = w4[q];
                            acc[4 * q + 0] = fmaf(x, w.x, acc[4 * q + 0]);
                            acc[4 * q + 1] = fmaf(x, w.y, acc[4 * q + 1]);
                            acc[4 * q + 2] = fmaf(x, w.z, acc[4 * q + 2]);
                            acc[4 * q + 3] = fmaf(x, w.w, acc[4 * q + 3]);
                        }
                    }
        }
    }
    const int gd = td0 + ld, gh = th0 + lh, gw = tw0 + lw;
    if (gd < S && gh < S && gw < S) {
        float* o = out + ((int64_t)n * C_total + c_off + cog * CV_COT) * vol + ((int64_t)gd * S + gh) * S + gw;
#pragma unroll
        for (int c = 0; c < CV_COT; ++c) {
            const int co = cog * CV_COT + c;
            if (co < Cout) o[(int64_t)c * vol] = apply_act(fmaf(acc[c], __ldg(scale + co), __ldg(shift + co)), act);
        }
    }
}

int conv_fp32_launch(const ConvUnit& u, const float* in, int n, int S, float* out, int C_total, int c_off, cudaStream_t st) {
    const int pad = u.dil * (u.K / 2);
    const int tile_vox = (CV_TD + 2 * pad) * (CV_TH + 2 * pad) * (CV_TW + 2 * pad);
    const int K3 = u.K * u.K * u.K;
    const size_t smem = (size_t)(CV_CI * tile_vox + CV_CI * K3 * CV_COT) * sizeof(float);
    const int64_t tiles = (int64_t)n * cdiv(S, CV_TD) * cdiv(S, CV_TH) * cdiv(S, CV_TW);
    SN_CHECK_ARG(tiles <= 0x7fffffff, "conv: too many tiles");
    dim3 grid((unsigned)tiles, (unsigned)cdiv(u.Cout, CV_COT));
    prof_begin(u.id, st);
    if (u.K == 3) {
        static bool attr_set = false;
        if (!attr_set) { SN_CUDA(cudaFuncSetAttribute(conv3d_fp32_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); attr_set = true; }
        conv3d_fp32_kernel<3><<<grid, CV_THREADS, smem, st>>>(in, u.w_fp32, u.scale, u.shift, out, u.Cin, u.Cin_pad, u.Cout, S, u.dil, u.act, C_total, c_off);
    } else {
        conv3d_fp32_kernel<1><<<grid, CV_THREADS, smem, st>>>(in, u.w_fp32, u.scale, u.shift, out, u.Cin, u.Cin_pad, u.Cout, S, u.dil, u.act, C_total, c_off);
    }
    prof_end(u.id, st);
    g_conv_path[0].fetch_add(1, std::memory_order_relaxed);
    SN_LAUNCHED();
    return SN_OK;
}

// ---------------------------------------------------------------------------------------------
__global__ void maxpool2_kernel(const float* __restrict__ in, int64_t total_out, int S, float* __restrict__ out) {
    const int So = S / 2;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total_out; i += stride) {
        const int w = (int)(i % So), h = (int)((i / So) % So), d = (int)((i / ((int64_t)So * So)) % So);
        const int64_t nc = i / ((int64_t)So * So * So);
        const float* p = in + nc * (int64_t)S * S * S + ((int64_t)(2 * d) * S + 2 * h) * S + 2 * w;
        float m = -INFINITY;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const float2 v = *reinterpret_cast<const float2*>(p + ((int64_t)a * S + b) * S);
                m = fmaxf(m, fmaxf(v.x, v.y));
            }
        out[i] = m;
    }
}

int maxpool2_launch(const float* in, int n, int C, int S, float* out, cudaStream_t st) {
    SN_CHECK_ARG(S % 2 == 0, "maxpool2: S=%d must be even", S);
    const int64_t total = (int64_t)n * C * (S / 2) * (S / 2) * (S / 2);
    if (total == 0) return SN_OK;
    const int blocks = (int)std::min<int64_t>(cdiv(total, 256), 148 * 32);
    maxpool2_kernel<<<blocks, 256, 0, st>>>(in, total, S, out);
    SN_LAUNCHED();
    return SN_OK;
}

// ---------------------------------------------------------------------------------------------
// out[o] = sum_t W[t] * stuffed[o + t - k/2],  stuffed[p] = in[p/f] when p % f == 0 in all 3 dims.
// Per dimension at most ceil(k/f) taps are non-zero: t = t0, t0+f, ... with t0 = (k/2 - o) mod f.
__global__ void upsample_kernel(const float* __restrict__ in, const float* __restrict__ W, int k, int f, int C, int S,
                                int64_t total_out, float* __restrict__ out, int C_total, int c_off) {
    const int So = S * f, c0 = k / 2;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total_out; i += stride) {
        const int ow = (int)(i % So), oh = (int)((i / So) % So), od = (int)((i / ((int64_t)So * So)) % So);
        const int64_t nc = i / ((int64_t)So * So * So);
        const int c = (int)(nc % C);
        const int64_t n = nc / C;
        const float* src = in + nc * (int64_t)S * S * S;
        const int td0 = ((c0 - od) % f + f) % f, th0 = ((c0 - oh) % f + f) % f, tw0 = ((c0 - ow) % f + f) % f;
        float acc = 0.f;
        for (int td = td0; td < k; td += f) {
            const int pd = od + td - c0;
            if (pd < 0 || pd >= So) continue;
            for (int th = th0; th < k; th += f) {
                const int ph = oh + th - c0;
                if (ph < 0 || ph >= So) continue;
                for (int tw = tw0; tw < k; tw += f) {
                    const int pw = ow + tw - c0;
                    if (pw < 0 || pw >= So) continue;
                    acc = fmaf(__ldg(W + (td * k + th) * k + tw), __ldg(src + ((int64_t)(pd / f) * S + ph / f) * S + pw / f), acc);
                }
            }
        }
        out[((n * C_total + c_off + c) * So + od) * (int64_t)So * So + (int64_t)oh * So + ow] = acc;
    }
}

int upsample_launch(const float* in, const float* W, int k, int f, int n, int C, int S, float* out, int C_total, int c_off, cudaStream_t st) {
    const int64_t total = (int64_t)n * C * S * f * S * f * S * f;
    if (total == 0) return SN_OK;
    const int blocks = (int)std::min<int64_t>(cdiv(total, 256), 148 * 32);
    upsample_kernel<<<blocks, 256, 0, st>>>(in, W, k, f, C, S, total, out, C_total, c_off);
    SN_LAUNCHED();
    return SN_OK;
}

// ---------------------------------------------------------------------------------------------
// fused[b, x] = sum_v (w[b,v] / sum_v w[b,v]) * p[b,v,x]      nets/layers.py:330-335
__global__ void fuse_kernel(const float* __restrict__ p, const float* __restrict__ w, int n_vp, int64_t vol, int64_t total,
                            float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int64_t b = i / vol, x = i - b * vol;
        const float* wb = w + b * n_vp;
        float sum = 0.f;
        for (int v = 0; v < n_vp; ++v) sum += wb[v];
        float acc = 0.f;
        for (int v = 0; v < n_vp; ++v) acc += p[(b * n_vp + v) * vol + x] * (wb[v] / sum);
        out[i] = acc;
    }
}

int fuse_launch(const float* p, const float* w, int n_cubes, int n_vp, int64_t vol, float* out, cudaStream_t st) {
    const int64_t total = (int64_t)n_cubes * vol;
    if (total == 0) return SN_OK;
    const int blocks = (int)std::min<int64_t>(cdiv(total, 256), 148 * 32);
    fuse_kernel<<<blocks, 256, 0, st>>>(p, w, n_vp, vol, total, out);
    SN_LAUNCHED();
    return SN_OK;
}

// ---------------------------------------------------------------------------------------------
// one block per feature row: h = sigmoid(BN(f . W1)); o = h . W2 + b        SurfaceNet.py:94-95
__global__ void relimp_mlp_kernel(const float* __restrict__ f, const float* __restrict__ W1, const float* __restrict__ scale,
                                  const float* __restrict__ shift, const float* __restrict__ W2, const float* __restrict__ b2,
                                  int Din, int H, float* __restrict__ logit) {
    extern __shared__ float sf[];
    float* sh = sf + Din;
    const int64_t row = blockIdx.x;
    for (int i = threadIdx.x; i < Din; i += blockDim.x) sf[i] = f[row * Din + i];
    __syncthreads();
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        float acc = 0.f;
        for (int i = 0; i < Din; ++i) acc = fmaf(sf[i], __ldg(W1 + (int64_t)i * H + j), acc);
        sh[j] = 1.f / (1.f + expf(-fmaf(acc, scale[j], shift[j]))) * __ldg(W2 + j);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float o = 0.f;
        for (int j = 0; j < H; ++j) o += sh[j];
        logit[row] = o + b2[0];
    }
}

// softmax over each group of n_per_group consecutive rows                      SurfaceNet.py:96-97
__global__ void group_softmax_kernel(const float* __restrict__ logit, int64_t groups, int g, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= groups) return;
    const float* l = logit + i * g;
    float m = -INFINITY;
    for (int j = 0; j < g; ++j) m = fmaxf(m, l[j]);
    float s = 0.f;
    for (int j = 0; j < g; ++j) s += expf(l[j] - m);
    for (int j = 0; j < g; ++j) out[i * g + j] = expf(l[j] - m) / s;
}

int relimp_launch(const Net& net, const float* features, int64_t n_rows, int n_per_group, float* logit_tmp, float* out, cudaStream_t st) {
    const int Din = 258, H = 100;
    relimp_mlp_kernel<<<(unsigned)n_rows, 128, (Din + H) * sizeof(float), st>>>(features, net.fc1_W, net.fc1_scale, net.fc1_shift, net.lin_W, net.lin_b, Din, H, logit_tmp);
    SN_LAUNCHED();
    const int64_t groups = n_rows / n_per_group;
    group_softmax_kernel<<<(unsigned)cdiv(groups, 128), 128, 0, st>>>(logit_tmp, groups, n_per_group, out);
    SN_LAUNCHED();
    return SN_OK;
}

}  // namespace sn
