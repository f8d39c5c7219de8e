// K1 -- Colored Voxel Cube construction (utils/CVC.py:6-53,56-104,108-111) and
// perspectiveProj (utils/camera.py:123-184) on the device.
//
// One thread colours VPT consecutive voxels (z fastest) of one (cube, view slot): fp64 projection,
// rint, int32 index, bounds mask, 3-byte nearest-pixel gather, optional mean subtraction, and
// 16-byte coalesced stores into the reference layout (B*N_vp, 6, D,D,D) f32.
// HBM-bound: 24 B written + <= 6 B gathered per pair-voxel (DESIGN.md "K1").
#include "geometry.cuh"

namespace sn {

constexpr int CVC_THREADS = 256;
constexpr int CVC_VPT = 4;   // voxels per thread -> float4 stores

template <bool WRITE_X, bool WRITE_IDX>
__global__ void __launch_bounds__(CVC_THREADS)
cvc_gather_kernel(const uint8_t* __restrict__ images, const int64_t* __restrict__ img_offset,
                  const int32_t* __restrict__ img_hw, int n_views, const double* __restrict__ P,
                  const float* __restrict__ xyz, const float* __restrict__ resol,
                  const int32_t* __restrict__ views, int n_vp, int D, const float* __restrict__ mean6,
                  float* __restrict__ X, int32_t* __restrict__ idx_w, int32_t* __restrict__ idx_h,
                  uint8_t* __restrict__ in_scope) {
    const int slot = blockIdx.y;                 // (cube, view slot) = b * 2*n_vp + v
    const int b = slot / (2 * n_vp);
    const int v = slot - b * 2 * n_vp;
    const int64_t vol = (int64_t)D * D * D;
    const int64_t n0 = ((int64_t)blockIdx.x * CVC_THREADS + threadIdx.x) * CVC_VPT;
    if (n0 >= vol) return;

    const int view = views[slot];
    const bool view_ok = (view >= 0 && view < n_views);
    double Pm[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) Pm[i] = view_ok ? __ldg(P + (int64_t)view * 12 + i) : 0.0;
    const int H = view_ok ? img_hw[2 * view] : 0;
    const int W = view_ok ? img_hw[2 * view + 1] : 0;
    const uint8_t* __restrict__ img = images + (view_ok ? img_offset[view] : 0);
    const float rs = resol[b];
    const float x0 = xyz[3 * b], y0 = xyz[3 * b + 1], z0 = xyz[3 * b + 2];

    float rgb[3][CVC_VPT];
    int32_t wv[CVC_VPT], hv[CVC_VPT];
    uint8_t inv[CVC_VPT];
#pragma unroll
    for (int e = 0; e < CVC_VPT; ++e) {
        const int64_t n = n0 + e;
        const int k = (int)(n % D);
        const int j = (int)((n / D) % D);
        const int i = (int)(n / ((int64_t)D * D));
        const Proj pr = project(Pm, voxel_coord(i, rs, x0), voxel_coord(j, rs, y0), voxel_coord(k, rs, z0));
        const int32_t w = round_to_i32(__ddiv_rn(pr.u, pr.q));      // CVC.py:38-39: [w, h, 1]
        const int32_t h = round_to_i32(__ddiv_rn(pr.t, pr.q));
        const bool in = (n < vol) && (w < W) && (h < H) && (w >= 0) && (h >= 0);   // CVC.py:45
        wv[e] = w; hv[e] = h; inv[e] = in ? 1 : 0;
        float r = 0.f, g = 0.f, bl = 0.f;                                           // CVC.py:42
        if (in) {
            const uint8_t* px = img + ((int64_t)h * W + w) * 3;                     // CVC.py:46 img[h, w]
            r = (float)__ldg(px); g = (float)__ldg(px + 1); bl = (float)__ldg(px + 2);
        }
        rgb[0][e] = r; rgb[1][e] = g; rgb[2][e] = bl;
    }

    const bool full = (n0 + CVC_VPT <= vol) && (vol % CVC_VPT == 0);
    if (WRITE_X) {
        const int64_t pc = (int64_t)b * n_vp + v / 2;       // (N_cubes, N_vp*2, 3, ...) -> (N_cubes*N_vp, 6, ...): CVC.py:67,104
        const int ch0 = 3 * (v & 1);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float m = mean6 ? __ldg(mean6 + ch0 + c) : 0.f;
            float* dst = X + (pc * 6 + ch0 + c) * vol + n0;
            if (full) {
                *reinterpret_cast<float4*>(dst) = make_float4(rgb[c][0] - m, rgb[c][1] - m, rgb[c][2] - m, rgb[c][3] - m);
            } else {
                for (int e = 0; e < CVC_VPT && n0 + e < vol; ++e) dst[e] = rgb[c][e] - m;
            }
        }
    }
    if (WRITE_IDX) {
        const int64_t o = (int64_t)slot * vol + n0;
        for (int e = 0; e < CVC_VPT && n0 + e < vol; ++e) {
            idx_w[o + e] = wv[e]; idx_h[o + e] = hv[e]; in_scope[o + e] = inv[e];
        }
    }
}

__global__ void sub_channel_mean_kernel(float* __restrict__ X, int64_t total, int channels, int64_t spatial,
                                        const float* __restrict__ mean) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int c = (int)((i / spatial) % channels);
        X[i] -= __ldg(mean + c);
    }
}

__global__ void perspective_proj_kernel(const double* __restrict__ P, int n_mats, const double* __restrict__ xyz,
                                        int64_t n_pts, int round_to_int, double* __restrict__ h_out,
                                        double* __restrict__ w_out, double* __restrict__ depth_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (i >= n_pts) return;
    const Proj pr = project(P + (int64_t)m * 12, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    double w = __ddiv_rn(pr.u, pr.q), h = __ddiv_rn(pr.t, pr.q);      // camera.py:177
    if (round_to_int) { w = rint(w); h = rint(h); }                   // camera.py:179
    h_out[(int64_t)m * n_pts + i] = h;
    w_out[(int64_t)m * n_pts + i] = w;
    if (depth_out) depth_out[(int64_t)m * n_pts + i] = pr.q;          // camera.py:182
}

}  // namespace sn

using namespace sn;

extern "C" int sn_cvc_gather(const uint8_t* images_dev, const int64_t* img_offset_dev, const int32_t* img_hw_dev,
                             int n_views, const double* P_dev, const float* xyz_dev, const float* resol_dev,
                             const int32_t* views_dev, int n_cubes, int n_vp, int D, const float* mean6_dev,
                             float* X_out_dev, int32_t* idx_w_out_dev, int32_t* idx_h_out_dev,
                             uint8_t* in_scope_out_dev, void* stream) {
    SN_CHECK_ARG(n_cubes >= 0 && n_vp >= 1 && D >= 1 && n_views >= 1, "sn_cvc_gather: bad sizes (n_cubes=%d n_vp=%d D=%d n_views=%d)", n_cubes, n_vp, D, n_views);
    SN_CHECK_ARG(images_dev && img_offset_dev && img_hw_dev && P_dev && xyz_dev && resol_dev && views_dev, "sn_cvc_gather: NULL input");
    const bool idx = idx_w_out_dev || idx_h_out_dev || in_scope_out_dev;
    SN_CHECK_ARG(!idx || (idx_w_out_dev && idx_h_out_dev && in_scope_out_dev), "sn_cvc_gather: index-map outputs must be all set or all NULL");
    SN_CHECK_ARG(X_out_dev || idx, "sn_cvc_gather: no output requested");
    if (n_cubes == 0) return SN_OK;
    SN_CHECK_ARG((int64_t)n_cubes * 2 * n_vp <= 65535, "sn_cvc_gather: n_cubes*2*n_vp = %lld exceeds 65535 per call", (long long)n_cubes * 2 * n_vp);
    const int64_t vol = (int64_t)D * D * D;
    dim3 grid((unsigned)cdiv(vol, CVC_THREADS * CVC_VPT), (unsigned)(n_cubes * 2 * n_vp));
    cudaStream_t st = (cudaStream_t)stream;
#define SN_CVC_LAUNCH(WX, WI) cvc_gather_kernel<WX, WI><<<grid, CVC_THREADS, 0, st>>>(images_dev, img_offset_dev, img_hw_dev, n_views, P_dev, xyz_dev, resol_dev, views_dev, n_vp, D, mean6_dev, X_out_dev, idx_w_out_dev, idx_h_out_dev, in_scope_out_dev)
    if (X_out_dev && idx) SN_CVC_LAUNCH(true, true);
    else if (X_out_dev) SN_CVC_LAUNCH(true, false);
    else SN_CVC_LAUNCH(false, true);
#undef SN_CVC_LAUNCH
    SN_LAUNCHED();
    return SN_OK;
}

extern "C" int sn_sub_channel_mean(float* X_dev, int64_t n, int channels, int64_t spatial, const float* mean_dev, void* stream) {
    SN_CHECK_ARG(X_dev && mean_dev && n >= 0 && channels >= 1 && spatial >= 1, "sn_sub_channel_mean: bad arguments");
    const int64_t total = n * channels * spatial;
    if (total == 0) return SN_OK;
    const int blocks = (int)std::min<int64_t>(cdiv(total, 256), 148 * 16);
    sub_channel_mean_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(X_dev, total, channels, spatial, mean_dev);
    SN_LAUNCHED();
    return SN_OK;
}

extern "C" int sn_perspective_proj(const double* P_dev, int n_mats, const double* xyz_dev, int64_t n_pts, int round_to_int,
                                   double* h_out_dev, double* w_out_dev, double* depth_out_dev, void* stream) {
    SN_CHECK_ARG(P_dev && xyz_dev && h_out_dev && w_out_dev, "sn_perspective_proj: NULL argument");
    SN_CHECK_ARG(n_mats >= 1 && n_mats <= 65535 && n_pts >= 0, "sn_perspective_proj: bad sizes (n_mats=%d n_pts=%lld)", n_mats, (long long)n_pts);
    if (n_pts == 0) return SN_OK;
    dim3 grid((unsigned)cdiv(n_pts, 256), (unsigned)n_mats);
    perspective_proj_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P_dev, n_mats, xyz_dev, n_pts, round_to_int, h_out_dev, w_out_dev, depth_out_dev);
    SN_LAUNCHED();
    return SN_OK;
}
