// Parameter loading (nets/SurfaceNet.py:385-402 SurfaceNet_inference) and the forward graph
// orchestration (nets/SurfaceNet.py:18-76 + :126,343-357) for SN_MODE_FP32.
#include "net.cuh"
#include <string.h>

namespace sn {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
std::atomic<int64_t> g_conv_path[3];

void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- per-launch profiling (bench.py: roofline of the dominant kernel, measured live with CUDA events) ----
struct ProfRec { int unit; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_prof_pool;

static cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
void prof_begin(int unit, cudaStream_t st) {
    if (!g_prof_on) return;
    ProfRec r{unit, prof_event(), prof_event()};
    cudaEventRecord(r.a, st);
    g_prof.push_back(r);
}
void prof_end(int unit, cudaStream_t st) {
    if (!g_prof_on || g_prof.empty() || g_prof.back().unit != unit) return;
    cudaEventRecord(g_prof.back().b, st);
}

static int upload(const std::vector<float>& h, float** d) {
    SN_CUDA(cudaMalloc((void**)d, std::max<size_t>(h.size(), 1) * sizeof(float)));
    SN_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    return SN_OK;
}

static void fold_bn(const float* beta, const float* gamma, const float* mean, const float* inv_std, int C,
                    std::vector<float>& scale, std::vector<float>& shift) {
    scale.resize(C); shift.resize(C);
    for (int c = 0; c < C; ++c) {
        scale[c] = gamma[c] * inv_std[c];                 // (x - mean) * (gamma * inv_std) + beta
        shift[c] = beta[c] - mean[c] * scale[c];
    }
}

static int64_t expected_size(int i_array, int* unit_out, int* slot_out) {
    // walk the App. B layout
    int i = 0;
    for (int u = 0; u < kNumUnits; ++u) {
        const UnitSpec& s = kUnits[u];
        const int64_t k3 = (int64_t)s.K * s.K * s.K;
        if (s.kind == UNIT_UP) { if (i == i_array) { *unit_out = u; *slot_out = 0; return k3; } i += 1; continue; }
        if (i_array < i + 5) { *unit_out = u; *slot_out = i_array - i; return (i_array == i) ? (int64_t)s.Cin * s.Cout * k3 : s.Cout; }
        i += 5;
    }
    *unit_out = -1; *slot_out = i_array - i;
    static const int64_t tail[7] = {258 * 100, 100, 100, 100, 100, 100, 1};
    return (i_array - i < 7) ? tail[i_array - i] : -1;
}

static int first_array_of_unit(int unit) {
    int i = 0;
    for (int u = 0; u < unit; ++u) i += (kUnits[u].kind == UNIT_UP) ? 1 : 5;
    return i;
}

int tc_prepare(Net& net);     // conv_tc.cu
void tc_destroy(Net& net);

}  // namespace sn

using namespace sn;

extern "C" const char* sn_last_error(void) { return g_err; }
extern "C" int sn_version(void) { return 100; }
extern "C" int64_t sn_launch_count(void) { return g_launches.load(); }
extern "C" void sn_launch_count_reset(void) { g_launches.store(0); for (auto& c : g_conv_path) c.store(0); }
extern "C" void sn_conv_path_counts(int64_t counts[3]) { for (int i = 0; i < 3; ++i) counts[i] = g_conv_path[i].load(); }

extern "C" void sn_profile_enable(int on) {
    for (auto& r : g_prof) { g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b); }
    g_prof.clear();
    g_prof_on = on != 0;
}
extern "C" int sn_profile_collect(double* ms_per_unit, int64_t* launches_per_unit, int n_units) {
    SN_CHECK_ARG(ms_per_unit && launches_per_unit && n_units == kNumUnits, "sn_profile_collect: expects %d units", kNumUnits);
    for (int u = 0; u < n_units; ++u) { ms_per_unit[u] = 0; launches_per_unit[u] = 0; }
    for (auto& r : g_prof) {
        SN_CUDA(cudaEventSynchronize(r.b));
        float ms = 0.f;
        SN_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        ms_per_unit[r.unit] += ms; launches_per_unit[r.unit] += 1;
    }
    return SN_OK;
}

extern "C" int sn_net_create(const float* const* arrays_host, const int64_t* sizes, int n_arrays, sn_net** out) {
    SN_CHECK_ARG(arrays_host && sizes && out, "sn_net_create: NULL argument");
    SN_CHECK_ARG(n_arrays == 105, "sn_net_create: the SurfaceNet parameter list holds 105 arrays, got %d", n_arrays);
    for (int i = 0; i < n_arrays; ++i) {
        int u, s;
        const int64_t e = expected_size(i, &u, &s);
        SN_CHECK_ARG(arrays_host[i] && sizes[i] == e, "sn_net_create: parameter %d has %lld elements, expected %lld", i, (long long)sizes[i], (long long)e);
    }
    sn_net* h = new sn_net();
    Net& net = h->net;
    int rc = SN_OK;
    for (int u = 0; u < kNumUnits && rc == SN_OK; ++u) {
        const UnitSpec& s = kUnits[u];
        ConvUnit& cu = net.units[u];
        const int i0 = first_array_of_unit(u);
        const int K3 = s.K * s.K * s.K;
        cu.id = u; cu.kind = s.kind; cu.Cin = s.Cin; cu.Cout = s.Cout; cu.K = s.K;
        cu.dil = (s.kind == UNIT_DIL) ? 2 : 1;
        if (s.kind == UNIT_UP) {
            std::vector<float> w(arrays_host[i0], arrays_host[i0] + K3);
            rc = upload(w, &cu.up_W);
            continue;
        }
        const bool sig = (u == U_SIDE1 || u == U_SIDE2 || u == U_SIDE3 || u == U_SIDE4 || u == U_MERGE3);
        cu.act = sig ? SN_ACT_SIGMOID : SN_ACT_RELU;
        cu.Cin_pad = (int)align_up(s.Cin, CV_CI);
        // canonical (Cout, Cin, K3); DilatedConv3DLayer stores (Cin, Cout, k,k,k) (nets/layers.py:200-253)
        cu.h_w.resize((size_t)s.Cout * s.Cin * K3);
        const float* W = arrays_host[i0];
        for (int co = 0; co < s.Cout; ++co)
            for (int ci = 0; ci < s.Cin; ++ci)
                for (int t = 0; t < K3; ++t)
                    cu.h_w[((size_t)co * s.Cin + ci) * K3 + t] =
                        (s.kind == UNIT_DIL) ? W[((size_t)ci * s.Cout + co) * K3 + t] : W[((size_t)co * s.Cin + ci) * K3 + t];
        fold_bn(arrays_host[i0 + 1], arrays_host[i0 + 2], arrays_host[i0 + 3], arrays_host[i0 + 4], s.Cout, cu.h_scale, cu.h_shift);
        const int groups = (int)cdiv(s.Cout, CV_COT);
        std::vector<float> wl((size_t)groups * cu.Cin_pad * K3 * CV_COT, 0.f);
        for (int co = 0; co < s.Cout; ++co)
            for (int ci = 0; ci < s.Cin; ++ci)
                for (int t = 0; t < K3; ++t)
                    wl[(((size_t)(co / CV_COT) * cu.Cin_pad + ci) * K3 + t) * CV_COT + co % CV_COT] = cu.h_w[((size_t)co * s.Cin + ci) * K3 + t];
        if ((rc = upload(wl, &cu.w_fp32)) != SN_OK) break;
        if ((rc = upload(cu.h_scale, &cu.scale)) != SN_OK) break;
        rc = upload(cu.h_shift, &cu.shift);
    }
    if (rc == SN_OK) {
        const int i0 = first_array_of_unit(kNumUnits);           // 98
        std::vector<float> sc, sh;
        fold_bn(arrays_host[i0 + 1], arrays_host[i0 + 2], arrays_host[i0 + 3], arrays_host[i0 + 4], 100, sc, sh);
        rc = upload(std::vector<float>(arrays_host[i0], arrays_host[i0] + 258 * 100), &net.fc1_W);
        if (rc == SN_OK) rc = upload(sc, &net.fc1_scale);
        if (rc == SN_OK) rc = upload(sh, &net.fc1_shift);
        if (rc == SN_OK) rc = upload(std::vector<float>(arrays_host[i0 + 5], arrays_host[i0 + 5] + 100), &net.lin_W);
        if (rc == SN_OK) rc = upload(std::vector<float>(arrays_host[i0 + 6], arrays_host[i0 + 6] + 1), &net.lin_b);
    }
    if (rc == SN_OK) rc = upload(std::vector<float>{123.68f, 116.779f, 103.939f, 123.68f, 116.779f, 103.939f}, &net.mean6);
    if (rc == SN_OK) rc = tc_prepare(net);
    if (rc != SN_OK) { sn_net_destroy(h); return rc; }
    *out = h;
    return SN_OK;
}

extern "C" void sn_net_destroy(sn_net* h) {
    if (!h) return;
    Net& net = h->net;
    tc_destroy(net);
    for (int u = 0; u < kNumUnits; ++u) {
        cudaFree(net.units[u].w_fp32); cudaFree(net.units[u].scale); cudaFree(net.units[u].shift); cudaFree(net.units[u].up_W);
    }
    cudaFree(net.mean6);
    cudaFree(net.fc1_W); cudaFree(net.fc1_scale); cudaFree(net.fc1_shift); cudaFree(net.lin_W); cudaFree(net.lin_b);
    delete h;
}

// ---- single layers -------------------------------------------------------------------------------
namespace sn { int tc_layer_conv(const Net& net, int u, const float* in, int n, int S, float* out, int mode, cudaStream_t st); }

extern "C" int sn_net_layer_conv(const sn_net* h, int unit, const float* in_dev, int n, int S, float* out_dev, int mode, void* stream) {
    SN_CHECK_ARG(h && in_dev && out_dev, "sn_net_layer_conv: NULL argument");
    SN_CHECK_ARG(unit >= 0 && unit < kNumUnits && kUnits[unit].kind != UNIT_UP, "sn_net_layer_conv: unit %d is not a conv unit", unit);
    SN_CHECK_ARG(n >= 0 && S >= 1, "sn_net_layer_conv: bad sizes");
    SN_CHECK_ARG(mode == SN_MODE_FP32 || mode == SN_MODE_TC_EXACT || mode == SN_MODE_TC_FAST, "sn_net_layer_conv: unknown mode %d", mode);
    if (n == 0) return SN_OK;
    if (mode != SN_MODE_FP32) return tc_layer_conv(h->net, unit, in_dev, n, S, out_dev, mode, (cudaStream_t)stream);
    return conv_fp32_launch(h->net.units[unit], in_dev, n, S, out_dev, kUnits[unit].Cout, 0, (cudaStream_t)stream);
}

extern "C" int sn_maxpool2(const float* in_dev, int n, int C, int S, float* out_dev, void* stream) {
    SN_CHECK_ARG(in_dev && out_dev && n >= 0 && C >= 1 && S >= 2, "sn_maxpool2: bad arguments");
    return maxpool2_launch(in_dev, n, C, S, out_dev, (cudaStream_t)stream);
}

extern "C" int sn_net_layer_upsample(const sn_net* h, int unit, const float* in_dev, int n, int C, int S, float* out_dev,
                                     int C_total, int c_off, void* stream) {
    SN_CHECK_ARG(h && in_dev && out_dev, "sn_net_layer_upsample: NULL argument");
    SN_CHECK_ARG(unit >= 0 && unit < kNumUnits && kUnits[unit].kind == UNIT_UP, "sn_net_layer_upsample: unit %d is not an up-sampling unit", unit);
    SN_CHECK_ARG(n >= 0 && C >= 1 && S >= 1 && c_off >= 0 && c_off + C <= C_total, "sn_net_layer_upsample: bad sizes");
    const int k = kUnits[unit].K, f = (k == 3) ? 2 : 4;      // k = f/2*2+1 (nets/layers.py:383)
    return upsample_launch(in_dev, h->net.units[unit].up_W, k, f, n, C, S, out_dev, C_total, c_off, (cudaStream_t)stream);
}

extern "C" int sn_fuse_weighted_average(const float* p_dev, const float* w_dev, int n_cubes, int n_vp, int64_t vol, float* out_dev, void* stream) {
    SN_CHECK_ARG(p_dev && w_dev && out_dev && n_cubes >= 0 && n_vp >= 1 && vol >= 1, "sn_fuse_weighted_average: bad arguments");
    return fuse_launch(p_dev, w_dev, n_cubes, n_vp, vol, out_dev, (cudaStream_t)stream);
}

extern "C" int sn_net_relative_importance(const sn_net* h, const float* features_dev, int64_t n_rows, int n_per_group, float* out_dev, void* stream) {
    SN_CHECK_ARG(h && features_dev && out_dev, "sn_net_relative_importance: NULL argument");
    SN_CHECK_ARG(n_rows >= 0 && n_per_group >= 1 && n_rows % n_per_group == 0, "sn_net_relative_importance: %lld rows is not a multiple of n_samples_perGroup=%d", (long long)n_rows, n_per_group);
    if (n_rows == 0) return SN_OK;
    float* tmp = nullptr;
    SN_CUDA(cudaMallocAsync((void**)&tmp, n_rows * sizeof(float), (cudaStream_t)stream));
    int rc = relimp_launch(h->net, features_dev, n_rows, n_per_group, tmp, out_dev, (cudaStream_t)stream);
    cudaFreeAsync(tmp, (cudaStream_t)stream);
    return rc;
}

// ---- whole forward, fp32 -------------------------------------------------------------------------
namespace sn {

// floats of workspace per pair-cube (V = D^3)
static int64_t fp32_floats_per_pc(int D) {
    const int64_t V = (int64_t)D * D * D, V2 = V / 8, V4 = V / 64;
    return 32 * V * 2 + 32 * V2 + 80 * V2 * 2 + 16 * V2 + 80 * V4 + 160 * V4 * 2 + 300 * V4 * 2 + 16 * V4 + 64 * V + 100 * V * 2 + 64;
}

int tc_forward(const Net& net, const float* X, int n_pc, int D, float* prob_out, void* ws, int64_t ws_bytes, int mode, cudaStream_t st,
               const CvcSource* src);  // conv_tc.cu
int64_t tc_workspace_bytes(const Net& net, int n_pc, int D, int mode);

static int fp32_forward_chunk(const Net& net, const float* X, int n, int D, float* prob_out, float* ws, cudaStream_t st) {
    const int64_t V = (int64_t)D * D * D, V2 = V / 8, V4 = V / 64;
    const int S1 = D, S2 = D / 2, S4 = D / 4;
    float* a1 = ws;               float* a2 = a1 + 32 * V * n;
    float* p1 = a2 + 32 * V * n;  float* b1 = p1 + 32 * V2 * n;   float* b2 = b1 + 80 * V2 * n;
    float* s2 = b2 + 80 * V2 * n; float* p2 = s2 + 16 * V2 * n;   float* c1 = p2 + 80 * V4 * n;
    float* c2 = c1 + 160 * V4 * n; float* d1 = c2 + 160 * V4 * n; float* d2 = d1 + 300 * V4 * n;
    float* s4 = d2 + 300 * V4 * n; float* cat = s4 + 16 * V4 * n; float* m1 = cat + 64 * V * n; float* m2 = m1 + 100 * V * n;
    const ConvUnit* U = net.units;
    int rc;
#define RUN(x) do { rc = (x); if (rc != SN_OK) return rc; } while (0)
    RUN(conv_fp32_launch(U[U_CONV1_1], X, n, S1, a1, 32, 0, st));
    RUN(conv_fp32_launch(U[U_CONV1_2], a1, n, S1, a2, 32, 0, st));
    RUN(conv_fp32_launch(U[U_CONV1_3], a2, n, S1, a1, 32, 0, st));
    RUN(conv_fp32_launch(U[U_SIDE1], a1, n, S1, cat, 64, 0, st));                       // side_op1 -> concat[0:16]
    RUN(maxpool2_launch(a1, n, 32, S1, p1, st));
    RUN(conv_fp32_launch(U[U_CONV2_1], p1, n, S2, b1, 80, 0, st));
    RUN(conv_fp32_launch(U[U_CONV2_2], b1, n, S2, b2, 80, 0, st));
    RUN(conv_fp32_launch(U[U_CONV2_3], b2, n, S2, b1, 80, 0, st));
    RUN(conv_fp32_launch(U[U_SIDE2], b1, n, S2, s2, 16, 0, st));
    RUN(upsample_launch(s2, U[U_UP2].up_W, 3, 2, n, 16, S2, cat, 64, 16, st));         // -> concat[16:32]
    RUN(maxpool2_launch(b1, n, 80, S2, p2, st));
    RUN(conv_fp32_launch(U[U_CONV3_1], p2, n, S4, c1, 160, 0, st));
    RUN(conv_fp32_launch(U[U_CONV3_2], c1, n, S4, c2, 160, 0, st));
    RUN(conv_fp32_launch(U[U_CONV3_3], c2, n, S4, c1, 160, 0, st));
    RUN(conv_fp32_launch(U[U_SIDE3], c1, n, S4, s4, 16, 0, st));
    RUN(upsample_launch(s4, U[U_UP3].up_W, 5, 4, n, 16, S4, cat, 64, 32, st));         // -> concat[32:48]
    RUN(conv_fp32_launch(U[U_CONV4_1], c1, n, S4, d1, 300, 0, st));
    RUN(conv_fp32_launch(U[U_CONV4_2], d1, n, S4, d2, 300, 0, st));
    RUN(conv_fp32_launch(U[U_CONV4_3], d2, n, S4, d1, 300, 0, st));
    RUN(conv_fp32_launch(U[U_SIDE4], d1, n, S4, s4, 16, 0, st));
    RUN(upsample_launch(s4, U[U_UP4].up_W, 5, 4, n, 16, S4, cat, 64, 48, st));         // -> concat[48:64]
    RUN(conv_fp32_launch(U[U_MERGE1], cat, n, S1, m1, 100, 0, st));
    RUN(conv_fp32_launch(U[U_MERGE2], m1, n, S1, m2, 100, 0, st));
    RUN(conv_fp32_launch(U[U_MERGE3], m2, n, S1, prob_out, 1, 0, st));
#undef RUN
    return SN_OK;
}

constexpr int kFp32MaxChunk = 8;

}  // namespace sn

extern "C" int64_t sn_net_workspace_bytes(const sn_net* h, int n_pair_cubes, int D, int mode) {
    if (!h || n_pair_cubes < 0 || D < 4 || D % 4) return -1;
    const int64_t V = (int64_t)D * D * D;
    const int64_t unf = align_up((int64_t)n_pair_cubes * V * 4, 256);       // unfused probabilities when the caller passes NULL
    if (mode == SN_MODE_FP32)
        return unf + align_up(fp32_floats_per_pc(D) * 4 * std::min(n_pair_cubes, kFp32MaxChunk), 256) + 256;
    const int64_t t = tc_workspace_bytes(h->net, n_pair_cubes, D, mode);
    return t < 0 ? t : unf + t + 256;
}

extern "C" int sn_net_forward(const sn_net* h, const float* X_dev, int n_cubes, int n_vp, int D, const float* w_dev,
                              float* fused_out_dev, float* unfused_out_dev, void* workspace_dev, int64_t workspace_bytes,
                              int mode, void* stream) {
    SN_CHECK_ARG(h && X_dev, "sn_net_forward: NULL argument");
    return sn::net_forward(h, X_dev, nullptr, n_cubes, n_vp, D, w_dev, fused_out_dev, unfused_out_dev, workspace_dev, workspace_bytes, mode, stream);
}

// the network input is either the fp32 CVC tensor X_dev or, for the forwards tc_gathers_directly() names, the gather's arguments (src)
int sn::net_forward(const sn_net* h, const float* X_dev, const CvcSource* src, int n_cubes, int n_vp, int D, const float* w_dev, float* fused_out_dev,
                    float* unfused_out_dev, void* workspace_dev, int64_t workspace_bytes, int mode, void* stream) {
    SN_CHECK_ARG(h && (X_dev || src) && fused_out_dev, "sn_net_forward: NULL argument");
    SN_CHECK_ARG(!src || tc_gathers_directly(h->net, D, mode), "sn_net_forward: this mode / cube size needs the CVC tensor");
    SN_CHECK_ARG(n_cubes >= 0 && n_vp >= 1, "sn_net_forward: bad sizes (n_cubes=%d n_vp=%d)", n_cubes, n_vp);
    SN_CHECK_ARG(D >= 4 && D % 4 == 0, "sn_net_forward: cube side %d must be a multiple of 4 (two 2^3 poolings)", D);
    SN_CHECK_ARG(w_dev || n_vp == 1, "sn_net_forward: w is required when N_viewPairs4inference >= 2 (nets/SurfaceNet.py:343-347)");
    SN_CHECK_ARG(mode == SN_MODE_FP32 || mode == SN_MODE_TC_EXACT || mode == SN_MODE_TC_FAST, "sn_net_forward: unknown mode %d", mode);
    if (n_cubes == 0) return SN_OK;
    const int n_pc = n_cubes * n_vp;
    const int64_t V = (int64_t)D * D * D;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t need = sn_net_workspace_bytes(h, n_pc, D, mode);
    if (need < 0) return SN_ERR_INVALID;
    if (!workspace_dev || workspace_bytes < need) { set_error("sn_net_forward: workspace %lld B < %lld B", (long long)workspace_bytes, (long long)need); return SN_ERR_NOMEM; }
    Arena ar(workspace_dev, workspace_bytes);
    float* unf_ws = ar.take<float>((int64_t)n_pc * V);
    // N_vp == 1: fused and unfused are the same tensor (SurfaceNet.py:354-357)
    float* prob = (n_vp == 1) ? fused_out_dev : (unfused_out_dev ? unfused_out_dev : unf_ws);
    int rc = SN_OK;
    if (mode == SN_MODE_FP32) {
        const int chunk = std::min(n_pc, kFp32MaxChunk);
        float* ws = ar.take<float>(fp32_floats_per_pc(D) * chunk);
        for (int i = 0; i < n_pc && rc == SN_OK; i += chunk) {
            const int n = std::min(chunk, n_pc - i);
            rc = fp32_forward_chunk(h->net, X_dev + (int64_t)i * 6 * V, n, D, prob + (int64_t)i * V, ws, st);
        }
    } else {
        void* ws = workspace_dev ? (char*)workspace_dev + ar.off : nullptr;
        rc = tc_forward(h->net, X_dev, n_pc, D, prob, ws, workspace_bytes - ar.off, mode, st, src);
    }
    if (rc != SN_OK) return rc;
    if (n_vp == 1) {
        if (unfused_out_dev && unfused_out_dev != fused_out_dev)
            SN_CUDA(cudaMemcpyAsync(unfused_out_dev, fused_out_dev, (int64_t)n_pc * V * 4, cudaMemcpyDeviceToDevice, st));
        return SN_OK;
    }
    return fuse_launch(prob, w_dev, n_cubes, n_vp, V, fused_out_dev, st);
}
