// Host-side tables of the tensor-core path (weights in operand layout, folded BatchNorm, tuned tile configurations).
#pragma once
#include "net.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <map>

namespace sn {

constexpr int TC_MAX_NT = 4;

struct TileCfg { int AD, NB, persist, tps; };          // d-planes per CTA, weight-ring depth, CTA scheduling (see conv_tc_launch_cfg)
struct TcVariant {                                 // [0] exact (P = 2), [1] fast (P = 1): own N tiling and weight image
    int n_ntiles = 0, nt_size[TC_MAX_NT] = {0, 0, 0, 0}, nt_off[TC_MAX_NT] = {0, 0, 0, 0}, nt_nc[TC_MAX_NT] = {0, 0, 0, 0};
    long long nt_woff[TC_MAX_NT] = {0, 0, 0, 0};
    unsigned char* w = nullptr;
};
struct TcUnit {
    int Cin_pad = 0, Cout_pad = 0, taps = 0, pair_last = 0, nv_last = 0;
    float* side_w = nullptr;                       // for side units with one-N-tile producers: [Cin_pad][16] fp32 (transposed)
    TcVariant v[2];
    float* scale = nullptr;                        // Cout_pad entries, zero for padded channels
    float* shift = nullptr;
};

// w-axis Winograd F(2,3) variant of a 3x3x3 unit (conv_wg.cu): per frequency f = 0..3 the weights U_f = G g over kw, in the same
// per-(channel block, tap) operand stages as the direct kernel, taps = the 9 (kd, kh) pairs
struct WgUnit {
    bool on = false;
    int Cin_pad = 0, Cout_pad = 0, n_cblk = 0, pair_last = 0, stages_per_f = 0;
    int n_ntiles = 0, nt_size[TC_MAX_NT] = {0, 0, 0, 0}, nt_off[TC_MAX_NT] = {0, 0, 0, 0}, nt_nc[TC_MAX_NT] = {0, 0, 0, 0};
    long long nt_woff[TC_MAX_NT] = {0, 0, 0, 0};   // byte offset of the N tile's weights; one frequency = stages_per_f * 2*nt_nc*32 bytes
    unsigned char* w = nullptr;
    float* scale = nullptr;                        // folded BatchNorm with this variant's round-toward-zero compensation
    float* shift = nullptr;
};

struct TcState {
    TcUnit units[kNumUnits];
    WgUnit wg[kNumUnits];
    float* w3 = nullptr; float scale3 = 0.f, shift3 = 0.f;
    PFN_cuTensorMapEncodeTiled encode = nullptr;
    cudaStream_t side_stream = nullptr;            // the side-output branch (side convs + up-samplers) runs beside the main chain
    cudaEvent_t side_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    std::map<int, std::pair<TileCfg, long long>> tuned;     // (unit, mode, S) -> best tile configuration, work it was measured at
};

int tc_get_encode(TcState* st);                    // conv_tc.cu: resolves cuTensorMapEncodeTiled once

// conv_wg.cu
constexpr int WG_OUT_RAW = 0, WG_OUT_WINO = 1, WG_OUT_FINAL = 2;
int wg_prepare(Net& net);
void wg_destroy(Net& net);
bool wg_supported(const Net& net, int u, int S);
// in: Winograd-domain blk tensor (n_pc, 2, 4, Cin_pad/8, S, S, S/2, 8) fp16.  out_fmt RAW: blk (n_pc, 2, cg_total, S^3, 8) at group cg_off;
// WINO: Winograd-domain blk of the next unit; FINAL: fused merge_conv3 + sigmoid -> prob_out fp32 (n_pc, S^3)
int wg_conv_launch(const Net& net, int u, const __half* in_wino, int n_pc, int S, int out_fmt, __half* out, int cg_out_total, int cg_out_off,
                   float* prob_out, cudaStream_t stream, int cg_in_total = 0);      // cg_in_total: groups of the input tensor when > Cin_pad/8
// blk raw (n, 2, cg_in_total, S^3, 8) groups [cg_in_off, +cg_count) -> Winograd-domain blk (n, 2, 4, cg_out_total, S, S, S/2, 8) at cg_out_off;
// dil = 2: the pairs of a dilated consumer, t = 2j + parity <-> voxels (parity + 4j, parity + 4j + 2)
int raw_to_wino_launch(const __half* in_raw, int n, int cg_in_total, int cg_in_off, int cg_count, int S, __half* out_wino, int cg_out_total,
                       int cg_out_off, cudaStream_t stream, int dil = 1);

// fp32 NCDHW input (n, C <= 8, S^3) -> group 0 of a 2-group Winograd-domain tensor (conv1_1's operand)
int pack_wino_launch(const float* x, int n, int C, int S, __half* out_wino, cudaStream_t stream);
// the same operand for pair-cubes [pc0, pc0 + n) coloured straight from the images: CVC gather + mean subtraction + input transform in one pass
int cvc_wino_launch(const CvcSource& src, int pc0, int n, int S, __half* out_wino, cudaStream_t stream);

}  // namespace sn
