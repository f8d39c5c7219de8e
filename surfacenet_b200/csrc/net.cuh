// SurfaceNet parameter store + launch helpers shared by the fp32 and tensor-core paths.
#pragma once
#include "common.cuh"
#include <algorithm>
#include <vector>

namespace sn {

constexpr int CV_COT = 16;   // output channels per block (fp32 direct conv)
constexpr int CV_CI = 8;     // input channels per shared-memory stage

enum UnitKind { UNIT_CONV = 0, UNIT_DIL = 1, UNIT_UP = 2 };

struct UnitSpec { const char* name; int kind, Cin, Cout, K; };
// mirrors surfacenet_b200/weights.py:UNITS (nets/SurfaceNet.py:33-74)
static const UnitSpec kUnits[] = {
    {"conv1_1", UNIT_CONV, 6, 32, 3},    {"conv1_2", UNIT_CONV, 32, 32, 3},   {"conv1_3", UNIT_CONV, 32, 32, 3},
    {"side_op1", UNIT_CONV, 32, 16, 1},
    {"conv2_1", UNIT_CONV, 32, 80, 3},   {"conv2_2", UNIT_CONV, 80, 80, 3},   {"conv2_3", UNIT_CONV, 80, 80, 3},
    {"side_op2", UNIT_CONV, 80, 16, 1},  {"up2", UNIT_UP, 1, 1, 3},
    {"conv3_1", UNIT_CONV, 80, 160, 3},  {"conv3_2", UNIT_CONV, 160, 160, 3}, {"conv3_3", UNIT_CONV, 160, 160, 3},
    {"side_op3", UNIT_CONV, 160, 16, 1}, {"up3", UNIT_UP, 1, 1, 5},
    {"conv4_1", UNIT_DIL, 160, 300, 3},  {"conv4_2", UNIT_DIL, 300, 300, 3},  {"conv4_3", UNIT_DIL, 300, 300, 3},
    {"side_op4", UNIT_DIL, 300, 16, 1},  {"up4", UNIT_UP, 1, 1, 5},
    {"merge_conv", UNIT_CONV, 64, 100, 3}, {"merge_conv2", UNIT_CONV, 100, 100, 3}, {"merge_conv3", UNIT_CONV, 100, 1, 1},
};
constexpr int kNumUnits = sizeof(kUnits) / sizeof(kUnits[0]);
enum UnitId { U_CONV1_1 = 0, U_CONV1_2, U_CONV1_3, U_SIDE1, U_CONV2_1, U_CONV2_2, U_CONV2_3, U_SIDE2, U_UP2,
              U_CONV3_1, U_CONV3_2, U_CONV3_3, U_SIDE3, U_UP3, U_CONV4_1, U_CONV4_2, U_CONV4_3, U_SIDE4, U_UP4,
              U_MERGE1, U_MERGE2, U_MERGE3 };

struct ConvUnit {
    int id = -1;              // index into kUnits
    int kind, Cin, Cout, K, dil, act;
    int Cin_pad;              // Cin rounded up to CV_CI
    float* w_fp32 = nullptr;  // [ceil(Cout/16)][Cin_pad][K^3][16], zero padded
    float* scale = nullptr;   // gamma * inv_std                       (lasagne BatchNormLayer, deterministic)
    float* shift = nullptr;   // beta - mean * gamma * inv_std
    float* up_W = nullptr;    // UNIT_UP: the (k,k,k) fixed kernel taken from the parameter list
    // host copies kept for the tensor-core weight preparation
    std::vector<float> h_w;   // canonical (Cout, Cin, K^3)
    std::vector<float> h_scale, h_shift;
};

struct Net {
    ConvUnit units[kNumUnits];
    float* mean6 = nullptr;   // params.py:129 __MEAN_CVC_RGBRGB
    float *fc1_W = nullptr, *fc1_scale = nullptr, *fc1_shift = nullptr, *lin_W = nullptr, *lin_b = nullptr;
    void* tc = nullptr;       // tensor-core side tables (conv_tc.cu), created lazily at sn_net_create
};

// Where the network input comes from when the caller does not need the fp32 CVC tensor X itself: the arguments of the CVC gather
// (utils/CVC.py:56-111).  The Winograd forward then colours conv1_1's operand straight from the images (conv_wg.cu:cvc_wino_kernel)
// and X (24 B per pair-voxel written + read back) is never materialised.
struct CvcSource {
    const uint8_t* images; const int64_t* img_offset; const int32_t* img_hw; int n_views;
    const double* P; const float* xyz; const float* resol; const int32_t* views;     // views: (N_cubes, N_vp, 2) = 2 slots per pair-cube
    int n_vp; const float* mean6;
};
bool tc_gathers_directly(const Net& net, int D, int mode);                           // conv_tc.cu: the forward of (D, mode) can take a CvcSource

// per-launch CUDA-event timing of the conv units (sn_profile_enable / sn_profile_collect; bench.py roofline)
void prof_begin(int unit, cudaStream_t st);
void prof_end(int unit, cudaStream_t st);

int conv_fp32_launch(const ConvUnit& u, const float* in, int n, int S, float* out, int C_total, int c_off, cudaStream_t st);
int maxpool2_launch(const float* in, int n, int C, int S, float* out, cudaStream_t st);
int upsample_launch(const float* in, const float* W, int k, int f, int n, int C, int S, float* out, int C_total, int c_off, cudaStream_t st);
int fuse_launch(const float* p, const float* w, int n_cubes, int n_vp, int64_t vol, float* out, cudaStream_t st);
int relimp_launch(const Net& net, const float* features, int64_t n_rows, int n_per_group, float* logit_tmp, float* out, cudaStream_t st);

}  // namespace sn

struct sn_net { sn::Net net; };

namespace sn {
// net.cu: sn_net_forward's body; the input is X_dev or, for the forwards tc_gathers_directly() names, the gather's arguments
int net_forward(const sn_net* h, const float* X_dev, const CvcSource* src, int n_cubes, int n_vp, int D, const float* w_dev, float* fused_out_dev,
                float* unfused_out_dev, void* workspace_dev, int64_t workspace_bytes, int mode, void* stream);
}
