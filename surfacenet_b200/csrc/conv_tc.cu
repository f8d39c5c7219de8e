// Tensor-core (tcgen05) convolution path -- placeholder wiring until the kernels land.
#include "net.cuh"

namespace sn {

int tc_prepare(Net& net) { (void)net; return SN_OK; }
void tc_destroy(Net& net) { (void)net; }
int64_t tc_workspace_bytes(const Net& net, int n_pc, int D, int mode) {
    (void)net; (void)n_pc; (void)D; (void)mode;
    set_error("tensor-core modes are not available in this build");
    return -1;
}
int tc_forward(const Net& net, const float* X, int n_pc, int D, float* prob_out, void* ws, int64_t ws_bytes, int mode, cudaStream_t st) {
    (void)net; (void)X; (void)n_pc; (void)D; (void)prob_out; (void)ws; (void)ws_bytes; (void)mode; (void)st;
    set_error("tensor-core modes are not available in this build");
    return SN_ERR_INVALID;
}

}  // namespace sn
