// Shared helpers for the surfacenet_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/surfacenet_b200.h"

namespace sn {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
extern std::atomic<int64_t> g_conv_path[3];     // launches of fp32 / direct tensor-core / Winograd tensor-core conv units

#define SN_CHECK_ARG(cond, ...)                         \
    do {                                                \
        if (!(cond)) {                                  \
            sn::set_error(__VA_ARGS__);                 \
            return SN_ERR_INVALID;                      \
        }                                               \
    } while (0)

#define SN_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            sn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return SN_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

// call after every kernel launch: counts it and surfaces launch-configuration errors
#define SN_LAUNCHED()                                   \
    do {                                                \
        sn::g_launches.fetch_add(1, std::memory_order_relaxed); \
        SN_CUDA(cudaGetLastError());                    \
    } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t align_up(int64_t a, int64_t b) { return cdiv(a, b) * b; }

// bump allocator over a caller-provided workspace
struct Arena {
    char* base; int64_t size; int64_t off;
    Arena(void* p, int64_t n) : base((char*)p), size(n), off(0) {}
    template <typename T> T* take(int64_t count) {
        int64_t bytes = align_up(count * (int64_t)sizeof(T), 256);
        T* r = (T*)(base + off);
        off += bytes;
        return r;
    }
    bool ok() const { return off <= size; }
};

int raypool_enqueue(const void* pred_dev, int pred_is_f16, int has_thresh, float thresh, const int32_t* viewpairs_dev,
                    const double* P_dev, int n_views, const float* xyz_dev, const float* resol_dev, int n_cubes, int n_vp,
                    int D, uint8_t* votes_out_dev, void* workspace_dev, int64_t workspace_bytes, void* stream,
                    int32_t** flags_dev_out);
int raypool_check(const int32_t* flags_dev, void* stream);

}  // namespace sn
