// fp64 projection arithmetic shared by the CVC gather, perspectiveProj and ray pooling.
// Follows utils/CVC.py:13-20,36-39 and utils/camera.py:172-182 operation by operation.
#pragma once
#include "common.cuh"

namespace sn {

// voxel corner coordinate:  i * resol + min   with int64*f32 -> f64 promotion (CVC.py:16-18,
// rayPooling.py:223).  Explicit round-to-nearest mul/add: numpy does NOT contract them into an fma.
__device__ __forceinline__ double voxel_coord(int i, float resol, float mn) {
    return __dadd_rn(__dmul_rn((double)i, (double)resol), (double)mn);
}

// one row of P(3x4) . (x,y,z,1): the K=4 dgemm inner product, accumulated in k order with fused
// multiply-adds from zero, which is what the BLAS micro-kernel behind np.dot / np.matmul does
// (CVC.py:37, camera.py:174).  See DESIGN.md "index exactness".
__device__ __forceinline__ double proj_row(const double* __restrict__ p, double x, double y, double z) {
    double acc = __dmul_rn(p[0], x);
    acc = __fma_rn(p[1], y, acc);
    acc = __fma_rn(p[2], z, acc);
    acc = __fma_rn(p[3], 1.0, acc);
    return acc;
}

struct Proj { double u, t, q; };   // u -> image column (w), t -> image row (h), q -> depth (3rd homogeneous coordinate)

__device__ __forceinline__ Proj project(const double* __restrict__ P, double x, double y, double z) {
    Proj r;
    r.u = proj_row(P, x, y, z);
    r.t = proj_row(P + 4, x, y, z);
    r.q = proj_row(P + 8, x, y, z);
    return r;
}

// `.round().astype(np.int32)` (CVC.py:39): rint = half-to-even; values that do not fit (or NaN)
// become INT32_MIN, which is what the x86 conversion numpy uses produces, and are out of scope.
__device__ __forceinline__ int32_t round_to_i32(double v) {
    double r = rint(v);
    return (r >= -2147483648.0 && r <= 2147483647.0) ? (int32_t)r : (int32_t)0x80000000;
}

// `.round().astype(np.int64)` (camera.py:179); same convention for unrepresentable values.
__device__ __forceinline__ long long round_to_i64(double v) {
    double r = rint(v);
    return (r >= -9223372036854775808.0 && r < 9223372036854775808.0) ? (long long)r : (long long)0x8000000000000000ull;
}

}  // namespace sn
