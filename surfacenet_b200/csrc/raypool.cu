// K6 -- ray-pool votes (utils/rayPooling.py:143-260 as called from utils/sparseCubes.py:57-62).
//
// Reference semantics per cube and per UNIQUE selected view (rayPooling.py:237-256):
//   S       = voxels with pred > thresh, ascending flat index n
//   cell    = (pixel (w,h), depth bin dint): the dense (pixel, depth) table is filled by fancy
//             assignment, so a cell holds the LAST voxel written = the largest n mapping to it (:251)
//   winner  = per pixel, argmax of pred over the cells; np.argmax returns the first maximum =
//             the smallest dint among equal preds (:253); empty cells hold 0 and selected preds are > 0
//   votes   = sum over the 2*N_vp view slots of "voxel is the winner of its pixel in that view" (:258)
//
// Device algorithm (exact, order independent): per (cube, first-occurrence view slot) two open
// addressing hash tables in global memory keyed by the voxel's own projection (the key of a
// claimed slot is recomputed from the claimant's index, so keys are full int64/int32 triples):
//   pass A  cell table:  claim (w,h,dint) cell, atomicMax of n                    -> cell representative
//   pass B  pixel table: representatives only, atomicMax of (pred bits, ~dint)    -> pixel winner key
//   pass C  representatives whose key equals the pixel's maximum add the view's multiplicity.
#include "geometry.cuh"
#include <algorithm>

namespace sn {

constexpr int RP_THREADS = 256;
constexpr int32_t RP_EMPTY = -1;

struct RpKey { long long w, h; int32_t d; };

// tables are ALLOCATED for the worst case (every voxel selected, rp_capacity) but only the first rp_dyn_cap(selected count) slots of
// each are cleared and probed: a power of two >= 2 x the cube's selected voxels (load factor <= 1/2 as before).  Capacity only changes
// where a key lands, never the result (the three passes are order independent and exact).
__device__ __forceinline__ uint32_t rp_dyn_cap(int cnt, uint32_t cap) {
    uint32_t c = 256;
    while (c < 2u * (uint32_t)cnt) c <<= 1;
    return c < cap ? c : cap;
}

struct RpCube {
    double P[12];
    float x0, y0, z0, rs;
    int D;
    __device__ __forceinline__ RpKey key(int n) const {
        const int k = n % D, j = (n / D) % D, i = n / (D * D);
        const Proj pr = project(P, voxel_coord(i, rs, x0), voxel_coord(j, rs, y0), voxel_coord(k, rs, z0));  // rayPooling.py:223, camera.py:173-174
        RpKey r;
        r.w = round_to_i64(__ddiv_rn(pr.u, pr.q));                      // camera.py:177-180
        r.h = round_to_i64(__ddiv_rn(pr.t, pr.q));
        r.d = round_to_i32(__ddiv_rn(pr.q, (double)rs));                // rayPooling.py:233
        return r;
    }
};

__device__ __forceinline__ uint32_t rp_hash3(const RpKey& k) {
    uint64_t x = (uint64_t)k.w * 0x9E3779B97F4A7C15ull ^ (uint64_t)k.h * 0xC2B2AE3D27D4EB4Full ^ (uint64_t)(uint32_t)k.d * 0x165667B19E3779F9ull;
    x ^= x >> 31; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 29;
    return (uint32_t)x;
}
__device__ __forceinline__ uint32_t rp_hash2(const RpKey& k) {
    uint64_t x = (uint64_t)k.w * 0xC2B2AE3D27D4EB4Full ^ (uint64_t)k.h * 0x9E3779B97F4A7C15ull;
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    return (uint32_t)x;
}

// find-or-claim the slot whose claimant has the same key (cell: w,h,d; pixel: w,h)
template <bool CELL, bool INSERT>
__device__ __forceinline__ uint32_t rp_probe(int32_t* claim, uint32_t mask, const RpCube& cube, int n, const RpKey& kn) {
    uint32_t s = (CELL ? rp_hash3(kn) : rp_hash2(kn)) & mask;
    while (true) {
        int32_t c = *((volatile int32_t*)&claim[s]);
        if (c == RP_EMPTY) {
            if (!INSERT) return 0xFFFFFFFFu;
            c = atomicCAS(&claim[s], RP_EMPTY, n);
            if (c == RP_EMPTY) return s;
        }
        if (c == n) return s;
        const RpKey kc = cube.key(c);
        if (kc.w == kn.w && kc.h == kn.h && (!CELL || kc.d == kn.d)) return s;
        s = (s + 1) & mask;
    }
}

template <typename T> __device__ __forceinline__ float rp_load(const T* p, int64_t i);
template <> __device__ __forceinline__ float rp_load<float>(const float* p, int64_t i) { return __ldg(p + i); }
template <> __device__ __forceinline__ float rp_load<__half>(const __half* p, int64_t i) { return __half2float(__ldg(p + i)); }

// selection `pred > thresh` (rayPooling.py:218-219) + unordered compaction of the selected indices
template <typename T>
__global__ void __launch_bounds__(RP_THREADS)
rp_select_kernel(const T* __restrict__ pred, int has_thresh, float thresh, int vol, int32_t* __restrict__ sel_count,
                 int32_t* __restrict__ sel_list, int32_t* __restrict__ flags) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * RP_THREADS + threadIdx.x;
    bool sel = false;
    if (n < vol) {
        const float p = rp_load<T>(pred, (int64_t)b * vol + n);
        sel = has_thresh ? (p > thresh) : true;
        if (sel && !(p > 0.f)) { atomicOr(&flags[0], 1); sel = false; }
    }
    const unsigned m = __ballot_sync(0xffffffffu, sel);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&sel_count[b], __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (sel) sel_list[(int64_t)b * vol + base + __popc(m & ((1u << lane) - 1))] = n;
}

__device__ __forceinline__ bool rp_setup(int b, int t, int T, int n_views, const int32_t* __restrict__ viewpairs,
                                         const double* __restrict__ P, const float* __restrict__ xyz,
                                         const float* __restrict__ resol, int D, RpCube& cube, int& mult) {
    const int32_t* vp = viewpairs + (int64_t)b * T;
    const int view = vp[t];
    if (view < 0 || view >= n_views) return false;
    for (int u = 0; u < t; ++u) if (vp[u] == view) return false;       // only the first occurrence works (np.unique, :210)
    mult = 0;
    for (int u = 0; u < T; ++u) mult += (vp[u] == view);               // a view used twice counts twice (:258)
#pragma unroll
    for (int i = 0; i < 12; ++i) cube.P[i] = __ldg(P + (int64_t)view * 12 + i);
    cube.x0 = xyz[3 * b]; cube.y0 = xyz[3 * b + 1]; cube.z0 = xyz[3 * b + 2]; cube.rs = resol[b]; cube.D = D;
    return true;
}

__device__ __forceinline__ unsigned long long rp_rank(float p, int32_t d) {
    // larger is better: pred first (positive floats order like their bit patterns), then smaller dint
    return ((unsigned long long)__float_as_uint(p) << 32) | (unsigned long long)(0xFFFFFFFFu - ((uint32_t)d ^ 0x80000000u));
}

// one selected voxel n of (cube b, view slot bt) in pass PASS
template <typename T, int PASS>
__device__ __forceinline__ void rp_item(const T* __restrict__ pred, const RpCube& cube, int b, int64_t bt, int n, int mult, int vol, uint32_t cap, uint32_t mask,
                                        int32_t* cell_claim, int32_t* cell_best, int32_t* pix_claim, unsigned long long* pix_val, uint8_t* votes) {
    int32_t* cc = cell_claim + bt * cap;
    int32_t* cb = cell_best + bt * cap;
    const RpKey kn = cube.key(n);
    if (PASS == 0) {
        const uint32_t s = rp_probe<true, true>(cc, mask, cube, n, kn);
        atomicMax(&cb[s], n);                                           // last write wins (:251)
        return;
    }
    const uint32_t s = rp_probe<true, false>(cc, mask, cube, n, kn);
    if (cb[s] != n) return;                                             // not the cell's representative
    int32_t* pc = pix_claim + bt * cap;
    unsigned long long* pv = pix_val + bt * cap;
    const unsigned long long rank = rp_rank(rp_load<T>(pred, (int64_t)b * vol + n), kn.d);
    if (PASS == 1) {
        const uint32_t q = rp_probe<false, true>(pc, mask, cube, n, kn);
        atomicMax(&pv[q], rank);                                        // argmax over depth (:253)
    } else {
        const uint32_t q = rp_probe<false, false>(pc, mask, cube, n, kn);
        if (pv[q] == rank) {                                            // (:255-256) + multiplicity (:258)
            const int64_t o = (int64_t)b * vol + n;
            atomicAdd(reinterpret_cast<unsigned int*>(votes + (o & ~(int64_t)3)), (unsigned)mult << (8 * (int)(o & 3)));
        }
    }
}

// grid = (blocks per slot, cube x view slot): the blocks of a slot stride over the cube's SELECTED voxels (their number is only known on
// the device; a grid sized for the whole volume would be mostly blocks that exit at once)
template <typename T, int PASS>
__global__ void __launch_bounds__(RP_THREADS)
rp_pass_kernel(const T* __restrict__ pred, const int32_t* __restrict__ viewpairs, const double* __restrict__ P, int n_views,
               const float* __restrict__ xyz, const float* __restrict__ resol, int n_vp, int D, int vol, uint32_t cap,
               const int32_t* __restrict__ sel_count, const int32_t* __restrict__ sel_list,
               int32_t* cell_claim, int32_t* cell_best, int32_t* pix_claim,
               unsigned long long* pix_val, uint8_t* votes) {
    const int NT = 2 * n_vp;
    const int bt = blockIdx.y;
    const int b = bt / NT, t = bt - b * NT;
    const int cnt = sel_count[b];
    if ((int)(blockIdx.x * RP_THREADS) >= cnt) return;
    RpCube cube; int mult;
    if (!rp_setup(b, t, NT, n_views, viewpairs, P, xyz, resol, D, cube, mult)) return;
    const uint32_t mask = rp_dyn_cap(cnt, cap) - 1;
    for (int e = blockIdx.x * RP_THREADS + threadIdx.x; e < cnt; e += gridDim.x * RP_THREADS)
        rp_item<T, PASS>(pred, cube, b, bt, sel_list[(int64_t)b * vol + e], mult, vol, cap, mask, cell_claim, cell_best, pix_claim, pix_val, votes);
}

__global__ void cast_f32_f16_kernel(const float* __restrict__ in, int64_t n, __half* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = __float2half_rn(in[i]);       // sparseCubes.py:115 astype(np.float16)
}

// claims = EMPTY, cell representatives = -1, pixel maxima = 0 over the slots the passes will use
__global__ void __launch_bounds__(RP_THREADS)
rp_clear_kernel(const int32_t* __restrict__ sel_count, int n_slots, uint32_t cap, int32_t* __restrict__ cell_claim,
                int32_t* __restrict__ cell_best, int32_t* __restrict__ pix_claim, unsigned long long* __restrict__ pix_val) {
    const int bt = blockIdx.y;
    const int cnt = sel_count[bt / n_slots];
    if (cnt == 0) return;
    const uint32_t capb = rp_dyn_cap(cnt, cap);
    const int64_t base = (int64_t)bt * cap;
    // 4 slots per thread and step: 16-byte stores (cap and capb are multiples of 256, the tables 256-byte aligned)
    const int4 m1 = make_int4(-1, -1, -1, -1);
    const ulonglong2 z = make_ulonglong2(0ull, 0ull);
    for (uint32_t i = (blockIdx.x * RP_THREADS + threadIdx.x) * 4; i < capb; i += gridDim.x * RP_THREADS * 4) {
        *reinterpret_cast<int4*>(cell_claim + base + i) = m1;
        *reinterpret_cast<int4*>(cell_best + base + i) = m1;
        *reinterpret_cast<int4*>(pix_claim + base + i) = m1;
        *reinterpret_cast<ulonglong2*>(pix_val + base + i) = z;
        *reinterpret_cast<ulonglong2*>(pix_val + base + i + 2) = z;
    }
}

uint32_t rp_capacity(int D) {
    uint64_t need = 2ull * D * D * D, c = 1024;
    while (c < need) c <<= 1;
    return (uint32_t)c;
}

}  // namespace sn

using namespace sn;

extern "C" int64_t sn_raypool_workspace_bytes(int n_cubes, int n_vp, int D) {
    if (n_cubes < 0 || n_vp < 1 || D < 1) return -1;
    const int64_t vol = (int64_t)D * D * D, bt = (int64_t)n_cubes * 2 * n_vp, cap = rp_capacity(D);
    int64_t bytes = 256;                                    // flags
    bytes += align_up(n_cubes * 4, 256);                    // sel_count
    bytes += align_up(n_cubes * vol * 4, 256);              // sel_list
    bytes += 3 * align_up(bt * cap * 4, 256);               // cell_claim, cell_best, pix_claim
    bytes += align_up(bt * cap * 8, 256);                   // pix_val
    bytes += align_up(n_cubes * vol + 4, 256);              // word-aligned votes accumulator
    return bytes;
}

// enqueue only; *flags_dev_out receives the device address of the domain-error flag word
int sn::raypool_enqueue(const void* pred_dev, int pred_is_f16, int has_thresh, float thresh, const int32_t* viewpairs_dev,
                        const double* P_dev, int n_views, const float* xyz_dev, const float* resol_dev, int n_cubes,
                        int n_vp, int D, uint8_t* votes_out_dev, void* workspace_dev, int64_t workspace_bytes, void* stream,
                        int32_t** flags_dev_out) {
    SN_CHECK_ARG(pred_dev && viewpairs_dev && P_dev && xyz_dev && resol_dev && votes_out_dev, "sn_raypool_votes: NULL argument");
    SN_CHECK_ARG(n_cubes >= 0 && n_vp >= 1 && n_vp <= 127 && D >= 1 && D <= 256 && n_views >= 1, "sn_raypool_votes: bad sizes (n_cubes=%d n_vp=%d D=%d)", n_cubes, n_vp, D);
    if (n_cubes == 0) return SN_OK;
    SN_CHECK_ARG(((uintptr_t)workspace_dev & 15) == 0, "sn_raypool_votes: the workspace must be 16-byte aligned");
    const int64_t need = sn_raypool_workspace_bytes(n_cubes, n_vp, D);
    if (!workspace_dev || workspace_bytes < need) { set_error("sn_raypool_votes: workspace %lld B < %lld B", (long long)workspace_bytes, (long long)need); return SN_ERR_NOMEM; }
    const int vol = D * D * D;
    const int64_t bt = (int64_t)n_cubes * 2 * n_vp;
    SN_CHECK_ARG(bt <= 65535, "sn_raypool_votes: n_cubes*2*n_vp = %lld exceeds 65535 per call", (long long)bt);
    const uint32_t cap = rp_capacity(D);
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(workspace_dev, workspace_bytes);
    int32_t* flags = ar.take<int32_t>(64);
    int32_t* sel_count = ar.take<int32_t>(n_cubes);
    int32_t* sel_list = ar.take<int32_t>((int64_t)n_cubes * vol);
    int32_t* cell_claim = ar.take<int32_t>(bt * cap);
    int32_t* cell_best = ar.take<int32_t>(bt * cap);
    int32_t* pix_claim = ar.take<int32_t>(bt * cap);
    unsigned long long* pix_val = ar.take<unsigned long long>(bt * cap);
    uint8_t* votes_acc = ar.take<uint8_t>((int64_t)n_cubes * vol + 4);
    if (flags_dev_out) *flags_dev_out = flags;
    // flags + sel_count are contiguous at the head of the arena
    SN_CUDA(cudaMemsetAsync(flags, 0, (char*)sel_list - (char*)flags, st));
    SN_CUDA(cudaMemsetAsync(votes_acc, 0, align_up((int64_t)n_cubes * vol, 4), st));      // rayPooling.py:235
    // >= 16 blocks per SM over all slots when the volume allows it; each slot's blocks loop over its selected voxels
    const unsigned per_slot = (unsigned)std::min<int64_t>(cdiv(vol, RP_THREADS), std::max<int64_t>(8, cdiv(148 * 16, bt)));
    dim3 gsel((unsigned)cdiv(vol, RP_THREADS), (unsigned)n_cubes), gpass(per_slot, (unsigned)bt);
    dim3 gclr((unsigned)std::min<int64_t>(cdiv((int64_t)cap, RP_THREADS * 4), 32), (unsigned)bt);
#define SN_RP_CLEAR rp_clear_kernel<<<gclr, RP_THREADS, 0, st>>>(sel_count, 2 * n_vp, cap, cell_claim, cell_best, pix_claim, pix_val); SN_LAUNCHED()
#define SN_RP_ARGS viewpairs_dev, P_dev, n_views, xyz_dev, resol_dev, n_vp, D, vol, cap, sel_count, sel_list, cell_claim, cell_best, pix_claim, pix_val, votes_acc
    if (pred_is_f16) {
        const __half* p = (const __half*)pred_dev;
        rp_select_kernel<__half><<<gsel, RP_THREADS, 0, st>>>(p, has_thresh, thresh, vol, sel_count, sel_list, flags); SN_LAUNCHED();
        SN_RP_CLEAR;
        rp_pass_kernel<__half, 0><<<gpass, RP_THREADS, 0, st>>>(p, SN_RP_ARGS); SN_LAUNCHED();
        rp_pass_kernel<__half, 1><<<gpass, RP_THREADS, 0, st>>>(p, SN_RP_ARGS); SN_LAUNCHED();
        rp_pass_kernel<__half, 2><<<gpass, RP_THREADS, 0, st>>>(p, SN_RP_ARGS); SN_LAUNCHED();
    } else {
        const float* p = (const float*)pred_dev;
        rp_select_kernel<float><<<gsel, RP_THREADS, 0, st>>>(p, has_thresh, thresh, vol, sel_count, sel_list, flags); SN_LAUNCHED();
        SN_RP_CLEAR;
        rp_pass_kernel<float, 0><<<gpass, RP_THREADS, 0, st>>>(p, SN_RP_ARGS); SN_LAUNCHED();
        rp_pass_kernel<float, 1><<<gpass, RP_THREADS, 0, st>>>(p, SN_RP_ARGS); SN_LAUNCHED();
        rp_pass_kernel<float, 2><<<gpass, RP_THREADS, 0, st>>>(p, SN_RP_ARGS); SN_LAUNCHED();
    }
#undef SN_RP_ARGS
#undef SN_RP_CLEAR
    SN_CUDA(cudaMemcpyAsync(votes_out_dev, votes_acc, (int64_t)n_cubes * vol, cudaMemcpyDeviceToDevice, st));
    return SN_OK;
}

// synchronises the stream and turns the device flag into an error code
int sn::raypool_check(const int32_t* flags_dev, void* stream) {
    int32_t hflag = 0;
    SN_CUDA(cudaMemcpyAsync(&hflag, flags_dev, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    SN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (hflag & 1) {
        set_error("ray pooling: a selected prediction is <= 0 (prediction_thresh must be >= 0, or all predictions > 0 when it is None)");
        return SN_ERR_DOMAIN;
    }
    return SN_OK;
}

extern "C" int sn_raypool_votes(const void* pred_dev, int pred_is_f16, int has_thresh, float thresh, const int32_t* viewpairs_dev,
                                const double* P_dev, int n_views, const float* xyz_dev, const float* resol_dev, int n_cubes,
                                int n_vp, int D, uint8_t* votes_out_dev, void* workspace_dev, int64_t workspace_bytes, void* stream) {
    int32_t* flags = nullptr;
    int rc = raypool_enqueue(pred_dev, pred_is_f16, has_thresh, thresh, viewpairs_dev, P_dev, n_views, xyz_dev, resol_dev,
                             n_cubes, n_vp, D, votes_out_dev, workspace_dev, workspace_bytes, stream, &flags);
    if (rc != SN_OK || n_cubes == 0) return rc;
    return raypool_check(flags, stream);
}

extern "C" int sn_cast_f32_to_f16(const float* in_dev, int64_t n, void* out_f16_dev, void* stream) {
    SN_CHECK_ARG(in_dev && out_f16_dev && n >= 0, "sn_cast_f32_to_f16: bad arguments");
    if (n == 0) return SN_OK;
    const int blocks = (int)std::min<int64_t>(cdiv(n, 256), 148 * 16);
    cast_f32_f16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in_dev, n, (__half*)out_f16_dev);
    SN_LAUNCHED();
    return SN_OK;
}
