// "Next" row N4 of the scope table: the consumers of the sparse voxel lists.
//   utils/sparseCubes.py:205-243  filter_voxels          -> sn_sparse_filter_voxels
//   utils/denoising.py:8-184      __cluster_inCube__ / __mark_overlappingLabels__ / denoise_crossCubes -> sn_sparse_denoise
//   utils/adapthresh.py:91-178    adapthresh refinement loop (cost = XOR - beta*AND of half-cube occupancies) -> sn_sparse_adapthresh
// All integer / byte / float16-compare work, bit-exact against the reference; HBM / atomic bound.
//
// Data layout: the scene's sparse cubes are ONE flat voxel array (ijk u8[N,3], pred f16[N], votes u8[N], mask u8[N]) with
// cube n owning [cube_offset[n], cube_offset[n+1]).  Per cube the workspace holds a G^3-bit occupancy bitmap and a per-word
// exclusive popcount prefix, so that   rank(pos) = prefix[word] + popc(bits below)   maps an occupied position to its index in
// raster (C-order) position order -- the order scipy.ndimage.label scans in.  Connected components are a union-find over
// those ranks with min-rank roots, so root order == first-voxel raster order == scipy's label numbering (denoising.py:54).
// Neighbouring cubes are found through a hash map cube ijk -> index that reproduces the reference dict (non-empty cubes only,
// the later index wins for repeated keys: denoising.py:106-108, adapthresh.py:121-124).
#include "common.cuh"
#include <algorithm>

namespace sn {

constexpr int PP_THREADS = 256;

struct PostWs {
    int32_t* vox_cube;     // [N]   cube of each voxel
    uint32_t* bitmap;      // [C*W] occupancy bits of the masked voxels, position = (i*G + j)*G + k
    uint32_t* prefix;      // [C*W] exclusive popcount prefix per word
    int32_t* n_masked;     // [C]   distinct masked positions per cube
    int32_t* parent;       // [N]   union-find over local ranks (slot cube_offset[c] + rank)
    int32_t* root;         // [N]
    int32_t* labelnum;     // [N]   scipy label number of a root rank
    uint8_t* ovl;          // [N]   root rank -> its cluster overlaps a neighbouring cube
    unsigned long long* hkeys;  // [H]
    int32_t* hvals;        // [H]
    int32_t* canon;        // [C]   1 = this cube is the dict's entry for its ijk (non-empty, last of its key)
    int32_t* nb27;         // [C*27] dict lookup of cube ijk + shift t (t = (si+1)*9 + (sj+1)*3 + (sk+1)), -1 = none
    int32_t* counts;       // [C*48] adapthresh: n_cur[6][3], n_and[6][3], nb[6], n_half0[6]
    int32_t* flags;        // [4]   error flags
    int64_t W, H;
};

static int64_t post_hash_size(int n_cubes) {
    int64_t h = 64;
    while (h < 2 * (int64_t)n_cubes) h <<= 1;
    return h;
}

static int64_t post_layout(void* ws, int64_t ws_bytes, int n_cubes, int64_t n_vox, int G, PostWs* out) {
    Arena a(ws, ws_bytes);
    PostWs w;
    w.W = ((int64_t)G * G * G + 31) / 32;
    w.H = post_hash_size(n_cubes);
    const int64_t C = std::max(n_cubes, 1), N = std::max<int64_t>(n_vox, 1);
    w.vox_cube = a.take<int32_t>(N);
    w.bitmap = a.take<uint32_t>(C * w.W);
    w.prefix = a.take<uint32_t>(C * w.W);
    w.n_masked = a.take<int32_t>(C);
    w.parent = a.take<int32_t>(N);
    w.root = a.take<int32_t>(N);
    w.labelnum = a.take<int32_t>(N);
    w.ovl = a.take<uint8_t>(N);
    w.hkeys = a.take<unsigned long long>(w.H);
    w.hvals = a.take<int32_t>(w.H);
    w.canon = a.take<int32_t>(C);
    w.nb27 = a.take<int32_t>(C * 27);
    w.counts = a.take<int32_t>(C * 48);
    w.flags = a.take<int32_t>(4);
    if (out) *out = w;
    return a.off;
}

// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pp_key(int i, int j, int k) {
    return (((unsigned long long)i << 42) | ((unsigned long long)j << 21) | (unsigned long long)k) + 1ull;
}
__device__ __forceinline__ uint32_t pp_hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return (uint32_t)k;
}
__device__ __forceinline__ int pp_lookup(const unsigned long long* keys, const int32_t* vals, int64_t H, int i, int j, int k) {
    if (i < 0 || j < 0 || k < 0 || i >= (1 << 21) || j >= (1 << 21) || k >= (1 << 21)) return -1;
    const unsigned long long key = pp_key(i, j, k);
    uint32_t s = pp_hash(key) & (uint32_t)(H - 1);
    while (true) {
        const unsigned long long kk = keys[s];
        if (kk == key) return vals[s];
        if (kk == 0ull) return -1;
        s = (s + 1) & (uint32_t)(H - 1);
    }
}

__global__ void pp_vox_cube_kernel(const int64_t* __restrict__ off, int n_cubes, int64_t n_vox, int32_t* __restrict__ vox_cube) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vox) return;
    int lo = 0, hi = n_cubes;                    // largest c with off[c] <= v
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= v) lo = mid; else hi = mid; }
    vox_cube[v] = lo;
}

// mask bits of every cube; positions outside the G^3 grid raise flag 0
__global__ void pp_bitmap_kernel(const uint8_t* __restrict__ ijk, const uint8_t* __restrict__ mask, const int32_t* __restrict__ vox_cube,
                                 int64_t n_vox, int G, int64_t W, uint32_t* __restrict__ bitmap, int32_t* __restrict__ flags) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vox || !mask[v]) return;
    const int i = ijk[3 * v], j = ijk[3 * v + 1], k = ijk[3 * v + 2];
    if (i >= G || j >= G || k >= G) { flags[0] = 1; return; }
    const int pos = (i * G + j) * G + k;
    atomicOr(&bitmap[(int64_t)vox_cube[v] * W + (pos >> 5)], 1u << (pos & 31));
}

// per cube: exclusive popcount prefix over the bitmap words, number of occupied positions
__global__ void __launch_bounds__(PP_THREADS)
pp_prefix_kernel(const uint32_t* __restrict__ bitmap, int64_t W, uint32_t* __restrict__ prefix, int32_t* __restrict__ n_masked) {
    __shared__ int warp_sum[PP_THREADS / 32];
    __shared__ int carry_s;
    const int c = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t* bm = bitmap + (int64_t)c * W;
    uint32_t* pf = prefix + (int64_t)c * W;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < W; base += PP_THREADS) {
        const int64_t w = base + threadIdx.x;
        const int x = w < W ? __popc(bm[w]) : 0;
        int incl = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
        if (lane == 31) warp_sum[wid] = incl;
        __syncthreads();
        int woff = 0;
        for (int q = 0; q < wid; ++q) woff += warp_sum[q];
        const int carry = carry_s;
        if (w < W) pf[w] = (uint32_t)(carry + woff + incl - x);
        __syncthreads();
        if (threadIdx.x == PP_THREADS - 1) carry_s = carry + woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_masked[c] = carry_s;
}

__device__ __forceinline__ int pp_rank(const uint32_t* bm, const uint32_t* pf, int pos) {
    const int w = pos >> 5;
    return (int)pf[w] + __popc(bm[w] & ((1u << (pos & 31)) - 1u));
}
__device__ __forceinline__ bool pp_bit(const uint32_t* bm, int pos) { return (bm[pos >> 5] >> (pos & 31)) & 1u; }

// dict build: non-empty cubes, the later index wins                                   denoising.py:106-108
__global__ void pp_hash_insert_kernel(const int32_t* __restrict__ cube_ijk, const int32_t* __restrict__ n_masked, int n_cubes,
                                      unsigned long long* keys, int32_t* vals, int64_t H, int32_t* flags) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_cubes || n_masked[n] <= 0) return;
    const int i = cube_ijk[3 * n], j = cube_ijk[3 * n + 1], k = cube_ijk[3 * n + 2];
    if (i < 0 || j < 0 || k < 0 || i >= (1 << 21) || j >= (1 << 21) || k >= (1 << 21)) { flags[1] = 1; return; }
    const unsigned long long key = pp_key(i, j, k);
    uint32_t s = pp_hash(key) & (uint32_t)(H - 1);
    while (true) {
        const unsigned long long old = atomicCAS(&keys[s], 0ull, key);
        if (old == 0ull || old == key) { atomicMax(&vals[s], n); return; }
        s = (s + 1) & (uint32_t)(H - 1);
    }
}
__global__ void pp_canon_kernel(const int32_t* __restrict__ cube_ijk, const int32_t* __restrict__ n_masked, int n_cubes,
                                const unsigned long long* __restrict__ keys, const int32_t* __restrict__ vals, int64_t H,
                                int32_t* __restrict__ canon, int32_t* __restrict__ nb27) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_cubes * 27) return;
    const int n = g / 27, t = g % 27;
    const int m = pp_lookup(keys, vals, H, cube_ijk[3 * n] + t / 9 - 1, cube_ijk[3 * n + 1] + (t / 3) % 3 - 1, cube_ijk[3 * n + 2] + t % 3 - 1);
    nb27[g] = m;
    if (t == 13) canon[n] = (n_masked[n] > 0 && m == n) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------
// connected components (scipy.ndimage.label with generate_binary_structure(3, neighbor_dist))      denoising.py:53-54
// Label-equivalence scheme on the raster ranks: (1) every voxel points at its first present neighbour among the 13 that
// precede it in raster order (a strictly smaller rank, so the forest is acyclic), (2) the chains are flattened, (3) every
// voxel merges with each preceding neighbour whose root differs (atomicMin hooking, the smaller root wins), (4) final roots.
__device__ __forceinline__ int pp_find(volatile int32_t* par, int x) {
    int p;
    while ((p = par[x]) != x) x = p;
    return x;
}
// find with path halving: every second node on the way is re-pointed at its grandparent.  Stores race with atomicMin hooks
// only in the harmless direction (any stored value is an ancestor of the node, ancestors have smaller ranks).
__device__ __forceinline__ int pp_find_halve(volatile int32_t* par, int x) {
    while (true) {
        const int p = par[x];
        if (p == x) return x;
        const int gp = par[p];
        if (gp == p) return p;
        par[x] = gp;
        x = gp;
    }
}
// merge the trees of a and b (ra = current root of a, kept up to date for the caller)
__device__ __forceinline__ int pp_unite(int32_t* par, int ra, int b) {
    int rb = pp_find_halve(par, b);
    while (ra != rb) {
        if (ra > rb) { const int t = ra; ra = rb; rb = t; }
        const int old = atomicMin(&par[rb], ra);         // hook the larger root under the smaller
        if (old == rb) break;
        rb = pp_find_halve(par, old);
        ra = pp_find_halve(par, ra);
    }
    return ra < rb ? ra : rb;
}

template <bool REDUCE>
__global__ void pp_ccl_link_kernel(const uint8_t* __restrict__ ijk, const uint8_t* __restrict__ mask, const int32_t* __restrict__ vox_cube,
                                   const int64_t* __restrict__ off, int64_t n_vox, int G, int64_t W, int neighbor_dist,
                                   const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ prefix, int32_t* parent) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vox || !mask[v]) return;
    const int c = vox_cube[v];
    const int i = ijk[3 * v], j = ijk[3 * v + 1], k = ijk[3 * v + 2];
    if (i >= G || j >= G || k >= G) return;
    const uint32_t* bm = bitmap + (int64_t)c * W;
    const uint32_t* pf = prefix + (int64_t)c * W;
    int32_t* par = parent + off[c];
    const int pos = (i * G + j) * G + k;
    // Phase 1 (independent loads, issued together): which of the 13 neighbours that precede (i,j,k) in raster order exist.
    // 26-connectivity, reduce pass: when the previous voxel of the row (0,0,-1) is occupied, every preceding neighbour with
    // dk <= 0 is also a preceding neighbour of THAT voxel and gets merged from there; only the four (di,dj,+1) remain.
    int q[13];
#pragma unroll
    for (int t = 0; t < 13; ++t) {
        const int di = t / 9 - 1, dj = (t / 3) % 3 - 1, dk = t % 3 - 1;
        const int a = i + di, b = j + dj, cc = k + dk;
        const bool ok = ((di != 0) + (dj != 0) + (dk != 0) <= neighbor_dist) && a >= 0 && b >= 0 && cc >= 0 && a < G && b < G && cc < G;
        const int p = pos + (di * G + dj) * G + dk;
        q[t] = (ok && pp_bit(bm, p)) ? p : -1;
    }
    if (!REDUCE) {
        int first = -1;
#pragma unroll
        for (int t = 12; t >= 0; --t) if (q[t] >= 0) first = q[t];                 // smallest position = smallest rank
        const int r = pp_rank(bm, pf, pos);
        par[r] = first >= 0 ? pp_rank(bm, pf, first) : r;
        return;
    }
    if (neighbor_dist == 3 && q[12] >= 0) {
#pragma unroll
        for (int t = 0; t < 12; ++t) if (t % 3 != 2) q[t] = -1;
    }
    // Phase 2: ranks, Phase 3: current parents (plain loads: a stale parent is still an ancestor, good enough for the filter)
    int rq[13], pq[13];
#pragma unroll
    for (int t = 0; t < 13; ++t) rq[t] = q[t] >= 0 ? pp_rank(bm, pf, q[t]) : -1;
#pragma unroll
    for (int t = 0; t < 13; ++t) pq[t] = rq[t] >= 0 ? __ldcg(par + rq[t]) : -1;
    int ra = pp_find_halve(par, pp_rank(bm, pf, pos));
    // Phase 4: only neighbours that hang under another node than our root need the (serial, atomic) merge
#pragma unroll
    for (int t = 0; t < 13; ++t)
        if (rq[t] >= 0 && pq[t] != ra) ra = pp_unite(par, ra, rq[t]);
}

__global__ void pp_ccl_flatten_kernel(const int64_t* __restrict__ off, const int32_t* __restrict__ vox_cube, const int32_t* __restrict__ n_masked,
                                      int64_t n_vox, int32_t* parent) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vox) return;
    const int c = vox_cube[v];
    const int r = (int)(v - off[c]);
    if (r >= n_masked[c]) return;
    int32_t* par = parent + off[c];
    par[r] = pp_find(par, r);                            // concurrent writers only ever store ancestors: chains stay valid
}

__global__ void pp_ccl_root_kernel(const int64_t* __restrict__ off, const int32_t* __restrict__ vox_cube, const int32_t* __restrict__ n_masked,
                                   int64_t n_vox, int32_t* parent, int32_t* __restrict__ root) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vox) return;
    const int c = vox_cube[v];
    const int r = (int)(v - off[c]);
    root[v] = r < n_masked[c] ? pp_find(parent + off[c], r) : r;
}

// scipy numbering: label of a component = 1 + number of component roots with a smaller rank
__global__ void __launch_bounds__(PP_THREADS)
pp_label_rank_kernel(const int64_t* __restrict__ off, const int32_t* __restrict__ n_masked, const int32_t* __restrict__ root,
                     int32_t* __restrict__ labelnum, int32_t* __restrict__ n_labels) {
    __shared__ int warp_sum[PP_THREADS / 32];
    __shared__ int carry_s;
    const int c = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n = n_masked[c];
    const int32_t* rt = root + off[c];
    int32_t* ln = labelnum + off[c];
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += PP_THREADS) {
        const int r = base + threadIdx.x;
        const int x = (r < n && rt[r] == r) ? 1 : 0;
        int incl = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
        if (lane == 31) warp_sum[wid] = incl;
        __syncthreads();
        int woff = 0;
        for (int q = 0; q < wid; ++q) woff += warp_sum[q];
        const int carry = carry_s;
        if (r < n) ln[r] = carry + woff + incl;            // for a root: its 1-based number
        __syncthreads();
        if (threadIdx.x == PP_THREADS - 1) carry_s = carry + woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && n_labels) n_labels[c] = carry_s;
}

__global__ void pp_label_out_kernel(const uint8_t* __restrict__ ijk, const uint8_t* __restrict__ mask, const int32_t* __restrict__ vox_cube,
                                    const int64_t* __restrict__ off, int64_t n_vox, int G, int64_t W, const uint32_t* __restrict__ bitmap,
                                    const uint32_t* __restrict__ prefix, const int32_t* __restrict__ root, const int32_t* __restrict__ labelnum,
                                    uint32_t* __restrict__ labels_out) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vox) return;
    uint32_t lab = 0;
    const int i = ijk[3 * v], j = ijk[3 * v + 1], k = ijk[3 * v + 2];
    if (mask[v] && i < G && j < G && k < G) {
        const int c = vox_cube[v];
        const int r = pp_rank(bitmap + (int64_t)c * W, prefix + (int64_t)c * W, (i * G + j) * G + k);
        lab = (uint32_t)labelnum[off[c] + root[off[c] + r]];
    }
    labels_out[v] = lab;
}

// cross-cube overlap: a masked voxel p of cube n coincides with a masked voxel of neighbour m (cube ijk + s) when
// p == u + (D_cube/2)*s  (denoising.py:127-131); both clusters are then "overlapping" (the reference marks both sides, here each
// side marks itself because the relation is symmetric)
__global__ void pp_overlap_kernel(const uint8_t* __restrict__ ijk, const uint8_t* __restrict__ mask, const int32_t* __restrict__ vox_cube,
                                  const int64_t* __restrict__ off, const int32_t* __restrict__ cube_ijk, const int32_t* __restrict__ canon,
                                  int64_t n_vox, int G, int64_t W, int half, const uint32_t* __restrict__ bitmap,
                                  const uint32_t* __restrict__ prefix, const int32_t* __restrict__ root,
                                  const int32_t* __restrict__ nb27, uint8_t* __restrict__ ovl) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vox || !mask[v]) return;
    const int n = vox_cube[v];
    if (!canon[n]) return;
    const int i = ijk[3 * v], j = ijk[3 * v + 1], k = ijk[3 * v + 2];
    if (i >= G || j >= G || k >= G) return;
    bool hit = false;
#pragma unroll
    for (int t = 0; t < 27; ++t) {                       // unrolled: the shifts are compile-time constants
        if (t == 13 || hit) continue;
        const int si = t / 9 - 1, sj = (t / 3) % 3 - 1, sk = t % 3 - 1;
        const int a = i - half * si, b = j - half * sj, c = k - half * sk;
        if (a < 0 || b < 0 || c < 0 || a >= G || b >= G || c >= G) continue;
        const int m = nb27[n * 27 + t];
        if (m < 0) continue;
        hit = pp_bit(bitmap + (int64_t)m * W, (a * G + b) * G + c);
    }
    if (hit) {
        const int r = pp_rank(bitmap + (int64_t)n * W, prefix + (int64_t)n * W, (i * G + j) * G + k);
        ovl[off[n] + root[off[n] + r]] = 1;
    }
}

__global__ void pp_keep_kernel(const uint8_t* __restrict__ ijk, const uint8_t* __restrict__ mask, const int32_t* __restrict__ vox_cube,
                               const int64_t* __restrict__ off, int64_t n_vox, int G, int64_t W, const uint32_t* __restrict__ bitmap,
                               const uint32_t* __restrict__ prefix, const int32_t* __restrict__ root, const uint8_t* __restrict__ ovl,
                               uint8_t* __restrict__ keep) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vox) return;
    uint8_t kp = 0;
    const int i = ijk[3 * v], j = ijk[3 * v + 1], k = ijk[3 * v + 2];
    if (mask[v] && i < G && j < G && k < G) {
        const int c = vox_cube[v];
        const int r = pp_rank(bitmap + (int64_t)c * W, prefix + (int64_t)c * W, (i * G + j) * G + k);
        kp = ovl[off[c] + root[off[c] + r]];
    }
    keep[v] = kp;                                                               // np.in1d(labels, overlappingLabels)  denoising.py:181
}

// ---------------------------------------------------------------------------------------------------------
// filter_voxels                                                                           sparseCubes.py:205-243
__global__ void pp_filter_kernel(const __half* __restrict__ pred, const uint8_t* __restrict__ votes, const int32_t* __restrict__ vox_cube,
                                 int64_t n_vox, const double* __restrict__ thresh_per_cube, double thresh_scalar, int has_prob,
                                 int rp_thresh, int and_into, uint8_t* __restrict__ mask) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vox) return;
    bool m = and_into ? (mask[v] != 0) : true;
    if (has_prob) {
        const double t = thresh_per_cube ? thresh_per_cube[vox_cube[v]] : thresh_scalar;
        m = m && (__half2float(pred[v]) >= __half2float(__double2half(t)));      // float16 array >= python float: compared in float16
    }
    if (rp_thresh >= 0 && votes) m = m && ((int)votes[v] >= rp_thresh);
    mask[v] = m ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------
// adapthresh.  Shift index q: 0..2 = +x,+y,+z, 3..5 = -x,-y,-z (adapthresh.py:114).
__device__ __forceinline__ bool pp_in_half(int i, int j, int k, int q, int D, int Dmid) {         // access_partial_Occupancy_ijk(shift_q)
    if (i >= D || j >= D || k >= D) return false;
    const int c = (q % 3 == 0) ? i : ((q % 3 == 1) ? j : k);
    return q < 3 ? (c >= Dmid) : (c < Dmid);
}

// occupancy at the current threshold: bitmap + number of voxels in each half                       adapthresh.py:149-151
// grid (C, PP_SLICES): a cube's voxel list is dealt to PP_SLICES blocks (a scene has far fewer cubes than the GPU has warps)
constexpr int PP_SLICES = 8, PP_OCC_SLICES = 32;
__global__ void __launch_bounds__(PP_THREADS)
pp_ada_occ0_kernel(const uint8_t* __restrict__ ijk, const __half* __restrict__ pred, const uint8_t* __restrict__ mask,
                   const int64_t* __restrict__ off, const double* __restrict__ thresh, int G, int64_t W, int D, int Dmid,
                   uint32_t* __restrict__ bitmap, int32_t* __restrict__ counts, int32_t* __restrict__ flags) {
    __shared__ int nh[6];
    const int c = blockIdx.x;
    if (threadIdx.x < 6) nh[threadIdx.x] = 0;
    __syncthreads();
    const float thr = __half2float(__double2half(thresh[c]));
    uint32_t* bm = bitmap + (int64_t)c * W;
    int loc[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t v = off[c] + blockIdx.y * PP_THREADS + threadIdx.x; v < off[c + 1]; v += PP_THREADS * PP_OCC_SLICES) {
        if (!mask[v] || !(__half2float(pred[v]) >= thr)) continue;
        const int i = ijk[3 * v], j = ijk[3 * v + 1], k = ijk[3 * v + 2];
        if (i >= G || j >= G || k >= G) { flags[0] = 1; continue; }
        const int pos = (i * G + j) * G + k;
        atomicOr(&bm[pos >> 5], 1u << (pos & 31));
#pragma unroll
        for (int q = 0; q < 6; ++q) loc[q] += pp_in_half(i, j, k, q, D, Dmid) ? 1 : 0;
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        int x = loc[q];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(&nh[q], x);
    }
    __syncthreads();
    if (threadIdx.x < 6 && nh[threadIdx.x]) atomicAdd(&counts[c * 48 + 42 + threadIdx.x], nh[threadIdx.x]);
}

// per dict cube: for the 6 face neighbours x 3 threshold perturbations, |current half| and |current half AND neighbour half|
__global__ void __launch_bounds__(PP_THREADS)
pp_ada_count_kernel(const uint8_t* __restrict__ ijk, const __half* __restrict__ pred, const uint8_t* __restrict__ mask,
                    const int64_t* __restrict__ off, const int32_t* __restrict__ cube_ijk, const int32_t* __restrict__ canon,
                    const double* __restrict__ thresh, int G, int64_t W, int D, int Dmid, const uint32_t* __restrict__ bitmap,
                    const int32_t* __restrict__ nb27, int32_t* __restrict__ counts) {
    __shared__ int sc[36];
    __shared__ int nb[6];
    const int n = blockIdx.x;
    if (!canon[n]) return;
    if (threadIdx.x < 36) sc[threadIdx.x] = 0;
    if (threadIdx.x < 6) {
        const int q = threadIdx.x, sg = q < 3 ? 1 : -1, d = q % 3;                   // shift_q as index into the 27-neighbour table
        nb[q] = nb27[n * 27 + (1 + (d == 0 ? sg : 0)) * 9 + (1 + (d == 1 ? sg : 0)) * 3 + (1 + (d == 2 ? sg : 0))];
    }
    __syncthreads();
    const double t0 = thresh[n];
    const float thr[3] = {__half2float(__double2half(t0 + 0.1)), __half2float(__double2half(t0 + 0)),
                          __half2float(__double2half(t0 + -0.1))};                                        // adapthresh.py:115
    int ncur[6][3], nand[6][3];
#pragma unroll
    for (int q = 0; q < 6; ++q)
#pragma unroll
        for (int t = 0; t < 3; ++t) { ncur[q][t] = 0; nand[q][t] = 0; }
    for (int64_t v = off[n] + blockIdx.y * PP_THREADS + threadIdx.x; v < off[n + 1]; v += PP_THREADS * PP_SLICES) {
        if (!mask[v]) continue;
        const float p = __half2float(pred[v]);
        const int o0 = p >= thr[0], o1 = p >= thr[1], o2 = p >= thr[2];
        if (!(o0 | o1 | o2)) continue;
        const int i = ijk[3 * v], j = ijk[3 * v + 1], k = ijk[3 * v + 2];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            if (!pp_in_half(i, j, k, q, D, Dmid)) continue;
            ncur[q][0] += o0; ncur[q][1] += o1; ncur[q][2] += o2;
            const int m = nb[q];
            if (m < 0) continue;
            // the neighbour's voxel u with u - Dmid*[neighbour half is the upper one] == p - Dmid*[current half is the upper one]
            int a = i, b = j, c = k;
            const int delta = q < 3 ? -Dmid : Dmid;
            if (q % 3 == 0) a += delta; else if (q % 3 == 1) b += delta; else c += delta;
            const int u = (q % 3 == 0) ? a : ((q % 3 == 1) ? b : c);
            if (u < 0 || u >= G || u >= D) continue;
            if (q < 3 ? (u >= Dmid) : (u < Dmid)) continue;          // must lie in the neighbour's facing half
            if (pp_bit(bitmap + (int64_t)m * W, (a * G + b) * G + c)) { nand[q][0] += o0; nand[q][1] += o1; nand[q][2] += o2; }
        }
    }
#pragma unroll
    for (int q = 0; q < 6; ++q)
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            int x = ncur[q][t], y = nand[q][t];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) { x += __shfl_xor_sync(0xffffffffu, x, d); y += __shfl_xor_sync(0xffffffffu, y, d); }
            if ((threadIdx.x & 31) == 0) { if (x) atomicAdd(&sc[q * 3 + t], x); if (y) atomicAdd(&sc[18 + q * 3 + t], y); }
        }
    __syncthreads();
    if (threadIdx.x < 36 && sc[threadIdx.x]) atomicAdd(&counts[n * 48 + threadIdx.x], sc[threadIdx.x]);
    if (threadIdx.x < 6 && blockIdx.y == 0) counts[n * 48 + 36 + threadIdx.x] = nb[threadIdx.x];
}

// cost accumulation in float16 exactly as numpy 1.13 evaluates `element_cost[t] += int` (float64 sum, rounded to float16 on
// assignment: adapthresh.py:141,162,165), first argmin, threshold update clamped to max_probThresh (166-168)
__global__ void pp_ada_update_kernel(const int32_t* __restrict__ counts, const int32_t* __restrict__ canon, int n_cubes, double beta,
                                     double max_thresh, double* __restrict__ thresh, int32_t* __restrict__ argmin_out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_cubes) return;
    if (!canon[n]) { if (argmin_out) argmin_out[n] = -1; return; }
    const int32_t* cn = counts + n * 48;
    __half cost[3] = {__double2half(0.0), __double2half(0.0), __double2half(0.0)};
    for (int q = 0; q < 6; ++q) {
        const int m = cn[36 + q];
        const int n_ovlp = m >= 0 ? counts[m * 48 + 42 + (q + 3) % 6] : 0;        // neighbour's half facing this cube (shift * -1)
        for (int t = 0; t < 3; ++t) {
            const int n_cur = cn[q * 3 + t];
            const int n_and = (n_cur == 0 || n_ovlp == 0) ? 0 : cn[18 + q * 3 + t];
            const int n_xor = n_cur + n_ovlp - 2 * n_and;
            cost[t] = __double2half((double)__half2float(cost[t]) + (double)n_xor);
            if (n_cur >= 6 && n_ovlp >= 6) cost[t] = __double2half((double)__half2float(cost[t]) - beta * (double)n_and);
        }
    }
    int arg = 0;
    float best = __half2float(cost[0]);
    for (int t = 1; t < 3; ++t) { const float x = __half2float(cost[t]); if (x < best) { best = x; arg = t; } }
    const double perturb = arg == 0 ? 0.1 : (arg == 1 ? 0.0 : -0.1);
    double nt = thresh[n] + perturb;
    if (!(nt <= max_thresh)) nt = max_thresh;                                     // python min(nt, max)
    thresh[n] = nt;
    if (argmin_out) argmin_out[n] = arg;
}

static inline unsigned pp_grid(int64_t n) { return (unsigned)std::max<int64_t>(1, cdiv(n, PP_THREADS)); }

// vox_cube + mask bitmap + prefix + dict; shared by denoise and adapthresh
static int post_prepare(const PostWs& w, const int32_t* cube_ijk, const int64_t* off, const uint8_t* ijk, const uint8_t* mask,
                        int n_cubes, int64_t n_vox, int G, cudaStream_t st) {
    SN_CUDA(cudaMemsetAsync(w.flags, 0, 4 * sizeof(int32_t), st));
    SN_CUDA(cudaMemsetAsync(w.bitmap, 0, (size_t)n_cubes * w.W * 4, st));
    SN_CUDA(cudaMemsetAsync(w.hkeys, 0, (size_t)w.H * 8, st));
    SN_CUDA(cudaMemsetAsync(w.hvals, 0xff, (size_t)w.H * 4, st));
    pp_vox_cube_kernel<<<pp_grid(n_vox), PP_THREADS, 0, st>>>(off, n_cubes, n_vox, w.vox_cube); SN_LAUNCHED();
    pp_bitmap_kernel<<<pp_grid(n_vox), PP_THREADS, 0, st>>>(ijk, mask, w.vox_cube, n_vox, G, w.W, w.bitmap, w.flags); SN_LAUNCHED();
    pp_prefix_kernel<<<n_cubes, PP_THREADS, 0, st>>>(w.bitmap, w.W, w.prefix, w.n_masked); SN_LAUNCHED();
    pp_hash_insert_kernel<<<pp_grid(n_cubes), PP_THREADS, 0, st>>>(cube_ijk, w.n_masked, n_cubes, w.hkeys, w.hvals, w.H, w.flags); SN_LAUNCHED();
    pp_canon_kernel<<<pp_grid((int64_t)n_cubes * 27), PP_THREADS, 0, st>>>(cube_ijk, w.n_masked, n_cubes, w.hkeys, w.hvals, w.H, w.canon, w.nb27); SN_LAUNCHED();
    return SN_OK;
}

static int post_check_flags(const PostWs& w, cudaStream_t st, const char* who) {
    int32_t f[4];
    SN_CUDA(cudaMemcpyAsync(f, w.flags, sizeof(f), cudaMemcpyDeviceToHost, st));
    SN_CUDA(cudaStreamSynchronize(st));
    SN_CHECK_ARG(!f[0], "%s: a masked voxel has a coordinate >= grid_extent", who);
    SN_CHECK_ARG(!f[1], "%s: cube ijk outside [0, 2^21)", who);
    return SN_OK;
}

}  // namespace sn

using namespace sn;

extern "C" int64_t sn_sparse_post_workspace_bytes(int n_cubes, int64_t n_vox, int grid_extent) {
    if (n_cubes < 0 || n_vox < 0 || grid_extent < 1 || grid_extent > 256) return -1;
    return post_layout(nullptr, 0, n_cubes, n_vox, grid_extent, nullptr) + 256;
}

extern "C" int sn_sparse_filter_voxels(const void* pred16_dev, const uint8_t* votes_dev, const int64_t* cube_offset_dev, int n_cubes,
                                       int64_t n_vox, const double* thresh_per_cube_dev, double thresh_scalar, int has_prob,
                                       int rayPool_thresh, int and_into, uint8_t* mask_inout_dev, void* workspace_dev,
                                       int64_t workspace_bytes, void* stream) {
    SN_CHECK_ARG(n_cubes >= 0 && n_vox >= 0, "sn_sparse_filter_voxels: negative size");
    if (n_vox == 0) return SN_OK;
    SN_CHECK_ARG(cube_offset_dev && mask_inout_dev && (!has_prob || pred16_dev), "sn_sparse_filter_voxels: NULL argument");
    SN_CHECK_ARG(workspace_dev && workspace_bytes >= align_up(n_vox * 4, 256), "sn_sparse_filter_voxels: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* vox_cube = (int32_t*)workspace_dev;
    pp_vox_cube_kernel<<<pp_grid(n_vox), PP_THREADS, 0, st>>>(cube_offset_dev, n_cubes, n_vox, vox_cube); SN_LAUNCHED();
    pp_filter_kernel<<<pp_grid(n_vox), PP_THREADS, 0, st>>>((const __half*)pred16_dev, votes_dev, vox_cube, n_vox, thresh_per_cube_dev,
                                                             thresh_scalar, has_prob, votes_dev ? rayPool_thresh : -1, and_into,
                                                             mask_inout_dev); SN_LAUNCHED();
    return SN_OK;
}

extern "C" int sn_sparse_denoise(const int32_t* cube_ijk_dev, const int64_t* cube_offset_dev, const uint8_t* ijk_dev,
                                 const uint8_t* mask_dev, int n_cubes, int64_t n_vox, int grid_extent, int D_cube, int neighbor_dist,
                                 uint8_t* keep_out_dev, uint32_t* labels_out_dev, int32_t* n_labels_out_dev, void* workspace_dev,
                                 int64_t workspace_bytes, void* stream) {
    SN_CHECK_ARG(n_cubes >= 0 && n_vox >= 0 && grid_extent >= 1 && grid_extent <= 256, "sn_sparse_denoise: bad sizes");
    SN_CHECK_ARG(neighbor_dist >= 1 && neighbor_dist <= 3, "sn_sparse_denoise: neighbor_dist must be 1, 2 or 3");
    SN_CHECK_ARG(D_cube >= 0, "sn_sparse_denoise: negative D_cube");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_cubes == 0) return SN_OK;
    SN_CHECK_ARG(cube_ijk_dev && cube_offset_dev && (n_vox == 0 || (ijk_dev && mask_dev)), "sn_sparse_denoise: NULL argument");
    PostWs w;
    const int64_t need = post_layout(workspace_dev, workspace_bytes, n_cubes, n_vox, grid_extent, &w);
    if (!workspace_dev || workspace_bytes < need) { set_error("sn_sparse_denoise: workspace %lld B < %lld B", (long long)workspace_bytes, (long long)need); return SN_ERR_NOMEM; }
    if (n_vox == 0) {
        if (n_labels_out_dev) SN_CUDA(cudaMemsetAsync(n_labels_out_dev, 0, (size_t)n_cubes * 4, st));
        return SN_OK;
    }
    const int G = grid_extent;
    int rc = post_prepare(w, cube_ijk_dev, cube_offset_dev, ijk_dev, mask_dev, n_cubes, n_vox, G, st);
    if (rc != SN_OK) return rc;
    SN_CUDA(cudaMemsetAsync(w.ovl, 0, (size_t)n_vox, st));
    pp_ccl_link_kernel<false><<<pp_grid(n_vox), PP_THREADS, 0, st>>>(ijk_dev, mask_dev, w.vox_cube, cube_offset_dev, n_vox, G, w.W, neighbor_dist,
                                                                      w.bitmap, w.prefix, w.parent); SN_LAUNCHED();
    pp_ccl_flatten_kernel<<<pp_grid(n_vox), PP_THREADS, 0, st>>>(cube_offset_dev, w.vox_cube, w.n_masked, n_vox, w.parent); SN_LAUNCHED();
    pp_ccl_link_kernel<true><<<pp_grid(n_vox), PP_THREADS, 0, st>>>(ijk_dev, mask_dev, w.vox_cube, cube_offset_dev, n_vox, G, w.W, neighbor_dist,
                                                                     w.bitmap, w.prefix, w.parent); SN_LAUNCHED();
    pp_ccl_root_kernel<<<pp_grid(n_vox), PP_THREADS, 0, st>>>(cube_offset_dev, w.vox_cube, w.n_masked, n_vox, w.parent, w.root); SN_LAUNCHED();
    if (labels_out_dev || n_labels_out_dev) {
        pp_label_rank_kernel<<<n_cubes, PP_THREADS, 0, st>>>(cube_offset_dev, w.n_masked, w.root, w.labelnum, n_labels_out_dev); SN_LAUNCHED();
        if (labels_out_dev) {
            pp_label_out_kernel<<<pp_grid(n_vox), PP_THREADS, 0, st>>>(ijk_dev, mask_dev, w.vox_cube, cube_offset_dev, n_vox, G, w.W, w.bitmap,
                                                                        w.prefix, w.root, w.labelnum, labels_out_dev); SN_LAUNCHED();
        }
    }
    if (keep_out_dev) {
        pp_overlap_kernel<<<pp_grid(n_vox), PP_THREADS, 0, st>>>(ijk_dev, mask_dev, w.vox_cube, cube_offset_dev, cube_ijk_dev, w.canon, n_vox, G,
                                                                  w.W, D_cube / 2, w.bitmap, w.prefix, w.root, w.nb27, w.ovl); SN_LAUNCHED();
        pp_keep_kernel<<<pp_grid(n_vox), PP_THREADS, 0, st>>>(ijk_dev, mask_dev, w.vox_cube, cube_offset_dev, n_vox, G, w.W, w.bitmap, w.prefix,
                                                               w.root, w.ovl, keep_out_dev); SN_LAUNCHED();
    }
    return post_check_flags(w, st, "sn_sparse_denoise");
}

extern "C" int sn_sparse_adapthresh(const int32_t* cube_ijk_dev, const int64_t* cube_offset_dev, const uint8_t* ijk_dev,
                                    const void* pred16_dev, const uint8_t* init_mask_dev, int n_cubes, int64_t n_vox, int grid_extent,
                                    int D_cube, double max_probThresh, double beta, int n_iter, double* thresh_inout_dev,
                                    uint8_t* mask_inout_dev, int32_t* argmin_out_dev, void* workspace_dev, int64_t workspace_bytes,
                                    void* stream) {
    SN_CHECK_ARG(n_cubes >= 0 && n_vox >= 0 && grid_extent >= 1 && grid_extent <= 256 && n_iter >= 0 && D_cube >= 0,
                 "sn_sparse_adapthresh: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_cubes == 0 || n_iter == 0) return SN_OK;
    SN_CHECK_ARG(cube_ijk_dev && cube_offset_dev && thresh_inout_dev && (n_vox == 0 || (ijk_dev && pred16_dev && init_mask_dev && mask_inout_dev)),
                 "sn_sparse_adapthresh: NULL argument");
    PostWs w;
    const int64_t need = post_layout(workspace_dev, workspace_bytes, n_cubes, n_vox, grid_extent, &w);
    if (!workspace_dev || workspace_bytes < need) { set_error("sn_sparse_adapthresh: workspace %lld B < %lld B", (long long)workspace_bytes, (long long)need); return SN_ERR_NOMEM; }
    const int G = grid_extent, Dmid = D_cube / 2;
    const __half* pred = (const __half*)pred16_dev;
    // the dict of cubes that are non-empty under the INITIAL mask (adapthresh.py:121-124); fixed for all iterations
    int rc = post_prepare(w, cube_ijk_dev, cube_offset_dev, ijk_dev, init_mask_dev, n_cubes, n_vox, G, st);
    if (rc != SN_OK) return rc;
    for (int it = 0; it < n_iter; ++it) {
        SN_CUDA(cudaMemsetAsync(w.bitmap, 0, (size_t)n_cubes * w.W * 4, st));
        SN_CUDA(cudaMemsetAsync(w.counts, 0, (size_t)n_cubes * 48 * 4, st));
        pp_ada_occ0_kernel<<<dim3(n_cubes, PP_OCC_SLICES), PP_THREADS, 0, st>>>(ijk_dev, pred, mask_inout_dev, cube_offset_dev, thresh_inout_dev, G, w.W, D_cube, Dmid,
                                                           w.bitmap, w.counts, w.flags); SN_LAUNCHED();
        pp_ada_count_kernel<<<dim3(n_cubes, PP_SLICES), PP_THREADS, 0, st>>>(ijk_dev, pred, mask_inout_dev, cube_offset_dev, cube_ijk_dev, w.canon,
                                                                             thresh_inout_dev, G, w.W, D_cube, Dmid, w.bitmap, w.nb27, w.counts); SN_LAUNCHED();
        pp_ada_update_kernel<<<pp_grid(n_cubes), PP_THREADS, 0, st>>>(w.counts, w.canon, n_cubes, beta, max_probThresh, thresh_inout_dev,
                                                                      argmin_out_dev ? argmin_out_dev + (int64_t)it * n_cubes : nullptr); SN_LAUNCHED();
        if (n_vox > 0) {
            pp_filter_kernel<<<pp_grid(n_vox), PP_THREADS, 0, st>>>(pred, nullptr, w.vox_cube, n_vox, thresh_inout_dev, 0.0, 1, -1, 1,
                                                                     mask_inout_dev); SN_LAUNCHED();           // adapthresh.py:174
        }
    }
    return post_check_flags(w, st, "sn_sparse_adapthresh");
}
