// bvh_build.cuh -- GPU BVH construction: replaces rtcCommitScene (pg1/raytracer.cpp:127).
//
// Pipeline (all on the device, one stream, no host round trips except the final stats read):
//   K1  k_scene_bounds      centroid bounds of all triangles (block reduce + ordered-int atomics)
//   K1b k_morton            63-bit Morton key of each centroid (21 bits / axis) + identity permutation
//   K2  radix sort          LSD, 8 bits / pass, stable (equal keys keep flat-id order => deterministic tree)
//   K4  k_ploc_*            PLOC: agglomerative refinement of the Morton order into a binary tree (default)
//   K3  k_karras, k_refit   Karras 2012 hierarchy over the sorted keys + bottom-up AABB fit (PGRT_BUILDER=lbvh, for comparison)
//   K5  k_collapse          collapse to the 8-wide traversal layout + triangle re-layout (3 x float4 per triangle), bvh8.cuh
#pragma once
#include "common.cuh"
#include "bvh8.cuh"

// ---------------------------------------------------------------------------------------------------
// ordered-int float atomics
__device__ __forceinline__ int float_to_ordered(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

struct SceneBounds { int lo[3]; int hi[3]; };   // ordered-int encoded centroid bounds

__global__ void k_init_bounds(SceneBounds* b) {
    if (threadIdx.x < 3) { b->lo[threadIdx.x] = float_to_ordered(FLT_MAX); b->hi[threadIdx.x] = float_to_ordered(-FLT_MAX); }
}

__device__ __forceinline__ void tri_aabb(const float* p, float lo[3], float hi[3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = fminf(p[a], fminf(p[3 + a], p[6 + a]));
        hi[a] = fmaxf(p[a], fmaxf(p[3 + a], p[6 + a]));
    }
}

__global__ void __launch_bounds__(256) k_scene_bounds(const float* __restrict__ pos, uint32_t n, SceneBounds* out) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float l[3], h[3];
        tri_aabb(pos + 9 * (size_t)i, l, h);
#pragma unroll
        for (int a = 0; a < 3; ++a) { const float c = 0.5f * (l[a] + h[a]); lo[a] = fminf(lo[a], c); hi[a] = fmaxf(hi[a], c); }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(&out->lo[a], float_to_ordered(lo[a])); atomicMax(&out->hi[a], float_to_ordered(hi[a])); }
    }
}

__device__ __forceinline__ uint64_t expand21(uint64_t v) {   // spread 21 bits to every third bit
    v &= 0x1FFFFFull;
    v = (v | v << 32) & 0x1F00000000FFFFull;
    v = (v | v << 16) & 0x1F0000FF0000FFull;
    v = (v | v << 8) & 0x100F00F00F00F00Full;
    v = (v | v << 4) & 0x10C30C30C30C30C3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

__global__ void __launch_bounds__(256) k_morton(const float* __restrict__ pos, uint32_t n, const SceneBounds* __restrict__ sb,
                                                uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float l[3], h[3];
    tri_aabb(pos + 9 * (size_t)i, l, h);
    uint64_t q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float blo = ordered_to_float(sb->lo[a]), bhi = ordered_to_float(sb->hi[a]);
        const float ext = bhi - blo;
        const float c = 0.5f * (l[a] + h[a]);
        float f = ext > 0.0f ? (c - blo) / ext : 0.0f;
        f = fminf(fmaxf(f, 0.0f), 1.0f);
        uint32_t g = (uint32_t)(f * 2097152.0f);
        q[a] = g > 2097151u ? 2097151u : g;
    }
    keys[i] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
    vals[i] = i;
}

// ---------------------------------------------------------------------------------------------------
// K2: LSD radix sort, 64-bit keys / 32-bit values, 8 bits per pass.
//   tile = 256 threads x 16 keys; each warp owns a contiguous 512-key run so that warp-level ranking
//   (match_any) + warp-major prefix is stable.
#define RS_THREADS 256
#define RS_ITEMS 16
#define RS_TILE (RS_THREADS * RS_ITEMS)
#define RS_WARPS (RS_THREADS / 32)

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t* __restrict__ keys, uint32_t n, int shift, uint32_t* __restrict__ hist,
                                                        uint32_t n_tiles) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const uint32_t i = base + k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];   // digit-major for the scan
}

// exclusive scan of every digit's row of the histogram (n_tiles counts, digit-major) in place, one block per digit; the row's
// total goes to totals[digit].  (One block over all 256 * n_tiles counts took 290 us per pass at 10 M keys: 2.3 of the 13 ms build.)
__global__ void __launch_bounds__(1024) k_rs_scan_rows(uint32_t* __restrict__ hist, uint32_t n, uint32_t* __restrict__ totals) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    uint32_t* data = hist + (size_t)blockIdx.x * n;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const uint32_t per_iter = 1024 * 4;
    for (uint32_t base = 0; base < n; base += per_iter) {
        const uint32_t i0 = base + threadIdx.x * 4;
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (i0 + k < n) ? data[i0 + k] : 0u;
        const uint32_t tsum = v[0] + v[1] + v[2] + v[3];
        uint32_t incl = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_sums[threadIdx.x], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (threadIdx.x >= o) wi += t; }
            warp_sums[threadIdx.x] = wi - w;   // exclusive
        }
        __syncthreads();
        uint32_t excl = carry + warp_sums[threadIdx.x >> 5] + incl - tsum;
#pragma unroll
        for (int k = 0; k < 4; ++k) { if (i0 + k < n) data[i0 + k] = excl; excl += v[k]; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                           uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                                                           const uint32_t* __restrict__ hist_scanned, const uint32_t* __restrict__ totals, uint32_t n_tiles) {
    __shared__ uint32_t cnt[RS_WARPS][256];
    __shared__ uint32_t dsum[RS_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // where digit d starts in the output: exclusive prefix of the digit totals (RS_THREADS = 256 = one thread per digit)
    uint32_t dbase;
    {
        const uint32_t t = totals[threadIdx.x];
        uint32_t incl = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) dsum[warp] = incl;
        __syncthreads();
        uint32_t before = 0;
        for (int w = 0; w < warp; ++w) before += dsum[w];
        dbase = before + incl - t;
    }
    for (int k = threadIdx.x; k < RS_WARPS * 256; k += RS_THREADS) (&cnt[0][0])[k] = 0;
    __syncthreads();
    const uint32_t wbase = blockIdx.x * RS_TILE + warp * (32 * RS_ITEMS);
    uint64_t key[RS_ITEMS];
    uint32_t rank[RS_ITEMS];
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const uint32_t i = wbase + k * 32 + lane;
        const bool valid = i < n;
        key[k] = valid ? keys_in[i] : ~0ull;
        const uint32_t d = valid ? ((uint32_t)(key[k] >> shift) & 255u) : 256u;   // 256: matches only other invalid lanes
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if (valid && lane == leader) { prev = cnt[warp][d]; cnt[warp][d] = prev + __popc(peers); }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[k] = prev + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    {   // exclusive prefix over warps per digit, folded with the tile's global base
        const uint32_t d = threadIdx.x;
        uint32_t run = dbase + hist_scanned[(size_t)d * n_tiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { const uint32_t c = cnt[w][d]; cnt[w][d] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const uint32_t i = wbase + k * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t)(key[k] >> shift) & 255u;
            const uint32_t dst = cnt[warp][d] + rank[k];
            keys_out[dst] = key[k];
            vals_out[dst] = vals_in[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K3: Karras 2012.  Internal nodes 0..N-2, leaves N-1..2N-2 (leaf k = sorted position k).
struct BinTree {
    int* left;        // [N-1] child node ids
    int* right;       // [N-1]
    int* parent;      // [2N-1]
    int* first;       // [N-1] first sorted position covered
    int* last;        // [N-1]
    float* lo;        // [2N-1][3]
    float* hi;        // [2N-1][3]
    int* flag;        // [N-1]
};

__device__ __forceinline__ int kdelta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

__global__ void __launch_bounds__(256) k_karras(const uint64_t* __restrict__ keys, int n, BinTree t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (kdelta(keys, n, i, i + 1) - kdelta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = kdelta(keys, n, i, i - d);
    int lmax = 2;
    while (kdelta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int s = lmax >> 1; s >= 1; s >>= 1)
        if (kdelta(keys, n, i, i + (l + s) * d) > dmin) l += s;
    const int j = i + l * d;
    const int dnode = kdelta(keys, n, i, j);
    int s = 0, tt = l;
    do {
        tt = (tt + 1) >> 1;
        if (kdelta(keys, n, i, i + (s + tt) * d) > dnode) s += tt;
    } while (tt > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int lc = (lo == gamma) ? (n - 1 + gamma) : gamma;
    const int rc = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    t.left[i] = lc; t.right[i] = rc; t.first[i] = lo; t.last[i] = hi;
    t.parent[lc] = i; t.parent[rc] = i;
    if (i == 0) t.parent[0] = -1;
    t.flag[i] = 0;
}

__global__ void __launch_bounds__(256) k_refit(const float* __restrict__ pos, const uint32_t* __restrict__ vals, int n, BinTree t) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float lo[3], hi[3];
    tri_aabb(pos + 9 * (size_t)vals[k], lo, hi);
    int node = n - 1 + k;
#pragma unroll
    for (int a = 0; a < 3; ++a) { t.lo[3 * (size_t)node + a] = lo[a]; t.hi[3 * (size_t)node + a] = hi[a]; }
    int p = t.parent[node];
    while (p >= 0) {
        __threadfence();
        if (atomicAdd(&t.flag[p], 1) == 0) return;   // first arrival: the sibling finishes this node
        __threadfence();
        const int l = t.left[p], r = t.right[p];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float bl = fminf(__ldcg(&t.lo[3 * (size_t)l + a]), __ldcg(&t.lo[3 * (size_t)r + a]));
            const float bh = fmaxf(__ldcg(&t.hi[3 * (size_t)l + a]), __ldcg(&t.hi[3 * (size_t)r + a]));
            t.lo[3 * (size_t)p + a] = bl; t.hi[3 * (size_t)p + a] = bh;
        }
        p = t.parent[p];
    }
}

// ===================================================================================================
// K4: PLOC -- agglomerative refinement of the Morton order (Meister & Bittner, "Parallel Locally-Ordered Clustering
// for BVH Construction", TVCG 2018).  Clusters start as the Morton-sorted triangles; every pass each cluster finds,
// among its PLOC_R neighbours on either side in the current order, the one whose merged box has the smallest
// surface area; mutual pairs merge into a new binary node, everything else is carried over, and the array is
// compacted in order.  The minimum of (area, lower index, higher index) over all pairs in range is always mutual,
// so every pass merges at least one pair.  Node ids are assigned by prefix sums, so the tree is deterministic.
//
// Binary tree layout shared with the collapse (Bvh2View in bvh8.cuh): leaves 0..N-1, internal nodes N..2N-2 in
// creation order (root = 2N-2), b0 = (lo, left), b1 = (hi, right), count = triangles below.
#define PLOC_R 8        // search radius.  8, not the 16 of round 1: better trees on every scene tried AND half the search (profiles/r2_ploc_radius.txt);
                        // tests/emul/bvh8_emul.cpp carries the same constant (the GPU builder must reproduce the emulator's tree)
#define PLOC_THREADS 256

__global__ void __launch_bounds__(256) k_ploc_init(const float* __restrict__ pos, const uint32_t* __restrict__ vals, uint32_t n,
                                                   float4* __restrict__ b0, float4* __restrict__ b1, uint32_t* __restrict__ count, uint32_t* __restrict__ cid) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float lo[3], hi[3];
    tri_aabb(pos + 9 * (size_t)vals[k], lo, hi);
    b0[k] = make_float4(lo[0], lo[1], lo[2], __uint_as_float(0xFFFFFFFFu));
    b1[k] = make_float4(hi[0], hi[1], hi[2], __uint_as_float(0xFFFFFFFFu));
    count[k] = 1u; cid[k] = k;
}

__device__ __forceinline__ float merged_half_area(const float* a_lo, const float* a_hi, float blx, float bly, float blz, float bhx, float bhy, float bhz) {
    const float dx = fmaxf(a_hi[0], bhx) - fminf(a_lo[0], blx), dy = fmaxf(a_hi[1], bhy) - fminf(a_lo[1], bly), dz = fmaxf(a_hi[2], bhz) - fminf(a_lo[2], blz);
    return dx * dy + dy * dz + dz * dx;
}

__global__ void __launch_bounds__(PLOC_THREADS) k_ploc_nn(const float4* __restrict__ b0, const float4* __restrict__ b1, const uint32_t* __restrict__ cid,
                                                          uint32_t n, int* __restrict__ nn, int ties) {
    __shared__ float sb[6][PLOC_THREADS + 2 * PLOC_R];
    const long first = (long)blockIdx.x * PLOC_THREADS - PLOC_R;
    for (int k = threadIdx.x; k < PLOC_THREADS + 2 * PLOC_R; k += PLOC_THREADS) {
        const long g = first + k;
        if (g >= 0 && g < (long)n) {
            const uint32_t node = cid[g];
            const float4 lo = b0[node], hi = b1[node];
            sb[0][k] = lo.x; sb[1][k] = lo.y; sb[2][k] = lo.z; sb[3][k] = hi.x; sb[4][k] = hi.y; sb[5][k] = hi.z;
        }
    }
    __syncthreads();
    const long i = (long)blockIdx.x * PLOC_THREADS + threadIdx.x;
    if (i >= (long)n) return;
    const int me = threadIdx.x + PLOC_R;
    const float lo[3] = {sb[0][me], sb[1][me], sb[2][me]}, hi[3] = {sb[3][me], sb[4][me], sb[5][me]};
    PlocBest best; ploc_best_init(best);
#pragma unroll 4
    for (int off = -PLOC_R; off <= PLOC_R; ++off) {
        const long j = i + off;
        if (off == 0 || j < 0 || j >= (long)n) continue;
        const int k = me + off;
        const float a = merged_half_area(lo, hi, sb[0][k], sb[1][k], sb[2][k], sb[3][k], sb[4][k], sb[5][k]);
        ploc_offer(best, a, i, j, ties);
    }
    nn[i] = best.j;
}

// flags of cluster i: bit 0 = survives into the next pass, bit 1 = creates a node (the lower index of a mutual pair)
__device__ __forceinline__ uint32_t ploc_flags(const int* __restrict__ nn, uint32_t i, uint32_t n) {
    if (i >= n) return 0u;
    const int j = nn[i];
    const bool mutual = j >= 0 && nn[j] == (int)i;
    if (mutual) return (uint32_t)j > i ? 3u : 0u;
    return 1u;
}

__global__ void __launch_bounds__(PLOC_THREADS) k_ploc_count(const int* __restrict__ nn, uint32_t n, uint2* __restrict__ block_sums) {
    const uint32_t f = ploc_flags(nn, blockIdx.x * PLOC_THREADS + threadIdx.x, n);
    const int v = __syncthreads_count(f & 1u), m = __syncthreads_count(f & 2u);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = make_uint2((uint32_t)v, (uint32_t)m);
}

// exclusive scan of the per-block (survivors, merges) in place; totals -> out[0]
__global__ void __launch_bounds__(1024) k_ploc_scan(uint2* __restrict__ sums, uint32_t nb, uint2* __restrict__ totals) {
    __shared__ uint2 wsum[32];
    __shared__ uint2 carry;
    if (threadIdx.x == 0) carry = make_uint2(0u, 0u);
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint2 v = i < nb ? sums[i] : make_uint2(0u, 0u);
        uint2 inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t tx = __shfl_up_sync(0xffffffffu, inc.x, o), ty = __shfl_up_sync(0xffffffffu, inc.y, o);
            if ((threadIdx.x & 31) >= o) { inc.x += tx; inc.y += ty; }
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint2 w = wsum[threadIdx.x]; uint2 wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t tx = __shfl_up_sync(0xffffffffu, wi.x, o), ty = __shfl_up_sync(0xffffffffu, wi.y, o);
                if (threadIdx.x >= o) { wi.x += tx; wi.y += ty; }
            }
            wsum[threadIdx.x] = make_uint2(wi.x - w.x, wi.y - w.y);
        }
        __syncthreads();
        const uint2 c = carry, ws = wsum[threadIdx.x >> 5];
        if (i < nb) sums[i] = make_uint2(c.x + ws.x + inc.x - v.x, c.y + ws.y + inc.y - v.y);
        __syncthreads();
        if (threadIdx.x == 1023) carry = make_uint2(c.x + ws.x + inc.x, c.y + ws.y + inc.y);
        __syncthreads();
    }
    if (threadIdx.x == 0) *totals = carry;
}

__global__ void __launch_bounds__(PLOC_THREADS) k_ploc_apply(const int* __restrict__ nn, const uint32_t* __restrict__ cid_in, uint32_t n,
                                                             const uint2* __restrict__ block_offs, uint32_t next_node,
                                                             float4* __restrict__ b0, float4* __restrict__ b1, uint32_t* __restrict__ count,
                                                             uint32_t* __restrict__ cid_out) {
    __shared__ uint32_t wv[PLOC_THREADS / 32], wm[PLOC_THREADS / 32];
    const uint32_t i = blockIdx.x * PLOC_THREADS + threadIdx.x;
    const uint32_t f = ploc_flags(nn, i, n);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bv = __ballot_sync(0xffffffffu, f & 1u), bm = __ballot_sync(0xffffffffu, f & 2u);
    if (lane == 0) { wv[warp] = __popc(bv); wm[warp] = __popc(bm); }
    __syncthreads();
    uint32_t pv = 0, pm = 0;
    for (int w = 0; w < warp; ++w) { pv += wv[w]; pm += wm[w]; }
    if (!(f & 1u)) return;
    const uint2 off = block_offs[blockIdx.x];
    const uint32_t dst = off.x + pv + __popc(bv & ((1u << lane) - 1u));
    uint32_t node = cid_in[i];
    if (f & 2u) {
        const uint32_t other = cid_in[nn[i]];
        const uint32_t id = next_node + off.y + pm + __popc(bm & ((1u << lane) - 1u));
        const float4 al = b0[node], ah = b1[node], bl = b0[other], bh = b1[other];
        b0[id] = make_float4(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z), __uint_as_float(node));
        b1[id] = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), __uint_as_float(other));
        count[id] = count[node] + count[other];
        node = id;
    }
    cid_out[dst] = node;
}

// the last passes (n <= PLOC_FINISH clusters) in one block, boxes and ids in shared memory
#define PLOC_FINISH 512
__global__ void __launch_bounds__(PLOC_FINISH) k_ploc_finish(const uint32_t* __restrict__ cid_in, uint32_t n, uint32_t next_node,
                                                             float4* __restrict__ b0, float4* __restrict__ b1, uint32_t* __restrict__ count,
                                                             uint32_t* __restrict__ left_over, int ties) {
    __shared__ uint32_t s_cid[2][PLOC_FINISH], s_cnt[2][PLOC_FINISH];
    __shared__ float s_box[2][6][PLOC_FINISH];
    __shared__ int s_nn[PLOC_FINISH];
    __shared__ uint32_t wv[PLOC_FINISH / 32], wm[PLOC_FINISH / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int cur = 0;
    if ((uint32_t)tid < n) {
        const uint32_t node = cid_in[tid];
        const float4 lo = b0[node], hi = b1[node];
        s_cid[0][tid] = node; s_cnt[0][tid] = count[node];
        s_box[0][0][tid] = lo.x; s_box[0][1][tid] = lo.y; s_box[0][2][tid] = lo.z; s_box[0][3][tid] = hi.x; s_box[0][4][tid] = hi.y; s_box[0][5][tid] = hi.z;
    }
    __syncthreads();
    int mode = ties;
    while (n > 1) {
        if ((uint32_t)tid < n) {
            const float lo[3] = {s_box[cur][0][tid], s_box[cur][1][tid], s_box[cur][2][tid]}, hi[3] = {s_box[cur][3][tid], s_box[cur][4][tid], s_box[cur][5][tid]};
            PlocBest best; ploc_best_init(best);
            for (int off = -PLOC_R; off <= PLOC_R; ++off) {
                const int j = tid + off;
                if (off == 0 || j < 0 || j >= (int)n) continue;
                const float a = merged_half_area(lo, hi, s_box[cur][0][j], s_box[cur][1][j], s_box[cur][2][j], s_box[cur][3][j], s_box[cur][4][j], s_box[cur][5][j]);
                ploc_offer(best, a, tid, j, mode);
            }
            s_nn[tid] = best.j;
        }
        __syncthreads();
        const uint32_t f = ploc_flags(s_nn, (uint32_t)tid, n);
        const unsigned bv = __ballot_sync(0xffffffffu, f & 1u), bm = __ballot_sync(0xffffffffu, f & 2u);
        if (lane == 0) { wv[warp] = __popc(bv); wm[warp] = __popc(bm); }
        __syncthreads();
        uint32_t pv = 0, pm = 0, tv = 0, tmg = 0;
        for (int w = 0; w < PLOC_FINISH / 32; ++w) { if (w < warp) { pv += wv[w]; pm += wm[w]; } tv += wv[w]; tmg += wm[w]; }
        if (mode != PLOC_TIES_BUDDY && ploc_pass_stalled(tmg, n)) { mode = PLOC_TIES_BUDDY; __syncthreads(); continue; }   // redo this pass (see bvh8.cuh)
        if (f & 1u) {
            const uint32_t dst = pv + __popc(bv & ((1u << lane) - 1u));
            uint32_t node = s_cid[cur][tid], cnt = s_cnt[cur][tid];
            float bx[6];
#pragma unroll
            for (int a = 0; a < 6; ++a) bx[a] = s_box[cur][a][tid];
            if (f & 2u) {
                const int j = s_nn[tid];
                const uint32_t other = s_cid[cur][j];
                const uint32_t id = next_node + pm + __popc(bm & ((1u << lane) - 1u));
#pragma unroll
                for (int a = 0; a < 3; ++a) { bx[a] = fminf(bx[a], s_box[cur][a][j]); bx[3 + a] = fmaxf(bx[3 + a], s_box[cur][3 + a][j]); }
                cnt += s_cnt[cur][j];
                b0[id] = make_float4(bx[0], bx[1], bx[2], __uint_as_float(node));
                b1[id] = make_float4(bx[3], bx[4], bx[5], __uint_as_float(other));
                count[id] = cnt;
                node = id;
            }
            s_cid[cur ^ 1][dst] = node; s_cnt[cur ^ 1][dst] = cnt;
#pragma unroll
            for (int a = 0; a < 6; ++a) s_box[cur ^ 1][a][dst] = bx[a];
        }
        __syncthreads();
        if (tmg == 0u) break;   // no mutual pair: only possible with non-finite boxes (NaN areas never compare); reported by the host
        n = tv; next_node += tmg; cur ^= 1; mode = ties;
    }
    if (tid == 0) *left_over = n;   // 1 = a single root remains
}

// the Karras tree (k_karras + k_refit) in the layout above: leaf k -> node k, internal j -> node N + j (root = N)
__global__ void __launch_bounds__(256) k_lbvh_to_b2(int n, BinTree t, const uint32_t* __restrict__ vals, float4* __restrict__ b0, float4* __restrict__ b1,
                                                    uint32_t* __restrict__ count) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 2 * n - 1) return;
    // source numbering: internal 0..n-2, leaves n-1..2n-2
    const bool leaf = k >= n - 1;
    const int dst = leaf ? k - (n - 1) : n + k;
    auto remap = [n](int c) { return (uint32_t)(c >= n - 1 ? c - (n - 1) : n + c); };
    const float* lo = t.lo + 3 * (size_t)k; const float* hi = t.hi + 3 * (size_t)k;
    b0[dst] = make_float4(lo[0], lo[1], lo[2], __uint_as_float(leaf ? 0xFFFFFFFFu : remap(t.left[k])));
    b1[dst] = make_float4(hi[0], hi[1], hi[2], __uint_as_float(leaf ? 0xFFFFFFFFu : remap(t.right[k])));
    count[dst] = leaf ? 1u : (uint32_t)(t.last[k] - t.first[k] + 1);
}

// ===================================================================================================
// K5: collapse to the wide layout, one tree level per launch, one thread per wide node (bvh8_gather / bvh8_emit).
// work item = (binary root, wide node index); children of a node are contiguous (child_base from one atomicAdd),
// as are its leaf triangles (tri_base).  The hit record never depends on the layout (ties resolve by flat id).
struct CollapseCounters { uint32_t nodes, tris, next_items, pad; float sah; };

__global__ void __launch_bounds__(128) k_collapse(Bvh2View t, const float* __restrict__ pos, const uint2* __restrict__ items, uint32_t n_items,
                                                  uint2* __restrict__ next_items, CollapseCounters* __restrict__ cc,
                                                  float4* __restrict__ nodes, float4* __restrict__ tris, int layout) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    float sah = 0.0f;
    if (k < n_items) {
        const uint2 it = items[k];
        Wide8 w;
        bvh8_gather(t, it.x, w);
        int n_int, n_tri;
        bvh8_counts(t, w, n_int, n_tri);
        const uint32_t child_base = n_int ? atomicAdd(&cc->nodes, (uint32_t)n_int) : 0u;
        const uint32_t tri_base = n_tri ? atomicAdd(&cc->tris, (uint32_t)n_tri) : 0u;
        uint32_t ic[8];
        bvh8_emit(t, it.x, w, child_base, tri_base, pos, nodes + (size_t)(layout == PGRT_LAYOUT_F32 ? PGRT_NODE_F4_F32 : PGRT_NODE_F4_Q8) * it.y, tris, ic, sah, layout);
        if (n_int) {
            const uint32_t q = atomicAdd(&cc->next_items, (uint32_t)n_int);
            for (int r = 0; r < n_int; ++r) next_items[q + r] = make_uint2(ic[r], child_base + (uint32_t)r);
        }
    }
    for (int o = 16; o > 0; o >>= 1) sah += __shfl_xor_sync(0xffffffffu, sah, o);
    if ((threadIdx.x & 31) == 0 && sah != 0.0f) atomicAdd(&cc->sah, sah);
}
