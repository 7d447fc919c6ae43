// bvh_build.cuh -- GPU BVH construction: replaces rtcCommitScene (pg1/raytracer.cpp:127).
//
// Pipeline (all on the device, one stream, no host round trips except the final stats read):
//   K1  k_scene_bounds      centroid bounds of all triangles (block reduce + ordered-int atomics)
//   K1b k_morton            63-bit Morton key of each centroid (21 bits / axis) + identity permutation
//   K2  radix sort          LSD, 8 bits / pass, stable (equal keys keep flat-id order => deterministic tree)
//   K3  k_karras            Karras 2012 hierarchy emit over the sorted keys (N-1 internal nodes)
//   K3b k_refit             bottom-up AABB fit with one atomic flag per internal node
//   K5  k_emit_*            collapse to the traversal layout + triangle re-layout (3 x float4 per triangle)
#pragma once
#include "common.cuh"

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

// exclusive scan of `n` uint32 in place, single block (n = 256 * n_tiles, a few 1e5 at most)
__global__ void __launch_bounds__(1024) k_rs_scan(uint32_t* __restrict__ data, uint32_t n) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
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
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                           uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                                                           const uint32_t* __restrict__ hist_scanned, uint32_t n_tiles) {
    __shared__ uint32_t cnt[RS_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
        uint32_t run = hist_scanned[(size_t)d * n_tiles + blockIdx.x];
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

// ---------------------------------------------------------------------------------------------------
// K5a: triangle re-layout into leaf order: (v0, flat id) (e1 = v0 - v1, 0) (e2 = v2 - v0, 0); 48 B, 16-B aligned.
__global__ void __launch_bounds__(256) k_emit_tris(const float* __restrict__ pos, const uint32_t* __restrict__ vals, uint32_t n, float4* __restrict__ tris) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t id = vals[k];
    const float* p = pos + 9 * (size_t)id;
    const V3 v0 = v3(p[0], p[1], p[2]), v1 = v3(p[3], p[4], p[5]), v2 = v3(p[6], p[7], p[8]);
    const V3 e1 = v0 - v1, e2 = v2 - v0;
    tris[3 * (size_t)k + 0] = make_float4(v0.x, v0.y, v0.z, __uint_as_float(id));
    tris[3 * (size_t)k + 1] = make_float4(e1.x, e1.y, e1.z, 0.0f);
    tris[3 * (size_t)k + 2] = make_float4(e2.x, e2.y, e2.z, 0.0f);
}

// K5b (binary layout): node i = 4 x float4
//   n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)   n1 = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//   n2 = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)   n3 = (ref0, ref1, 0, 0) as int bits
//   ref >= 0: internal node index; ref < 0: leaf, ~ref = (first triangle << 2) | (count - 1), count <= PGRT_LEAF_MAX.
__host__ __device__ __forceinline__ int leaf_ref(int first, int count) { return ~((first << 2) | (count - 1)); }

__global__ void __launch_bounds__(256) k_emit_bvh2(int n, BinTree t, float4* __restrict__ nodes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int ref[2];
    const int ch[2] = {t.left[i], t.right[i]};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (ch[c] >= n - 1) ref[c] = leaf_ref(ch[c] - (n - 1), 1);
        else {
            const int f = t.first[ch[c]], l = t.last[ch[c]];
            ref[c] = (l - f + 1 <= PGRT_LEAF_MAX) ? leaf_ref(f, l - f + 1) : ch[c];
        }
    }
    const float* l0 = t.lo + 3 * (size_t)ch[0]; const float* h0 = t.hi + 3 * (size_t)ch[0];
    const float* l1 = t.lo + 3 * (size_t)ch[1]; const float* h1 = t.hi + 3 * (size_t)ch[1];
    nodes[4 * (size_t)i + 0] = make_float4(l0[0], h0[0], l0[1], h0[1]);
    nodes[4 * (size_t)i + 1] = make_float4(l1[0], h1[0], l1[1], h1[1]);
    nodes[4 * (size_t)i + 2] = make_float4(l0[2], h0[2], l1[2], h1[2]);
    nodes[4 * (size_t)i + 3] = make_float4(__int_as_float(ref[0]), __int_as_float(ref[1]), 0.0f, 0.0f);
}

// SAH cost of the emitted binary tree: sum over reachable internal nodes of area(child)/area(root) * (leaf ? count : 1)
__global__ void __launch_bounds__(256) k_sah_cost(int n, BinTree t, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float c = 0.0f;
    if (i < n - 1 && (t.last[i] - t.first[i] + 1) > PGRT_LEAF_MAX) {
        const float rx = t.hi[0] - t.lo[0], ry = t.hi[1] - t.lo[1], rz = t.hi[2] - t.lo[2];
        const float ra = rx * ry + ry * rz + rz * rx;
        const int ch[2] = {t.left[i], t.right[i]};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float* l = t.lo + 3 * (size_t)ch[k]; const float* h = t.hi + 3 * (size_t)ch[k];
            const float dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2];
            const float a = dx * dy + dy * dz + dz * dx;
            int cnt = 1;
            if (ch[k] < n - 1) { const int m = t.last[ch[k]] - t.first[ch[k]] + 1; cnt = m <= PGRT_LEAF_MAX ? m : 1; }
            c += (ra > 0.0f ? a / ra : 0.0f) * (float)cnt;
        }
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c != 0.0f) atomicAdd(out, c);
}
