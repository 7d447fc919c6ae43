// common.cuh -- shared declarations of the B200 render-loop library (libpgrt_b200.so).
//
// Arithmetic contract of every kernel in this library (the oracle states the same contract on its own):
//   * compiled with -fmad=false -ftz=true -prec-div=true -prec-sqrt=true: no implicit FMA contraction
//     (the reference is MSVC /fp:precise x64), IEEE division and square root, FTZ/DAZ as the reference's
//     main() sets them (pg1_embree.cpp:8-9).  FMA appears only where written: e_dot / e_cross / interpolate
//     (Embree's AVX2 madd/msub forms) and the BVH slab tests (conservative culling, never result-bearing).
//   * float transcendentals the reference calls through float overloads (expf, powf, atan2f, asinf) are
//     evaluated in double and rounded once, which agrees with a correctly rounding libm; double pow stays double.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include <math.h>
#include <string.h>
#include "../../include/pgrt.h"

#define PGRT_TILE_W 32
#define PGRT_TILE_H 8
#define PGRT_TILE_PIXELS (PGRT_TILE_W * PGRT_TILE_H)
#define PGRT_MAX_LEVELS 33            // max_depth <= 32

struct V3 { float x, y, z; };

__host__ __device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
// vector3.cpp:99-122
__host__ __device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
__host__ __device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(s * a.x, s * a.y, s * a.z); }
__host__ __device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
// vector3.cpp:56-59, :38-44, :19-36
__host__ __device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ V3 cross3(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__host__ __device__ __forceinline__ float sqr_norm3(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__host__ __device__ __forceinline__ float l2norm3(V3 a) { return sqrtf(sqr_norm3(a)); }
__host__ __device__ __forceinline__ V3 normalize3(V3 a) {
    const float n = sqr_norm3(a);
    if (n != 0) { const float rn = 1 / sqrtf(n); a.x *= rn; a.y *= rn; a.z *= rn; }
    return a;
}
// row-major Matrix3x3 (matrix3x3.cpp:32-45, :68-73)
struct M3 { float m00, m01, m02, m10, m11, m12, m20, m21, m22; };
__host__ __device__ __forceinline__ V3 mul3(const M3& a, V3 b) {
    return v3(a.m00 * b.x + a.m01 * b.y + a.m02 * b.z, a.m10 * b.x + a.m11 * b.y + a.m12 * b.z, a.m20 * b.x + a.m21 * b.y + a.m22 * b.z);
}

// ---- host/device portability of the BVH code: the wide-node encoder and the traversal (bvh8.cuh, traverse.cuh) are
// plain functions that also compile with g++, so tests/ can run them on the CPU (tests/emul/) before a GPU is spent.
#if defined(__CUDACC__)
#define PG_HD __host__ __device__ __forceinline__
#else
#define PG_HD inline
#endif
PG_HD float pg_fma(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);   // one correctly rounded operation on both sides (host builds use -mfma -ffp-contract=off)
#endif
}
PG_HD uint32_t pg_f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
PG_HD float pg_u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
PG_HD int pg_bfind(uint32_t v) {   // index of the highest set bit, v != 0
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}
PG_HD int pg_popc(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
PG_HD float pg_rcp(float x) {      // IEEE-rounded reciprocal: the same value as 1.0f / x on the host, without the division routine
#ifdef __CUDA_ARCH__
    return __frcp_rn(x);
#else
    return 1.0f / x;
#endif
}
PG_HD float4 pg_ldg4(const float4* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
#if defined(__CUDACC__)
__device__ __forceinline__ float4 pg_lds4(uint32_t saddr) {      // 16-byte load from a shared-memory address (cvta'd)
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
#endif
// Embree 3 vec3 forms on FMA hardware (restated): dot = madd(x,x,madd(y,y,z*z)), cross = msub(..)
PG_HD float e_dot(V3 a, V3 b) { return pg_fma(a.x, b.x, pg_fma(a.y, b.y, a.z * b.z)); }
PG_HD V3 e_cross(V3 a, V3 b) {
    return v3(pg_fma(a.y, b.z, -(a.z * b.y)), pg_fma(a.z, b.x, -(a.x * b.z)), pg_fma(a.x, b.y, -(a.y * b.x)));
}
#ifdef __CUDACC__
__device__ __forceinline__ float f_expf(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float f_powf(float x, float y) { return (float)pow((double)x, (double)y); }
// the same for a small whole exponent (MTL "Ns 32"): square-and-multiply in double is within 3 ulp of a DOUBLE of pow(),
// which never survives the rounding to float; saves the ~200-instruction double pow per lit Phong hit
__device__ __forceinline__ float f_powf_whole(float x, float y) {
    const int n = (int)y;
    if ((float)n != y || n < 1 || n > 256) return f_powf(x, y);
    double b = (double)x, r = 1.0;
    for (int e = n; e; e >>= 1) { if (e & 1) r *= b; if (e > 1) b *= b; }
    return (float)r;
}
__device__ __forceinline__ float f_atan2f(float y, float x) { return (float)atan2((double)y, (double)x); }
__device__ __forceinline__ float f_asinf(float x) { return (float)asin((double)x); }
#endif

struct DevTexture { const uint8_t* data; int32_t width, height, pitch, bpp; };

struct DevCamera {            // PinHoleCamera.h:30-41
    int32_t width, height;
    float f_y;
    V3 from;
    M3 M;
};

// Per-frame constant block handed to the kernels by value.
struct DevScene {
    const float4* nodes;      // 8-wide nodes, root = node 0 (bvh8.cuh)
    const float4* tris;       // 3 x float4 per triangle in leaf order: (v0, id) (e1, -) (e2, -)
    const float4* shade;      // 4 x float4 per triangle in flat order: normals + uv + geomID
    const uint32_t* geom_first;     // geomID -> first flat triangle id
    const int32_t* geom_material;   // geomID -> material index
    const pgrt_material* materials;
    const DevTexture* textures;
    int32_t n_textures;
    DevTexture env;
    const pgrt_light* lights;
    int32_t n_lights;
    uint32_t n_tris;
    int32_t node_layout;      // PGRT_LAYOUT_Q8 (80 B nodes) or PGRT_LAYOUT_F32 (240 B nodes), bvh8.cuh
    int32_t loop_ww;          // traversal loop shape: 1 while-while, 0 if-if (traverse.cuh trav_advance)
    float bb_lo[3], bb_hi[3]; // bounds of the whole scene (the root of the binary tree): rays that miss them skip the traversal
};

#define CUDA_TRY(call)                                                                                  \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) return ctx->fail_cuda(e__, #call, __FILE__, __LINE__);                  \
    } while (0)
