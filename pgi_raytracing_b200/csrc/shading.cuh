// shading.cuh -- device restatement of the leaf functions of the reference's shading path.
// Citations are relative to src/pg/pg1_embree/ of the reference.  Bug-compatible on purpose
// (SURVEY.md Appendix A): channel swaps, the mis-scaled sRGB encode, black texels on integer
// coordinates, shadow rays that use the hit position as a direction, ... are part of the contract.
#pragma once
#include "common.cuh"

struct Col4 { float r, g, b, a; };   // structs.h:11-14
struct Col3 { float r, g, b; };      // structs.h:16

__device__ __forceinline__ float byte_over_255(uint8_t byte) {
    const float b = (float)byte, r = 0.003921568859368563f;   // (float)(1.0f / 255.0f)
    const float q = b * r;
    return __fmaf_rn(__fmaf_rn(-q, 255.0f, b), r, q);
}

// ---- Texture::get_pixel (texture.cpp:56-62): bytes are B,G,R(,A); the returned .r carries the BLUE byte.
// The reference does not bounds-check; out-of-range texel indices are defined here by clamping (as the oracle does).
__device__ __forceinline__ Col3 tex_get_pixel(const DevTexture& t, int x, int y) {
    x = min(max(x, 0), t.width - 1); y = min(max(y, 0), t.height - 1);
    const uint8_t* p = t.data + (size_t)y * t.pitch + (size_t)x * t.bpp;
    // byte / 255.0f, correctly rounded, without the IEEE division routine: q = b * (1/255) is off by at most one ulp and
    // one residual step repairs it -- exact for all 256 inputs (tests/test_host_logic.py checks the identity)
    Col3 c;
    c.r = byte_over_255(__ldg(p)); c.g = byte_over_255(__ldg(p + 1)); c.b = byte_over_255(__ldg(p + 2));
    return c;
}

// ---- Texture::get_texel (texture.cpp:77-130)
__device__ __forceinline__ Col3 tex_get_texel(const DevTexture& t, const float u, const float v) {
    const float x = u * (float)t.width;
    const float y = v * (float)t.height;
    const int tmp_x1 = (int)floorf(x), tmp_x2 = (int)ceilf(x);
    const int tmp_y1 = (int)floorf(y), tmp_y2 = (int)ceilf(y);
    const int x1 = (tmp_x1 < 1) ? tmp_x1 + 1 : tmp_x1;               // :96
    const int x2 = (tmp_x2 == t.width) ? 0 : tmp_x2;                 // :97
    const int y1 = (tmp_y1 < 1) ? tmp_y1 + 1 : tmp_y1;               // :98
    const int y2 = (tmp_y2 == t.height) ? tmp_y1 : tmp_y2;           // :99
    Col3 out; out.r = 0.0f; out.g = 0.0f; out.b = 0.0f;
    if (x1 == x2 || y1 == y2) return out;                            // :101-103
    const Col3 x1y1 = tex_get_pixel(t, x1, y1), x2y1 = tex_get_pixel(t, x2, y1);
    const Col3 x1y2 = tex_get_pixel(t, x1, y2), x2y2 = tex_get_pixel(t, x2, y2);
    const float Q11 = ((float)x2 - x) / (float)(x2 - x1), Q21 = (x - (float)x1) / (float)(x2 - x1);   // :110-113
    const float f1r = x1y1.r * Q11 + x2y1.r * Q21, f1g = x1y1.g * Q11 + x2y1.g * Q21, f1b = x1y1.b * Q11 + x2y1.b * Q21;
    const float f2r = x1y2.r * Q11 + x2y2.r * Q21, f2g = x1y2.g * Q11 + x2y2.g * Q21, f2b = x1y2.b * Q11 + x2y2.b * Q21;
    const float wy1 = ((float)y2 - y) / (float)(y2 - y1), wy2 = (y - (float)y1) / (float)(y2 - y1);   // :125-127
    out.r = f1r * wy1 + f2r * wy2; out.g = f1g * wy1 + f2g * wy2; out.b = f1b * wy1 + f2b * wy2;
    return out;
}

// ---- SphericalMap::get_texel (SphericalMap.cpp:17-29)
__device__ __forceinline__ Col4 env_get_texel(const DevTexture& env, const float x, const float y, const float z) {
    V3 vec = normalize3(v3(x, z, y));
    const float u = (float)(0.5 + (double)f_atan2f(vec.x, vec.z) / (2 * 3.14159265358979323846));
    const float v = (float)(0.5 - (double)f_asinf(vec.y) / 3.14159265358979323846);
    Col4 o; o.r = 0.0f; o.g = 0.0f; o.b = 0.0f; o.a = 1.0f;
    if (env.data == nullptr) return o;
    const Col3 c = tex_get_texel(env, u, v);
    o.r = c.b; o.g = c.g; o.b = c.r;
    return o;
}

// ---- utils.cpp:204-241
__device__ __forceinline__ float srgb_compress(float u) {
    if (u <= 0) return 0.0f;
    if (u >= 1) return 1.0f;
    if ((double)u <= 0.00313080) return (float)(12.92 * (double)u);
    return (float)(1.00 * pow((double)u, 1 / 2.4) - 0.055);   // sic: not 1.055
}
__device__ __forceinline__ float srgb_expand(float u) {
    if (u <= 0) return 0.0f;
    if (u >= 1) return 1.0f;
    if ((double)u <= 0.04045) return (float)((double)u / 12.92);
    return (float)pow(((double)u + 0.055) / 1.055, 2.4);
}
__device__ __forceinline__ Col4 mix_srgb(Col4 c0, Col4 c1, float alpha) {
    // expand / mix_linear / compress each build {f(in.b), f(in.g), f(in.r)}: three swaps = one net R<->B swap
    Col4 e0, e1, m, o;
    e0.r = srgb_expand(c0.b); e0.g = srgb_expand(c0.g); e0.b = srgb_expand(c0.r);
    e1.r = srgb_expand(c1.b); e1.g = srgb_expand(c1.g); e1.b = srgb_expand(c1.r);
    m.r = (alpha * e0.b + (1 - alpha) * e1.b); m.g = (alpha * e0.g + (1 - alpha) * e1.g); m.b = (alpha * e0.r + (1 - alpha) * e1.r);
    o.r = srgb_compress(m.b); o.g = srgb_compress(m.g); o.b = srgb_compress(m.r); o.a = 0.1f;
    return o;
}

// ---- Raytracer::gamma (raytracer.cpp:439-446)
__device__ __forceinline__ Col4 gamma_correct(Col4 in, float gamma_level) {
    if (gamma_level == 0.5f) {
        // the slider default (raytracer.cpp:450): pow(c, 0.5) correctly rounded IS sqrt(c) correctly rounded (NaN for c < 0
        // on both; -0 squares to +0 either way), and sqrtf is IEEE here (-prec-sqrt=true); six double pow per pixel saved
        const float sb = sqrtf(in.b), sg = sqrtf(in.g), sr = sqrtf(in.r);
        Col4 o; o.r = sb * sb; o.g = sg * sg; o.b = sr * sr; o.a = 1.0f;
        return o;
    }
    const float b = f_powf(in.b, gamma_level) * f_powf(in.b, gamma_level);
    const float g = f_powf(in.g, gamma_level) * f_powf(in.g, gamma_level);
    const float r = f_powf(in.r, gamma_level) * f_powf(in.r, gamma_level);
    Col4 o; o.r = b; o.g = g; o.b = r; o.a = 1.0f;
    return o;
}

// ---- secondary-ray makers (raytracer.cpp:178-235); both re-normalise their (already unit) inputs
struct RayRec { V3 o; float tnear; V3 d; float time; };

__device__ __forceinline__ RayRec make_reflection_ray(V3 direction, V3 normal, V3 hit_point, float ior) {
    direction = normalize3(direction); normal = normalize3(normal);
    RayRec r; r.d = direction - 2 * (dot3(direction, normal) * normal);
    r.o = hit_point; r.tnear = 0.01f; r.time = ior;
    return r;
}
__device__ __forceinline__ RayRec make_refraction_ray(V3 direction, V3 normal, float n1, float n2, V3 hit_point) {
    direction = normalize3(direction); normal = normalize3(normal);
    const float n1_n2 = n1 / n2;
    const float d_n_ = dot3(direction, normal);
    const V3 scaled = v3(direction.x * n1_n2, direction.y * n1_n2, direction.z * n1_n2);
    const float k = n1_n2 * d_n_ + sqrtf(1 - ((n1_n2 * n1_n2) * (1 - (d_n_ * d_n_))));   // NaN <=> total internal reflection
    RayRec r; r.d = scaled - k * normal;
    r.o = hit_point; r.tnear = 0.01f; r.time = n2;
    return r;
}

// ---- counter-based RNG shared (as a spec) with the oracle; replaces raytracer.cpp:407 / PinHoleCamera.cpp:77
__device__ __forceinline__ uint32_t mix32(uint32_t h) { h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16; return h; }
__device__ __forceinline__ float rng_u01(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t dim) {
    uint32_t h = mix32(seed + 0x9E3779B9u * (pixel + 1u));
    h = mix32(h ^ (sample * 4u + dim + 0x85EBCA6Bu));
    return (float)(h >> 8) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float rng_uniform(float a, float b, float u) { return (b - a) * u + a; }

// ---- diffuse bounce of the path-tracing mode (shader_mode 3; README.md:21 "To do: path tracing", no reference code).
// Cosine-weighted direction = normalize(n + q), q uniform on the unit sphere, found by rejection from the cube: only
// +, *, /, sqrt, so the oracle reproduces it bit for bit.  The random numbers are keyed by the BITS of the hit point, the
// level and the frame seed: a ray carries no sample id below level 0, and the hit point is deterministic.
__device__ __forceinline__ V3 path_bounce_direction(V3 n, V3 hitp, uint32_t seed, int level) {
    uint32_t k = mix32(seed ^ (0x9E3779B9u * (uint32_t)(level + 1)));
    k = mix32(k ^ __float_as_uint(hitp.x)); k = mix32(k ^ __float_as_uint(hitp.y)); k = mix32(k ^ __float_as_uint(hitp.z));
    V3 q = n;
    for (uint32_t t = 0; t < 16u; ++t) {
        const float x = 2.0f * ((float)(mix32(k + 0x85EBCA6Bu * (3u * t + 1u)) >> 8) * (1.0f / 16777216.0f)) - 1.0f;
        const float y = 2.0f * ((float)(mix32(k + 0x85EBCA6Bu * (3u * t + 2u)) >> 8) * (1.0f / 16777216.0f)) - 1.0f;
        const float z = 2.0f * ((float)(mix32(k + 0x85EBCA6Bu * (3u * t + 3u)) >> 8) * (1.0f / 16777216.0f)) - 1.0f;
        const V3 c = v3(x, y, z);
        const float l2 = dot3(c, c);
        if (l2 <= 1.0f && l2 > 1e-6f) { q = normalize3(c); break; }
    }
    const V3 d = v3(n.x + q.x, n.y + q.y, n.z + q.z);
    return dot3(d, d) > 1e-8f ? normalize3(d) : n;
}

// ---- PinHoleCamera::generate_ray, both overloads (PinHoleCamera.cpp:31-63, :65-105)
__device__ __forceinline__ RayRec camera_ray_pinhole(const DevCamera& c, const float x_i, const float y_i) {
    V3 d = normalize3(v3(x_i - (float)(c.width / 2), (float)(c.height / 2) - y_i, -c.f_y));
    RayRec r; r.d = mul3(c.M, d); r.o = c.from; r.tnear = 0.001f; r.time = 0.0f;
    return r;
}
__device__ __forceinline__ RayRec camera_ray_lens(const DevCamera& c, const float x_i, const float y_i, const float focal_length,
                                                  const float rand1, const float rand2) {
    V3 d = normalize3(v3(x_i - (float)(c.width / 2), (float)(c.height / 2) - y_i, -c.f_y));
    V3 dws = normalize3(mul3(c.M, d));
    const V3 focal_point = v3(c.from.x + dws.x * focal_length, c.from.y + dws.y * focal_length, c.from.z + dws.z * focal_length);
    const V3 shift = mul3(c.M, v3(rand1, rand2, 0.0f));
    RayRec r;
    r.o = v3(c.from.x + shift.x, c.from.y + shift.y, c.from.z + shift.z);
    r.d = normalize3(v3(focal_point.x - r.o.x, focal_point.y - r.o.y, focal_point.z - r.o.z));
    r.tnear = 0.01f; r.time = 0.0f;
    return r;
}
