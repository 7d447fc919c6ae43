// =====================================================================================
// pg_oracle.cpp -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT.
//
// CPU restatement of the per-pixel render loop of rddrdhd/PGI_RayTracing
// (reference paths below are relative to /root/reference/src/pg/pg1_embree/).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library; the CUDA product never links, imports or calls it.
//
// PINNING STATUS
//   * pinned to the reference's OWN object code above the Embree boundary: oracle/ref.mk compiles the unmodified pg1/*.cpp
//     (raytracer, PinHoleCamera, LightSource, SphericalMap, texture, utils, vector3, matrix3x3, material, objloader, tutorials, ...)
//     into oracle/_ref/libpg_ref.so over stub Embree / FreeImage / window layers, and tests/test_oracle_vs_ref.py holds this file
//     to it BIT FOR BIT: Raytracer::trace at levels 0..7 on 12 000 rays per scene, whole un-jittered frames, is_illuminated,
//     both ray makers, both camera overloads, mix_srgb, gamma, Texture::get_texel, SphericalMap::get_texel, LoadOBJ / LoadMTL,
//     and the reference's own tutorial_1 / tutorial_2 print-outs; the shipped (clock-seeded) get_pixel as a converged mean.
//   * and to the known answers the reference holds for this path:
//       T1  tutorials.cpp:39-41,67-69,87-97  (one triangle, one ray: t=2,u=.05,v=.0667, normal (0,0,1), uv (0.050,0.933))
//       T2  tutorials.cpp:173-175 + data/test4.png (texel (r=1.000,g=0.000,b=0.500))
//       MTL data/6887_allied_avenger.mtl (5 materials, Ks parses to (1.0,0.8,0.8))
//   * PARITY UNPINNED only at the Embree boundary itself: the intersection arithmetic lives in Intel Embree 3.11.0
//     (embree3.lib, not vendored, emb/include/embree3/rtcore_config.h:6-10); no reference test pins rtcIntersect1 results, and
//     the stub of oracle/_ref calls THIS file's intersector.  Embree's published algorithm (Moeller-Trumbore as formulated in
//     its TriangleM intersector: C=v0-O, R=C x D, den=Ng.D, U=R.e2, V=R.e1, T=Ng.C, sign-folded, tnear < t <= tfar, no
//     culling) is restated here with IEEE division in place of Embree's rcp+Newton step.
//   * The one place the reference has no defined behaviour: Texture::get_pixel reads out of bounds for u >= 1 or outside
//     [0,1] (texture.cpp:56-62, tiled UVs); this file clamps the texel index, and so does the CUDA path.
//
// ARITHMETIC CONVENTIONS (the CUDA path states the same ones independently)
//   * FP32 everywhere the reference uses float; double only where the reference promotes
//     (utils.cpp:204-216, SphericalMap.cpp:22-23, raytracer.cpp:316).
//   * Compiled with -ffp-contract=off: no implicit FMA (MSVC /fp:precise x64 does not
//     contract).  fmaf() appears only where Embree's AVX2 path uses madd/msub
//     (vec3 dot/cross, rtcInterpolate).
//   * FTZ/DAZ on, as main() sets them (pg1_embree.cpp:8-9).
//   * Closest hit ties (equal t) resolve to the lowest flat triangle id so the result
//     does not depend on traversal order.
// =====================================================================================
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <vector>
#include <algorithm>
#include <atomic>
#if defined(__x86_64__)
#include <xmmintrin.h>
#include <pmmintrin.h>
#endif
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_INVALID_ID 0xFFFFFFFFu            // RTC_INVALID_GEOMETRY_ID, rtcore_common.h:45
#define ORC_IOR_AIR 1.000293f                  // material.h:15

extern "C" {

// POD mirrors of the product's C-ABI structs (include/pgrt.h); redefined here on purpose.
struct orc_material {
    float diffuse[3];    // Kd as parsed (x,y,z)
    float specular[3];   // Ks as parsed
    float shininess;     // Ns
    float ior;           // Ni
    int32_t type;        // MTL "shader N"
    int32_t diffuse_tex; // texture id or -1
};
struct orc_light {
    float position[3];
    float ambient[3];
    float diffuse[3];
    float specular[3];
};
struct orc_params {
    int32_t sampling_width;  // raytracer.cpp:398  (3)
    int32_t jitter;          // raytracer.cpp:408-410 on/off
    float focal_distance;    // raytracer.cpp:399  (200)
    float aperture;          // raytracer.cpp:400  (5)
    int32_t max_depth;       // raytracer.cpp:282  (7)
    float gamma_level;       // raytracer.cpp:450  (0.5)
    uint32_t seed;
    int32_t camera_mode;     // 0 thin lens PinHoleCamera.cpp:65-105, 1 pinhole :31-63
    int32_t shader_mode;     // 0 Whitted (trace as shipped), 1 Lambert (diffuse term only), 2 normal shader (:274-280),
                             // 3 path tracing (README.md:21 to-do; spec in include/pgrt.h and path_bounce_direction below)
    int32_t shadow_mode;     // 0 is_illuminated as shipped (:150-176); 1 hard shadows (README.md:20 to-do; spec in include/pgrt.h):
                             // hit point -> light, in units of that segment, t in (1e-3, 1]; the closest occluder decides
    int32_t reserved[6];
};
struct orc_stats {
    uint64_t rays_primary, rays_shadow, rays_reflection, rays_refraction;
};
// RTCRayHit-compatible, rtcore_ray.h:11-49
struct orc_rayhit {
    float org_x, org_y, org_z, tnear;
    float dir_x, dir_y, dir_z, time;
    float tfar; uint32_t mask, id, flags;
    float Ng_x, Ng_y, Ng_z, u, v;
    uint32_t primID, geomID, instID;
};
}

namespace {

// ------------------------------------------------------------------ vector3.cpp / matrix3x3.cpp
struct V3 { float x, y, z; };
static inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
static inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }   // vector3.cpp:104-107
static inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }   // :109-112
static inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }                       // :99-102
static inline V3 operator*(V3 a, float s) { return v3(s * a.x, s * a.y, s * a.z); }      // :114-117
static inline V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }      // :119-122
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }        // :56-59
static inline V3 cross(V3 a, V3 b) {                                                    // :38-44
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float sqr_norm(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }         // :19-22
static inline float l2norm(V3 a) { return sqrtf(sqr_norm(a)); }                         // :14-17
static inline V3 normalize(V3 a) {                                                      // :24-36
    const float n = sqr_norm(a);
    if (n != 0) { const float rn = 1 / sqrtf(n); a.x *= rn; a.y *= rn; a.z *= rn; }
    return a;
}
struct M3 { float m00, m01, m02, m10, m11, m12, m20, m21, m22; };
static inline M3 m3_from_basis(V3 bx, V3 by, V3 bz) {                                   // matrix3x3.cpp:32-45
    M3 m = {bx.x, by.x, bz.x, bx.y, by.y, bz.y, bx.z, by.z, bz.z}; return m;
}
static inline V3 operator*(const M3& a, V3 b) {                                         // matrix3x3.cpp:68-73
    return v3(a.m00 * b.x + a.m01 * b.y + a.m02 * b.z,
              a.m10 * b.x + a.m11 * b.y + a.m12 * b.z,
              a.m20 * b.x + a.m21 * b.y + a.m22 * b.z);
}

struct Color4 { float r, g, b, a; };   // structs.h:11-14
struct Color3 { float r, g, b; };      // structs.h:16

// Embree-style fused vec ops (restated; AVX2 path uses madd/msub) -- intersection + interpolation only.
static inline float e_dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
static inline V3 e_cross(V3 a, V3 b) {
    return v3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}

// ------------------------------------------------------------------ scene containers
struct Texture {                       // texture.h:27-32
    int width = 0, height = 0, scan_width = 0, pixel_size = 0;
    std::vector<uint8_t> data;         // top-down BGR(A), texture.cpp:46-47
};
struct Ray {                           // RTCRay subset, rtcore_ray.h:11-27
    float ox, oy, oz, tnear, dx, dy, dz, time, tfar;
};
struct Hit { float t, u, v; uint32_t tri; };   // tri = flat triangle id, ORC_INVALID_ID on miss

struct BNode { float lo[3], hi[3]; uint32_t left, right, first, count; };  // count>0 => leaf

struct Camera {                        // PinHoleCamera.h:30-41
    int width = 640, height = 480; float fov_y = 0.785f;
    V3 from = {0, 0, 0}, at = {0, 0, 0}; float f_y = 1.0f; M3 M = {1, 0, 0, 0, 1, 0, 0, 0, 1};
};

struct Scene {
    // per flat triangle (3 corners each), un-indexed as uploaded at raytracer.cpp:96-120
    std::vector<float> pos, nrm, uv;            // 9T, 9T, 6T
    std::vector<uint32_t> tri_geom, tri_prim;   // flat id -> (geomID, primID)
    std::vector<uint32_t> geom_first;           // geomID -> first flat id
    std::vector<int32_t> geom_material;         // geomID -> material index (rtcSetGeometryUserData, :83)
    std::vector<orc_material> materials;
    std::vector<Texture> textures;
    Texture env;
    std::vector<orc_light> lights;
    Camera cam;
    // acceleration structure (own binary SAH BVH; replaces Embree's, results are order independent)
    std::vector<BNode> nodes;
    std::vector<uint32_t> order;                // leaf triangle ids
    bool committed = false;
    uint32_t ntris() const { return (uint32_t)tri_geom.size(); }
};

struct FtzGuard {                      // pg1_embree.cpp:8-9
#if defined(__x86_64__)
    unsigned int saved;
    FtzGuard() { saved = _mm_getcsr(); _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON); _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON); }
    ~FtzGuard() { _mm_setcsr(saved); }
#endif
};

// ------------------------------------------------------------------ closest hit (rtcIntersect1 restated)
// Embree 3.11 TriangleM / MoellerTrumboreIntersector1 formulation (see header).  Returns true and
// fills (t,u,v) when the triangle is hit inside (tnear, tfar].
static inline bool tri_test(const float* p, const Ray& r, float& t, float& u, float& v, V3* ng = nullptr) {
    const V3 v0 = v3(p[0], p[1], p[2]), v1 = v3(p[3], p[4], p[5]), v2 = v3(p[6], p[7], p[8]);
    const V3 e1 = v0 - v1, e2 = v2 - v0;
    const V3 Ng = e_cross(e2, e1);
    const V3 O = v3(r.ox, r.oy, r.oz), D = v3(r.dx, r.dy, r.dz);
    const V3 C = v0 - O;
    const V3 R = e_cross(C, D);
    const float den = e_dot(Ng, D);
    const float absDen = fabsf(den);
    const bool neg = std::signbit(den);
    float U = e_dot(R, e2); if (neg) U = -U;
    float V = e_dot(R, e1); if (neg) V = -V;
    if (!(den != 0.0f && U >= 0.0f && V >= 0.0f && U + V <= absDen)) return false;
    float T = e_dot(Ng, C); if (neg) T = -T;
    if (!(absDen * r.tnear < T && T <= absDen * r.tfar)) return false;
    const float rcp = 1.0f / absDen;
    t = T * rcp; u = U * rcp; v = V * rcp;
    if (ng) *ng = Ng;
    return true;
}

static inline void consider(const Scene& s, uint32_t id, const Ray& r, Hit& best) {
    float t, u, v;
    if (tri_test(&s.pos[9 * (size_t)id], r, t, u, v)) {
        if (t < best.t || (t == best.t && id < best.tri)) { best.t = t; best.u = u; best.v = v; best.tri = id; }
    }
}

static Hit intersect_brute(const Scene& s, const Ray& r) {
    Hit best = {r.tfar, 0, 0, ORC_INVALID_ID};
    for (uint32_t i = 0; i < s.ntris(); ++i) consider(s, i, r, best);
    return best;
}

static inline bool box_test(const BNode& n, const Ray& r, const float inv[3], float tbest, float& tentry) {
    const float o[3] = {r.ox, r.oy, r.oz};
    // The box test only culls: it must never reject a triangle that tri_test accepts.  (plane - o) * inv carries the rounding
    // of the subtraction seen through inv -- an ABSOLUTE error on t of about 2^-24 max(|plane|, |o|) |inv| -- so every axis is
    // widened by 2^-22 (|o * inv| + |t|), and the best hit so far (whose t has tri_test's own rounding of o - v0) gets a
    // slack of 2^-21 max|o| min|inv|.  Found by fuzzing axis-aligned sheets hit on their border (tools/fuzz_emul.py scenes).
    const float omax = std::max(std::fabs(o[0]), std::max(std::fabs(o[1]), std::fabs(o[2])));
    const float imin = std::min(std::fabs(inv[0]), std::min(std::fabs(inv[1]), std::fabs(inv[2])));
    float t0 = r.tnear, t1 = tbest * 1.0000005f + 4.7683716e-7f * omax * imin;
    for (int a = 0; a < 3; ++a) {
        float ta = (n.lo[a] - o[a]) * inv[a], tb = (n.hi[a] - o[a]) * inv[a];
        if (ta > tb) std::swap(ta, tb);
        const float po = 2.3841858e-7f * std::fabs(o[a] * inv[a]);
        ta -= po + 2.3841858e-7f * std::fabs(ta); tb += po + 2.3841858e-7f * std::fabs(tb);
        // NaN (0*inf, inf-inf) compares false on both and leaves the interval untouched: conservative
        if (ta > t0) t0 = ta;
        if (tb < t1) t1 = tb;
    }
    tentry = t0;
    return t0 <= t1 * 1.0000005f + 1e-30f;
}

static Hit intersect_bvh(const Scene& s, const Ray& r) {
    Hit best = {r.tfar, 0, 0, ORC_INVALID_ID};
    if (s.nodes.empty()) return best;
    const float inv[3] = {1.0f / r.dx, 1.0f / r.dy, 1.0f / r.dz};
    uint32_t stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
        const BNode& n = s.nodes[stack[--sp]];
        float te;
        if (!box_test(n, r, inv, best.t, te)) continue;
        if (n.count) {
            for (uint32_t i = 0; i < n.count; ++i) consider(s, s.order[n.first + i], r, best);
        } else {
            float tl, tr;
            const bool hl = box_test(s.nodes[n.left], r, inv, best.t, tl);
            const bool hr = box_test(s.nodes[n.right], r, inv, best.t, tr);
            if (hl && hr) {
                if (tl <= tr) { stack[sp++] = n.right; stack[sp++] = n.left; }
                else { stack[sp++] = n.left; stack[sp++] = n.right; }
            } else if (hl) stack[sp++] = n.left;
            else if (hr) stack[sp++] = n.right;
        }
    }
    return best;
}

// Binned-SAH top-down build (own; stands in for rtcCommitScene's builder, raytracer.cpp:127).
struct BuildCtx {
    Scene& s; std::vector<float> clo, chi, cen;   // per-triangle bounds + centroid
    explicit BuildCtx(Scene& sc) : s(sc) {}
};
static void tri_bounds(const float* p, float lo[3], float hi[3]) {
    for (int a = 0; a < 3; ++a) {
        lo[a] = std::min(p[a], std::min(p[3 + a], p[6 + a]));
        hi[a] = std::max(p[a], std::max(p[3 + a], p[6 + a]));
    }
}
static uint32_t build_rec(BuildCtx& c, uint32_t first, uint32_t count, int depth) {
    Scene& s = c.s;
    const uint32_t idx = (uint32_t)s.nodes.size();
    s.nodes.push_back(BNode());
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (uint32_t i = 0; i < count; ++i) {
        const uint32_t t = s.order[first + i];
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], c.clo[3 * (size_t)t + a]); hi[a] = std::max(hi[a], c.chi[3 * (size_t)t + a]);
            clo[a] = std::min(clo[a], c.cen[3 * (size_t)t + a]); chi[a] = std::max(chi[a], c.cen[3 * (size_t)t + a]);
        }
    }
    BNode nd; memcpy(nd.lo, lo, 12); memcpy(nd.hi, hi, 12); nd.left = nd.right = 0; nd.first = first; nd.count = count;
    auto make_leaf = [&]() { s.nodes[idx] = nd; return idx; };
    if (count <= 4 || depth > 100) return make_leaf();
    const int NB = 16;
    int best_axis = -1, best_split = -1; double best_cost = 1e300;
    for (int a = 0; a < 3; ++a) {
        const float ext = chi[a] - clo[a];
        if (!(ext > 0)) continue;
        float blo[NB][3], bhi[NB][3]; uint32_t bcnt[NB];
        for (int b = 0; b < NB; ++b) { bcnt[b] = 0; for (int k = 0; k < 3; ++k) { blo[b][k] = FLT_MAX; bhi[b][k] = -FLT_MAX; } }
        const float scale = NB / ext;
        for (uint32_t i = 0; i < count; ++i) {
            const uint32_t t = s.order[first + i];
            int b = (int)((c.cen[3 * (size_t)t + a] - clo[a]) * scale); if (b >= NB) b = NB - 1; if (b < 0) b = 0;
            bcnt[b]++;
            for (int k = 0; k < 3; ++k) { blo[b][k] = std::min(blo[b][k], c.clo[3 * (size_t)t + k]); bhi[b][k] = std::max(bhi[b][k], c.chi[3 * (size_t)t + k]); }
        }
        double ra[NB]; uint32_t rc[NB];
        { float l[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, h[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}; uint32_t n = 0;
          for (int b = NB - 1; b > 0; --b) {
              for (int k = 0; k < 3; ++k) { l[k] = std::min(l[k], blo[b][k]); h[k] = std::max(h[k], bhi[b][k]); }
              n += bcnt[b]; rc[b] = n;
              const double dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2];
              ra[b] = n ? 2.0 * (dx * dy + dy * dz + dz * dx) : 0.0;
          } }
        { float l[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, h[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}; uint32_t n = 0;
          for (int b = 0; b < NB - 1; ++b) {
              for (int k = 0; k < 3; ++k) { l[k] = std::min(l[k], blo[b][k]); h[k] = std::max(h[k], bhi[b][k]); }
              n += bcnt[b];
              if (n == 0 || rc[b + 1] == 0) continue;
              const double dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2];
              const double cost = 2.0 * (dx * dy + dy * dz + dz * dx) * n + ra[b + 1] * rc[b + 1];
              if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = b; }
          } }
    }
    uint32_t mid;
    if (best_axis < 0) {
        mid = first + count / 2;   // all centroids coincide: median split by order
    } else {
        const float ext = chi[best_axis] - clo[best_axis]; const float scale = NB / ext;
        auto it = std::partition(s.order.begin() + first, s.order.begin() + first + count, [&](uint32_t t) {
            int b = (int)((c.cen[3 * (size_t)t + best_axis] - clo[best_axis]) * scale); if (b >= NB) b = NB - 1; if (b < 0) b = 0;
            return b <= best_split; });
        mid = (uint32_t)(it - s.order.begin());
        if (mid == first || mid == first + count) mid = first + count / 2;
    }
    const uint32_t l = build_rec(c, first, mid - first, depth + 1);
    const uint32_t r = build_rec(c, mid, first + count - mid, depth + 1);
    nd.count = 0; nd.left = l; nd.right = r; s.nodes[idx] = nd;
    return idx;
}
static void build_bvh(Scene& s) {
    s.nodes.clear(); s.order.resize(s.ntris());
    if (!s.ntris()) return;
    BuildCtx c(s); const size_t n = s.ntris();
    c.clo.resize(3 * n); c.chi.resize(3 * n); c.cen.resize(3 * n);
    for (size_t i = 0; i < n; ++i) {
        s.order[i] = (uint32_t)i;
        tri_bounds(&s.pos[9 * i], &c.clo[3 * i], &c.chi[3 * i]);
        for (int a = 0; a < 3; ++a) c.cen[3 * i + a] = 0.5f * (c.clo[3 * i + a] + c.chi[3 * i + a]);
    }
    s.nodes.reserve(2 * n);
    build_rec(c, 0, (uint32_t)n, 0);
}

// ------------------------------------------------------------------ texture.cpp
static Color3 tex_get_pixel(const Texture& t, int x, int y) {                           // texture.cpp:56-62
    // The reference does no bounds checking (reads out of bounds for u,v outside [0,1]);
    // both this oracle and the CUDA path define that case by clamping the indices.
    x = std::min(std::max(x, 0), t.width - 1); y = std::min(std::max(y, 0), t.height - 1);
    const int offset = y * t.scan_width + x * t.pixel_size;
    const float b = t.data[offset] / 255.0f;
    const float g = t.data[offset + 1] / 255.0f;
    const float r = t.data[offset + 2] / 255.0f;
    Color3 c = {b, g, r};   // .r carries the BLUE byte (sic)
    return c;
}
static Color3 tex_get_texel(const Texture& t, const float u, const float v) {           // texture.cpp:77-130
    const float x = u * float(t.width);
    const float y = v * float(t.height);
    const int tmp_x1 = (int)floorf(x), tmp_x2 = (int)ceilf(x);
    const int tmp_y1 = (int)floorf(y), tmp_y2 = (int)ceilf(y);
    const int x1 = (tmp_x1 < 1) ? tmp_x1 + 1 : tmp_x1;                                  // :96
    const int x2 = (tmp_x2 == t.width) ? 0 : tmp_x2;                                    // :97
    const int y1 = (tmp_y1 < 1) ? tmp_y1 + 1 : tmp_y1;                                  // :98
    const int y2 = (tmp_y2 == t.height) ? tmp_y1 : tmp_y2;                              // :99
    if (x1 == x2 || y1 == y2) { Color3 k = {0, 0, 0}; return k; }                       // :101-103
    const Color3 x1y1 = tex_get_pixel(t, x1, y1), x2y1 = tex_get_pixel(t, x2, y1);
    const Color3 x1y2 = tex_get_pixel(t, x1, y2), x2y2 = tex_get_pixel(t, x2, y2);
    const float Q11 = float((x2 - x) / (x2 - x1)), Q21 = float((x - x1) / (x2 - x1));   // :110-113
    const float Q12 = Q11, Q22 = Q21;
    const Color3 f1 = {x1y1.r * Q11 + x2y1.r * Q21, x1y1.g * Q11 + x2y1.g * Q21, x1y1.b * Q11 + x2y1.b * Q21};
    const Color3 f2 = {x1y2.r * Q12 + x2y2.r * Q22, x1y2.g * Q12 + x2y2.g * Q22, x1y2.b * Q12 + x2y2.b * Q22};
    const float wy1 = (y2 - y) / (y2 - y1), wy2 = (y - y1) / (y2 - y1);                 // :125-127
    Color3 out = {f1.r * wy1 + f2.r * wy2, f1.g * wy1 + f2.g * wy2, f1.b * wy1 + f2.b * wy2};
    return out;
}

// ------------------------------------------------------------------ SphericalMap.cpp:17-29
static Color4 env_get_texel(const Scene& s, const float x, const float y, const float z) {
    V3 vec = v3(x, z, y);
    vec = normalize(vec);
    // atan2/asin bind to the float overloads (float arguments); the division and sum promote to double
    const float u = (float)(0.5 + atan2f(vec.x, vec.z) / (2 * M_PI));
    const float v = (float)(0.5 - asinf(vec.y) / M_PI);
    if (s.env.data.empty()) { Color4 k = {0, 0, 0, 1}; return k; }
    const Color3 c = tex_get_texel(s.env, u, v);
    Color4 out = {c.b, c.g, c.r, 1};
    return out;
}

// ------------------------------------------------------------------ utils.cpp:204-241
static float _compress(float u) {
    if (u <= 0) return 0.0f;
    if (u >= 1) return 1.0f;
    if (u <= 0.00313080) return (float)(12.92 * u);
    return (float)(1.00 * pow(u, 1 / 2.4) - 0.055);
}
static float _expand(float u) {
    if (u <= 0) return 0.0f;
    if (u >= 1) return 1.0f;
    if (u <= 0.04045) return (float)(u / 12.92);
    return (float)pow((u + 0.055) / 1.055, 2.4);
}
static Color4 compress(Color4 c) { Color4 o = {_compress(c.b), _compress(c.g), _compress(c.r), 0.1f}; return o; }
static Color4 expand(Color4 c) { Color4 o = {_expand(c.b), _expand(c.g), _expand(c.r), 0.1f}; return o; }
static Color4 mix_linear(Color4 c0, Color4 c1, float alpha) {
    Color4 o = {(alpha * c0.b + (1 - alpha) * c1.b), (alpha * c0.g + (1 - alpha) * c1.g), (alpha * c0.r + (1 - alpha) * c1.r), 1.0f};
    return o;
}
static Color4 mix_srgb(Color4 c0, Color4 c1, float alpha) { return compress(mix_linear(expand(c0), expand(c1), alpha)); }

// ------------------------------------------------------------------ camera (PinHoleCamera.cpp)
static void camera_init(Camera& c, int w, int h, float fov_y, V3 from, V3 at) {         // :5-29
    c.width = w; c.height = h; c.fov_y = fov_y; c.from = from; c.at = at;
    c.f_y = h / (2 * tanf(fov_y / 2));
    const V3 up = v3(0.0f, 0.0f, 1.0f);                                                 // PinHoleCamera.h:36
    V3 z_c = from - at;
    V3 x_c = cross(up, z_c);
    V3 y_c = cross(z_c, x_c);
    z_c = normalize(z_c); x_c = normalize(x_c); y_c = normalize(y_c);
    c.M = m3_from_basis(x_c, y_c, z_c);
}
static Ray camera_ray_pinhole(const Camera& c, const float x_i, const float y_i) {      // :31-63
    V3 d = v3(x_i - (c.width / 2), (c.height / 2) - y_i, -c.f_y);
    d = normalize(d);
    const V3 dir = c.M * d;
    Ray r = {c.from.x, c.from.y, c.from.z, 0.001f, dir.x, dir.y, dir.z, 0.0f, FLT_MAX};
    return r;
}
static Ray camera_ray_lens(const Camera& c, const float x_i, const float y_i, const float focal_length,
                           const float rand1, const float rand2) {                      // :65-105
    V3 d = v3(x_i - c.width / 2, c.height / 2 - y_i, -c.f_y);
    d = normalize(d);
    V3 dws = c.M * d;
    dws = normalize(dws);
    const V3 focal_point = v3(c.from.x + dws.x * focal_length, c.from.y + dws.y * focal_length, c.from.z + dws.z * focal_length);
    const V3 shift = c.M * v3(rand1, rand2, 0);
    const V3 org = v3(c.from.x + shift.x, c.from.y + shift.y, c.from.z + shift.z);
    V3 nd = v3(focal_point.x - org.x, focal_point.y - org.y, focal_point.z - org.z);
    nd = normalize(nd);
    Ray r = {org.x, org.y, org.z, 0.01f, nd.x, nd.y, nd.z, 0.0f, FLT_MAX};
    return r;
}

// ------------------------------------------------------------------ counter-based RNG (replaces the two
// clock-seeded std::mt19937 of raytracer.cpp:407 / PinHoleCamera.cpp:77; spec shared with the CUDA path)
static inline uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16; return h;
}
static inline float rng_u01(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t dim) {
    uint32_t h = mix32(seed + 0x9E3779B9u * (pixel + 1u));
    h = mix32(h ^ (sample * 4u + dim + 0x85EBCA6Bu));
    return (float)(h >> 8) * (1.0f / 16777216.0f);
}
static inline float rng_uniform(float a, float b, float u) { return (b - a) * u + a; }  // uniform_real_distribution

// ------------------------------------------------------------------ tracer (raytracer.cpp:130-446)
struct Tracer {
    const Scene& s; const orc_params& p; bool brute;
    uint64_t n_primary = 0, n_shadow = 0, n_refl = 0, n_refr = 0;
    Tracer(const Scene& sc, const orc_params& pp, bool b) : s(sc), p(pp), brute(b) {}

    Hit get_ray_hit(const Ray& r) const { return brute ? intersect_brute(s, r) : intersect_bvh(s, r); }   // :130-148

    const orc_material& material_of(uint32_t tri) const { return s.materials[s.geom_material[s.tri_geom[tri]]]; }

    // rtcInterpolate0 (rtcore_geometry.h:288): w*a0 + u*a1 + v*a2 with w = 1-u-v, fused as Embree's madd chain
    V3 interp_normal(const Hit& h) const {
        const float* n = &s.nrm[9 * (size_t)h.tri]; const float w = 1.0f - h.u - h.v;
        return v3(fmaf(w, n[0], fmaf(h.u, n[3], h.v * n[6])), fmaf(w, n[1], fmaf(h.u, n[4], h.v * n[7])),
                  fmaf(w, n[2], fmaf(h.u, n[5], h.v * n[8])));
    }
    void interp_uv(const Hit& h, float& tu, float& tv) const {
        const float* c = &s.uv[6 * (size_t)h.tri]; const float w = 1.0f - h.u - h.v;
        tu = fmaf(w, c[0], fmaf(h.u, c[2], h.v * c[4])); tv = fmaf(w, c[1], fmaf(h.u, c[3], h.v * c[5]));
    }

    bool is_illuminated(const orc_light& light, V3 hit_position, V3 normal) {           // :150-176
        const V3 lp = v3(light.position[0], light.position[1], light.position[2]);
        if (p.shadow_mode == 1) {
            // README "To do: hard shadows" (README.md:20), non-default, no reference code: the ray the reference meant to cast
            const V3 to_light = v3(lp.x - hit_position.x, lp.y - hit_position.y, lp.z - hit_position.z);
            if (dot(normal, to_light) < 0) return false;
            Ray r = {hit_position.x, hit_position.y, hit_position.z, 1e-3f, to_light.x, to_light.y, to_light.z, 0.0f, 1.0f};
            n_shadow++;
            const Hit h = get_ray_hit(r);
            if (h.tri != ORC_INVALID_ID) return material_of(h.tri).type == 4;
            return true;
        }
        if (dot(normal, lp) < 0) return false;
        // LightSource::GenerateRay, LightSource.cpp:11-32: dir = the hit POSITION (sic), not hit - light
        Ray r = {lp.x, lp.y, lp.z, 0.01f, hit_position.x, hit_position.y, hit_position.z, 0.0f,
                 l2norm(v3(lp.x - hit_position.x, lp.y - hit_position.y, lp.z - hit_position.z))};
        n_shadow++;
        const Hit h = get_ray_hit(r);
        if (h.tri != ORC_INVALID_ID) return material_of(h.tri).type == 4;                // :166-172
        return true;
    }

    static Ray get_refraction_ray(V3 direction, V3 normal, float n1, float n2, V3 hit_point) {   // :178-207
        direction = normalize(direction); normal = normalize(normal);
        const float n1_n2 = (float)(n1 / n2);
        const float d_n_ = (float)(dot(direction, normal));
        const V3 scaled = v3(direction.x * n1_n2, direction.y * n1_n2, direction.z * n1_n2);
        const float k = n1_n2 * d_n_ + sqrtf(1 - ((n1_n2 * n1_n2) * (1 - (d_n_ * d_n_))));
        const V3 rd = scaled - k * normal;
        Ray r = {hit_point.x, hit_point.y, hit_point.z, 0.01f, rd.x, rd.y, rd.z, n2, FLT_MAX};
        return r;
    }
    static Ray get_reflection_ray(V3 direction, V3 normal, V3 hit_point, float ior) {   // :209-235
        direction = normalize(direction); normal = normalize(normal);
        const V3 rd = direction - 2 * (dot(direction, normal) * normal);
        Ray r = {hit_point.x, hit_point.y, hit_point.z, 0.01f, rd.x, rd.y, rd.z, ior, FLT_MAX};
        return r;
    }

    // Diffuse bounce of shader_mode 3: normalize(n + q), q uniform on the unit sphere by rejection from the cube; keyed by
    // the bits of the hit point, the level and the frame seed (the same spec as shading.cuh, restated).
    static V3 path_bounce_direction(V3 n, V3 hitp, uint32_t seed, int level) {
        auto bits = [](float f) { uint32_t u; memcpy(&u, &f, 4); return u; };
        uint32_t k = mix32(seed ^ (0x9E3779B9u * (uint32_t)(level + 1)));
        k = mix32(k ^ bits(hitp.x)); k = mix32(k ^ bits(hitp.y)); k = mix32(k ^ bits(hitp.z));
        V3 q = n;
        for (uint32_t t = 0; t < 16u; ++t) {
            const float x = 2.0f * ((float)(mix32(k + 0x85EBCA6Bu * (3u * t + 1u)) >> 8) * (1.0f / 16777216.0f)) - 1.0f;
            const float y = 2.0f * ((float)(mix32(k + 0x85EBCA6Bu * (3u * t + 2u)) >> 8) * (1.0f / 16777216.0f)) - 1.0f;
            const float z = 2.0f * ((float)(mix32(k + 0x85EBCA6Bu * (3u * t + 3u)) >> 8) * (1.0f / 16777216.0f)) - 1.0f;
            const V3 c = v3(x, y, z);
            const float l2 = dot(c, c);
            if (l2 <= 1.0f && l2 > 1e-6f) { q = normalize(c); break; }
        }
        const V3 d = v3(n.x + q.x, n.y + q.y, n.z + q.z);
        return dot(d, d) > 1e-8f ? normalize(d) : n;
    }

    Color4 trace(const Ray& ray, int level, Hit* first_hit = nullptr) {                 // :237-394
        const Hit h = get_ray_hit(ray);
        if (first_hit) *first_hit = h;
        V3 direction_vector = normalize(v3(ray.dx, ray.dy, ray.dz));
        if (h.tri != ORC_INVALID_ID) {
            const orc_material& material = material_of(h.tri);
            V3 normal_vector = normalize(interp_normal(h));
            const V3 hit_vector = v3(ray.ox, ray.oy, ray.oz) + v3(ray.dx, ray.dy, ray.dz) * h.t;   // :257-258
            float n1, n2;
            if (ray.time == ORC_IOR_AIR) { n1 = ORC_IOR_AIR; n2 = material.ior; }       // :261-267
            else { n1 = material.ior; n2 = ORC_IOR_AIR; }
            if (dot(normal_vector, direction_vector) > 0) normal_vector = -normal_vector;   // :269-272
            if (p.shader_mode == 2) {                                                   // :274-280 (commented normal shader)
                Color4 c = {((normal_vector.x) + 1) / 2, ((normal_vector.y) + 1) / 2, ((normal_vector.z) + 1) / 2, 1.0f};
                return c;
            }
            if (level >= p.max_depth) { Color4 k = {0, 0, 0, 1}; return k; }            // :282-283
            const V3 v = -direction_vector;
            if (material.type == 4 && (p.shader_mode == 0 || p.shader_mode == 3)) {     // :294-323
                const Ray reflection_ray = get_reflection_ray(direction_vector, normal_vector, hit_vector, n1);
                n_refl++;
                const Color4 reflection_color = trace(reflection_ray, level + 1);
                Color4 attenuation;
                attenuation.r = expf(-(1 - material.diffuse[0]) * h.t);
                attenuation.g = expf(-(1 - material.diffuse[1]) * h.t);
                attenuation.b = expf(-(1 - material.diffuse[2]) * h.t);
                const Ray refraction_ray = get_refraction_ray(direction_vector, normal_vector, n1, n2, hit_vector);
                if (refraction_ray.dx == refraction_ray.dx) {                           // :309 NaN <=> TIR
                    n_refr++;
                    const Color4 refraction_color = trace(refraction_ray, level + 1);
                    const float cos1 = fabsf(dot(normal_vector, v));
                    const float alpha = (n1 - n2) / (n1 + n2);
                    const float R = (float)((alpha * alpha + (1 - (alpha * alpha))) * pow((double)(1 - cos1), 5.0));   // :316
                    const Color4 c = mix_srgb(reflection_color, refraction_color, R);
                    Color4 o = {c.r * attenuation.r, c.g * attenuation.g, c.b * attenuation.b, 1};
                    return o;
                }
                Color4 o = {reflection_color.r * attenuation.r, reflection_color.g * attenuation.g, reflection_color.b * attenuation.b, 1};
                return o;
            }
            // default: Phong (:325-386)
            float blue = 0, green = 0, red = 0;
            float m_d_r, m_d_g, m_d_b;
            if (material.diffuse_tex < 0) {
                m_d_r = material.diffuse[2]; m_d_g = material.diffuse[1]; m_d_b = material.diffuse[0];   // :339-341
            } else {
                float tu, tv; interp_uv(h, tu, tv);
                const Color3 texel = tex_get_texel(s.textures[material.diffuse_tex], tu, 1.0f - tv);    // :345
                m_d_r = texel.r; m_d_g = texel.g; m_d_b = texel.b;
            }
            for (const orc_light& light : s.lights) {                                   // :351-384
                if (is_illuminated(light, hit_vector, normal_vector)) {
                    V3 light_vector = normalize(v3(light.position[0] - hit_vector.x, light.position[1] - hit_vector.y, light.position[2] - hit_vector.z));
                    V3 camera_vector = normalize(v3(ray.ox - ray.dx, ray.oy - ray.dy, ray.oz - ray.dz));   // :360 (sic)
                    const float gamma = material.shininess;
                    const float i_d_r = light.diffuse[2], i_d_g = light.diffuse[1], i_d_b = light.diffuse[0];
                    const float i_s_r = light.specular[2], i_s_g = light.specular[1], i_s_b = light.specular[0];
                    float m_s_r = material.specular[2], m_s_g = material.specular[1], m_s_b = material.specular[0];
                    const float ndl = dot(normal_vector, light_vector);
                    V3 l_r = normalize(2 * (ndl)*normal_vector - light_vector);
                    if (p.shader_mode == 1) {   // "Lambert": the diffuse addend only (README "Lambert shader")
                        blue += i_d_b * m_d_b * ndl; green += i_d_g * m_d_g * ndl; red += i_d_r * m_d_r * ndl;
                    } else {
                        const float spec = powf(dot(camera_vector, l_r), gamma);
                        blue += (i_d_b * m_d_b * ndl + i_s_b * m_s_b * spec);
                        green += (i_d_g * m_d_g * ndl + i_s_g * m_s_g * spec);
                        red += (i_d_r * m_d_r * ndl + i_s_r * m_s_r * spec);
                    }
                }
            }
            if (p.shader_mode == 3) {
                // path tracing (no reference counterpart): the Phong value above + albedo x one cosine-weighted bounce
                const V3 bd = path_bounce_direction(normal_vector, hit_vector, p.seed, level);
                const Ray bounce = {hit_vector.x, hit_vector.y, hit_vector.z, 0.01f, bd.x, bd.y, bd.z, ray.time, FLT_MAX};
                n_refl++;
                const Color4 c = trace(bounce, level + 1);
                Color4 o = {blue + m_d_b * c.r, green + m_d_g * c.g, red + m_d_r * c.b, 1.0f};
                return o;
            }
            Color4 o = {blue, green, red, 1.0f};                                        // :385
            return o;
        }
        return env_get_texel(s, direction_vector.x, direction_vector.y, direction_vector.z);   // :390-393
    }

    static Color4 gamma(Color4 in, float gamma_level) {                                 // :439-446
        const float b = powf(in.b, gamma_level) * powf(in.b, gamma_level);
        const float g = powf(in.g, gamma_level) * powf(in.g, gamma_level);
        const float r = powf(in.r, gamma_level) * powf(in.r, gamma_level);
        Color4 o = {b, g, r, 1.0f};
        return o;
    }

    Ray primary_ray(int xp, int yp, int sx, int sy) const {                             // :405-416
        const int w = p.sampling_width;
        const uint32_t pixel = (uint32_t)(yp * s.cam.width + xp), sample = (uint32_t)(sx * w + sy);
        float rand1 = 0.0f, rand2 = 0.0f;
        if (p.jitter) {
            rand1 = rng_uniform(-0.5f / w, 0.5f / w, rng_u01(p.seed, pixel, sample, 0));
            rand2 = rng_uniform(-0.5f / w, 0.5f / w, rng_u01(p.seed, pixel, sample, 1));
        }
        const float new_x = xp + (sx * (1.0f / w)) + rand1;
        const float new_y = yp + (sy * (1.0f / w)) + rand2;
        Ray r;
        if (p.camera_mode == 1) r = camera_ray_pinhole(s.cam, new_x, new_y);
        else {
            const float l1 = rng_uniform(-p.aperture / 2.0f, p.aperture / 2.0f, rng_u01(p.seed, pixel, sample, 2));
            const float l2 = rng_uniform(-p.aperture / 2.0f, p.aperture / 2.0f, rng_u01(p.seed, pixel, sample, 3));
            r = camera_ray_lens(s.cam, new_x, new_y, p.focal_distance, l1, l2);
        }
        r.time = ORC_IOR_AIR;                                                           // :416
        return r;
    }

    Color4 get_pixel(int xp, int yp, uint32_t* geom, uint32_t* prim) {                  // :396-437
        const int w = p.sampling_width;
        Color4 fin = {0.0f, 0.0f, 0.0f, 1.0f};
        std::vector<Color4> res((size_t)w * w);
        for (int x = 0; x < w; ++x)
            for (int y = 0; y < w; ++y) {
                const Ray pr = primary_ray(xp, yp, x, y);
                n_primary++;
                Hit fh;
                res[(size_t)x * w + y] = trace(pr, 0, &fh);
                if (x == 0 && y == 0 && geom) {
                    if (fh.tri == ORC_INVALID_ID) { *geom = ORC_INVALID_ID; *prim = ORC_INVALID_ID; }
                    else { *geom = s.tri_geom[fh.tri]; *prim = s.tri_prim[fh.tri]; }
                }
            }
        for (int x = 0; x < w; ++x)
            for (int y = 0; y < w; ++y) { fin.r += res[(size_t)x * w + y].r; fin.g += res[(size_t)x * w + y].g; fin.b += res[(size_t)x * w + y].b; }
        const int area = w * w;                                                         // (int)pow(sampling_width,2)
        Color4 in = {fin.b / area, fin.g / area, fin.r / area, 1.0f};
        return gamma(in, p.gamma_level);
    }
};

}  // namespace

// =====================================================================================
// C entry points (ctypes)
// =====================================================================================
extern "C" {

void* orc_create() { return new Scene(); }
void orc_destroy(void* h) { delete (Scene*)h; }

int orc_add_mesh(void* h, const float* pos, const float* nrm, const float* uv, uint32_t T, int32_t material_id, uint32_t* geom_id) {
    Scene& s = *(Scene*)h;
    const uint32_t g = (uint32_t)s.geom_first.size();
    s.geom_first.push_back(s.ntris());
    s.geom_material.push_back(material_id);
    s.pos.insert(s.pos.end(), pos, pos + 9 * (size_t)T);
    s.nrm.insert(s.nrm.end(), nrm, nrm + 9 * (size_t)T);
    s.uv.insert(s.uv.end(), uv, uv + 6 * (size_t)T);
    for (uint32_t i = 0; i < T; ++i) { s.tri_geom.push_back(g); s.tri_prim.push_back(i); }
    if (geom_id) *geom_id = g;
    s.committed = false;
    return 0;
}
int orc_set_materials(void* h, const orc_material* m, int n) { Scene& s = *(Scene*)h; s.materials.assign(m, m + n); return 0; }
static void fill_tex(Texture& t, const uint8_t* bytes, int w, int hgt, int pitch, int bpp) {
    t.width = w; t.height = hgt; t.scan_width = pitch; t.pixel_size = bpp; t.data.assign(bytes, bytes + (size_t)pitch * hgt);
}
int orc_set_texture(void* h, int id, const uint8_t* bytes, int w, int hgt, int pitch, int bpp) {
    Scene& s = *(Scene*)h; if (id < 0) return 1;
    if ((size_t)id >= s.textures.size()) s.textures.resize(id + 1);
    fill_tex(s.textures[id], bytes, w, hgt, pitch, bpp); return 0;
}
int orc_set_envmap(void* h, const uint8_t* bytes, int w, int hgt, int pitch, int bpp) { fill_tex(((Scene*)h)->env, bytes, w, hgt, pitch, bpp); return 0; }
int orc_set_lights(void* h, const orc_light* l, int n) { ((Scene*)h)->lights.assign(l, l + n); return 0; }
int orc_set_camera(void* h, int w, int hgt, float fov_y, const float* from, const float* at) {
    FtzGuard g; camera_init(((Scene*)h)->cam, w, hgt, fov_y, v3(from[0], from[1], from[2]), v3(at[0], at[1], at[2])); return 0;
}
int orc_commit(void* h) { Scene& s = *(Scene*)h; build_bvh(s); s.committed = true; return 0; }
uint32_t orc_num_triangles(void* h) { return ((Scene*)h)->ntris(); }
uint32_t orc_num_nodes(void* h) { return (uint32_t)((Scene*)h)->nodes.size(); }

// camera constants (f_y, 9 matrix entries row-major) for host-logic parity checks
void orc_get_camera(void* h, float* out10) {
    const Camera& c = ((Scene*)h)->cam; out10[0] = c.f_y; memcpy(out10 + 1, &c.M, 36);
}

// rtcIntersect1 over a batch (RTCRayHit layout).  brute=1 tests every triangle.
int orc_intersect(void* h, orc_rayhit* rh, uint64_t n, int brute, int threads) {
    Scene& s = *(Scene*)h;
    if (!s.committed && !brute) return 1;
    (void)threads;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 256)
#endif
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        FtzGuard g;
        orc_rayhit& q = rh[i];
        Ray r = {q.org_x, q.org_y, q.org_z, q.tnear, q.dir_x, q.dir_y, q.dir_z, q.time, q.tfar};
        const Hit hit = brute ? intersect_brute(s, r) : intersect_bvh(s, r);
        if (hit.tri != ORC_INVALID_ID) {   // miss leaves the struct untouched (emb/doc/README.md:6372-6378)
            float t, u, v; V3 ng = {0, 0, 0}; tri_test(&s.pos[9 * (size_t)hit.tri], r, t, u, v, &ng);
            q.tfar = hit.t; q.u = hit.u; q.v = hit.v; q.Ng_x = ng.x; q.Ng_y = ng.y; q.Ng_z = ng.z;
            q.geomID = s.tri_geom[hit.tri]; q.primID = s.tri_prim[hit.tri]; q.instID = ORC_INVALID_ID;
        }
    }
    return 0;
}

// One Producer iteration (simpleguidx11.cpp:95-118): rgba = W*H*4 floats, row 0 = top.
// geom/prim (optional) receive the primary hit of sample (0,0) of every pixel.
// threads: 1 = as shipped (simpleguidx11.cpp:104 pragma commented out); >1 = pragma restored over rows.
// Sub-rectangle [x0,x1) x [y0,y1) lets a bounded sample of a big frame be timed; pixels outside stay untouched.
int orc_render_region(void* h, const orc_params* p, int x0, int y0, int x1, int y1, float* rgba, uint32_t* geom, uint32_t* prim,
                      orc_stats* stats, int brute, int threads) {
    Scene& s = *(Scene*)h;
    if (!s.committed && !brute) return 1;
    if (p->sampling_width < 1) return 2;
    const int W = s.cam.width;
    uint64_t np = 0, ns = 0, nrl = 0, nrr = 0;
    (void)threads;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : np, ns, nrl, nrr)
#endif
    for (int y = y0; y < y1; ++y) {
        FtzGuard g;
        Tracer tr(s, *p, brute != 0);
        for (int x = x0; x < x1; ++x) {
            const size_t px = (size_t)y * W + x;
            const Color4 c = tr.get_pixel(x, y, geom ? geom + px : nullptr, prim ? prim + px : nullptr);
            const size_t offset = px * 4;                                               // simpleguidx11.cpp:108-114
            rgba[offset] = c.r; rgba[offset + 1] = c.g; rgba[offset + 2] = c.b; rgba[offset + 3] = c.a;
        }
        np += tr.n_primary; ns += tr.n_shadow; nrl += tr.n_refl; nrr += tr.n_refr;
    }
    if (stats) { stats->rays_primary = np; stats->rays_shadow = ns; stats->rays_reflection = nrl; stats->rays_refraction = nrr; }
    return 0;
}
int orc_render(void* h, const orc_params* p, float* rgba, uint32_t* geom, uint32_t* prim, orc_stats* stats, int brute, int threads) {
    Scene& s = *(Scene*)h;
    return orc_render_region(h, p, 0, 0, s.cam.width, s.cam.height, rgba, geom, prim, stats, brute, threads);
}

// Raytracer::trace (raytracer.cpp:237-394) on caller-supplied rays: in = 9 floats per ray (org, tnear, dir, time, tfar), out = Color4f
int orc_trace(void* h, const orc_params* p, const float* rays9, uint64_t n, int level, float* out4, int brute, int threads) {
    Scene& s = *(Scene*)h;
    if (!s.committed && !brute) return 1;
    (void)threads;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 64)
#endif
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        FtzGuard g;
        Tracer tr(s, *p, brute != 0);
        const float* q = rays9 + 9 * i;
        const Ray r = {q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8]};
        const Color4 c = tr.trace(r, level);
        out4[4 * i] = c.r; out4[4 * i + 1] = c.g; out4[4 * i + 2] = c.b; out4[4 * i + 3] = c.a;
    }
    return 0;
}
// Raytracer::is_illuminated (raytracer.cpp:150-176): light position, hit position, normal per query -> 0 / 1
int orc_is_illuminated(void* h, const orc_params* p, const float* light3, const float* hit3, const float* nrm3, uint64_t n, int32_t* out, int brute) {
    Scene& s = *(Scene*)h;
    if (!s.committed && !brute) return 1;
    FtzGuard g;
    Tracer tr(s, *p, brute != 0);
    for (uint64_t i = 0; i < n; ++i) {
        orc_light l = {};
        l.position[0] = light3[3 * i]; l.position[1] = light3[3 * i + 1]; l.position[2] = light3[3 * i + 2];
        out[i] = tr.is_illuminated(l, v3(hit3[3 * i], hit3[3 * i + 1], hit3[3 * i + 2]), v3(nrm3[3 * i], nrm3[3 * i + 1], nrm3[3 * i + 2])) ? 1 : 0;
    }
    return 0;
}

// ---- per-function entry points (unit parity of each App. A quirk)
void orc_mix_srgb(const float* c0, const float* c1, const float* alpha, uint64_t n, float* out) {
    FtzGuard g;
    for (uint64_t i = 0; i < n; ++i) {
        Color4 a = {c0[4 * i], c0[4 * i + 1], c0[4 * i + 2], c0[4 * i + 3]}, b = {c1[4 * i], c1[4 * i + 1], c1[4 * i + 2], c1[4 * i + 3]};
        const Color4 o = mix_srgb(a, b, alpha[i]);
        out[4 * i] = o.r; out[4 * i + 1] = o.g; out[4 * i + 2] = o.b; out[4 * i + 3] = o.a;
    }
}
// tex_id >= 0: material texture; -1: the env-map texture.  out = Color3 {r,g,b} as returned by Texture::get_texel
int orc_texture_get_texel(void* h, int tex_id, const float* uv, uint64_t n, float* out) {
    Scene& s = *(Scene*)h; FtzGuard g;
    const Texture* t = tex_id < 0 ? &s.env : ((size_t)tex_id < s.textures.size() ? &s.textures[tex_id] : nullptr);
    if (!t || t->data.empty()) return 1;
    for (uint64_t i = 0; i < n; ++i) { const Color3 c = tex_get_texel(*t, uv[2 * i], uv[2 * i + 1]); out[3 * i] = c.r; out[3 * i + 1] = c.g; out[3 * i + 2] = c.b; }
    return 0;
}
int orc_env_get_texel(void* h, const float* dirs, uint64_t n, float* out) {
    Scene& s = *(Scene*)h; FtzGuard g;
    for (uint64_t i = 0; i < n; ++i) { const Color4 c = env_get_texel(s, dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]); out[4 * i] = c.r; out[4 * i + 1] = c.g; out[4 * i + 2] = c.b; out[4 * i + 3] = c.a; }
    return 0;
}
void orc_gamma(const float* in, float gamma_level, uint64_t n, float* out) {
    FtzGuard g;
    for (uint64_t i = 0; i < n; ++i) {
        Color4 c = {in[4 * i], in[4 * i + 1], in[4 * i + 2], in[4 * i + 3]};
        const Color4 o = Tracer::gamma(c, gamma_level);
        out[4 * i] = o.r; out[4 * i + 1] = o.g; out[4 * i + 2] = o.b; out[4 * i + 3] = o.a;
    }
}
// primary rays of one frame in sample order ((y*W+x)*S + sx*w+sy): out = 9 floats per ray
// (org xyz, tnear, dir xyz, time, tfar)
int orc_primary_rays(void* h, const orc_params* p, float* out) {
    Scene& s = *(Scene*)h; FtzGuard g;
    Tracer tr(s, *p, true);
    const int w = p->sampling_width; size_t k = 0;
    for (int y = 0; y < s.cam.height; ++y)
        for (int x = 0; x < s.cam.width; ++x)
            for (int sx = 0; sx < w; ++sx)
                for (int sy = 0; sy < w; ++sy) {
                    const Ray r = tr.primary_ray(x, y, sx, sy);
                    float* o = out + 9 * k++;
                    o[0] = r.ox; o[1] = r.oy; o[2] = r.oz; o[3] = r.tnear; o[4] = r.dx; o[5] = r.dy; o[6] = r.dz; o[7] = r.time; o[8] = r.tfar;
                }
    return 0;
}
// secondary-ray makers: in = dir(3) normal(3) hit(3) n1 n2 per item; out = 9 floats per ray (as above)
void orc_secondary_rays(const float* in, uint64_t n, int refraction, float* out) {
    FtzGuard g;
    for (uint64_t i = 0; i < n; ++i) {
        const float* q = in + 11 * i;
        const V3 d = v3(q[0], q[1], q[2]), nn = v3(q[3], q[4], q[5]), hp = v3(q[6], q[7], q[8]);
        const Ray r = refraction ? Tracer::get_refraction_ray(d, nn, q[9], q[10], hp) : Tracer::get_reflection_ray(d, nn, hp, q[9]);
        float* o = out + 9 * i;
        o[0] = r.ox; o[1] = r.oy; o[2] = r.oz; o[3] = r.tnear; o[4] = r.dx; o[5] = r.dy; o[6] = r.dz; o[7] = r.time; o[8] = r.tfar;
    }
}
// rtcInterpolate0 (raytracer.cpp:252, :344): slot 0 -> normal (3 floats), slot 1 -> uv (2 floats)
int orc_interpolate(void* h, const uint32_t* geom, const uint32_t* prim, const float* u, const float* v, uint64_t n, int slot, float* out) {
    Scene& s = *(Scene*)h; FtzGuard g;
    orc_params p = {}; Tracer tr(s, p, true);
    for (uint64_t i = 0; i < n; ++i) {
        if (geom[i] >= s.geom_first.size()) return 1;
        Hit hit = {0.0f, u[i], v[i], s.geom_first[geom[i]] + prim[i]};
        if (hit.tri >= s.ntris()) return 1;
        if (slot == 0) { const V3 nn = tr.interp_normal(hit); out[3 * i] = nn.x; out[3 * i + 1] = nn.y; out[3 * i + 2] = nn.z; }
        else tr.interp_uv(hit, out[2 * i], out[2 * i + 1]);
    }
    return 0;
}
float orc_rng_u01(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t dim) { return rng_u01(seed, pixel, sample, dim); }
// n directions of the diffuse bounce (shader_mode 3) for normal nrm at hit points hit[3*i..]: out[3*i..]
void orc_path_bounce(const float* nrm, const float* hit, uint64_t n, uint32_t seed, int32_t level, float* out) {
    for (uint64_t i = 0; i < n; ++i) {
        const V3 d = Tracer::path_bounce_direction(v3(nrm[0], nrm[1], nrm[2]), v3(hit[3 * i], hit[3 * i + 1], hit[3 * i + 2]), seed, level);
        out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
    }
}
int orc_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
