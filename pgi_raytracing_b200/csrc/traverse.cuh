// traverse.cuh -- closest-hit query: replaces rtcIntersect1 (pg1/raytracer.cpp:130-148).
//
// Semantics (emb/doc/README.md:5963-5975, :6392-6396, :2482-2492 of the reference's vendored Embree docs):
// closest hit with tnear < t <= tfar, direction not normalised, no culling, no masks; (u,v) barycentric with
// p0 as base.  Ties in t resolve to the lowest flat triangle id (order independent).
#pragma once
#include "common.cuh"

struct HitRec { float t, u, v; uint32_t tri; };   // tri = flat triangle id, PGRT_INVALID_ID = miss

// Embree 3 TriangleM / Moeller-Trumbore formulation, restated (same operation order as the oracle states):
//   C = v0 - O; R = C x D; Ng = e2 x e1; den = Ng.D; U = R.e2; V = R.e1; T = Ng.C  (sign of den folded in)
PG_HD void tri_test(const float4* __restrict__ tris, uint32_t k, V3 O, V3 D, float tnear, float tfar, HitRec& best) {
    const float4 q0 = pg_ldg4(tris + 3 * (size_t)k), q1 = pg_ldg4(tris + 3 * (size_t)k + 1), q2 = pg_ldg4(tris + 3 * (size_t)k + 2);
    const V3 v0 = v3(q0.x, q0.y, q0.z), e1 = v3(q1.x, q1.y, q1.z), e2 = v3(q2.x, q2.y, q2.z);
    const V3 Ng = e_cross(e2, e1);
    const V3 C = v0 - O;
    const V3 R = e_cross(C, D);
    const float den = e_dot(Ng, D);
    const float absDen = fabsf(den);
    const uint32_t sgn = pg_f2u(den) & 0x80000000u;
    const float U = pg_u2f(pg_f2u(e_dot(R, e2)) ^ sgn);
    const float V = pg_u2f(pg_f2u(e_dot(R, e1)) ^ sgn);
    if (!(den != 0.0f && U >= 0.0f && V >= 0.0f && U + V <= absDen)) return;
    const float T = pg_u2f(pg_f2u(e_dot(Ng, C)) ^ sgn);
    if (!(absDen * tnear < T && T <= absDen * tfar)) return;
    const float rcp = 1.0f / absDen;
    const float t = T * rcp;
    const uint32_t id = pg_f2u(q0.w);
    if (t < best.t || (t == best.t && id < best.tri)) { best.t = t; best.u = U * rcp; best.v = V * rcp; best.tri = id; }
}

// geometric normal as Embree reports it (un-normalised), for pgrt_intersect's Ng fields
__device__ __forceinline__ V3 tri_ng(const float* __restrict__ pos, uint32_t id) {
    const float* p = pos + 9 * (size_t)id;
    const V3 v0 = v3(p[0], p[1], p[2]), v1 = v3(p[3], p[4], p[5]), v2 = v3(p[6], p[7], p[8]);
    return e_cross(v2 - v0, v0 - v1);
}

#define PGRT_STACK 128

// Binary-node traversal (layout in bvh_build.cuh, k_emit_bvh2).  Slab tests use FMA and a padded far bound:
// they only cull, the hit record comes from tri_test alone.
struct TravCount { uint32_t nodes, tris; };   // instrumented renders only (profile bit 1)

template <bool COUNT>
__device__ __forceinline__ HitRec trace_closest_t(const DevScene& sc, V3 O, V3 D, float tnear, float tfar, TravCount& tc) {
    HitRec best; best.t = tfar; best.u = 0.0f; best.v = 0.0f; best.tri = PGRT_INVALID_ID;
    if (sc.n_tris == 0) return best;
    const float ooeps = 8.271806e-25f;   // 2^-80
    const float idx = 1.0f / (fabsf(D.x) > ooeps ? D.x : copysignf(ooeps, D.x));
    const float idy = 1.0f / (fabsf(D.y) > ooeps ? D.y : copysignf(ooeps, D.y));
    const float idz = 1.0f / (fabsf(D.z) > ooeps ? D.z : copysignf(ooeps, D.z));
    const float oodx = O.x * idx, oody = O.y * idy, oodz = O.z * idz;
    int stack[PGRT_STACK];
    int sp = 0;
    int cur = (int)sc.root;
    const float4* __restrict__ nodes = sc.nodes;
    while (true) {
        if (cur >= 0) {
            if (COUNT) tc.nodes++;
            const float4 n0 = __ldg(nodes + 4 * (size_t)cur), n1 = __ldg(nodes + 4 * (size_t)cur + 1);
            const float4 n2 = __ldg(nodes + 4 * (size_t)cur + 2), n3 = __ldg(nodes + 4 * (size_t)cur + 3);
            const float far_pad = best.t * 1.0000005f;
            const float c0lox = __fmaf_rn(n0.x, idx, -oodx), c0hix = __fmaf_rn(n0.y, idx, -oodx);
            const float c0loy = __fmaf_rn(n0.z, idy, -oody), c0hiy = __fmaf_rn(n0.w, idy, -oody);
            const float c0loz = __fmaf_rn(n2.x, idz, -oodz), c0hiz = __fmaf_rn(n2.y, idz, -oodz);
            const float c1lox = __fmaf_rn(n1.x, idx, -oodx), c1hix = __fmaf_rn(n1.y, idx, -oodx);
            const float c1loy = __fmaf_rn(n1.z, idy, -oody), c1hiy = __fmaf_rn(n1.w, idy, -oody);
            const float c1loz = __fmaf_rn(n2.z, idz, -oodz), c1hiz = __fmaf_rn(n2.w, idz, -oodz);
            const float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), tnear));
            const float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), far_pad));
            const float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), tnear));
            const float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), far_pad));
            const bool h0 = c0min <= c0max * 1.0000005f, h1 = c1min <= c1max * 1.0000005f;
            const int r0 = __float_as_int(n3.x), r1 = __float_as_int(n3.y);
            if (h0 && h1) {
                const bool swp = c1min < c0min;
                cur = swp ? r1 : r0;
                stack[sp++] = swp ? r0 : r1;
            } else if (h0) cur = r0;
            else if (h1) cur = r1;
            else { if (sp == 0) break; cur = stack[--sp]; }
        } else {
            const uint32_t code = (uint32_t)~cur;
            const uint32_t first = code >> 2, cnt = (code & 3u) + 1u;
            if (COUNT) tc.tris += cnt;
            for (uint32_t k = 0; k < cnt; ++k) tri_test(sc.tris, first + k, O, D, tnear, tfar, best);
            if (sp == 0) break;
            cur = stack[--sp];
        }
    }
    return best;
}

__device__ __forceinline__ HitRec trace_closest(const DevScene& sc, V3 O, V3 D, float tnear, float tfar) {
    TravCount tc;
    return trace_closest_t<false>(sc, O, D, tnear, tfar, tc);
}
