// traverse.cuh -- closest-hit query: replaces rtcIntersect1 (pg1/raytracer.cpp:130-148).
//
// Semantics (emb/doc/README.md:5963-5975, :6392-6396, :2482-2492 of the reference's vendored Embree docs):
// closest hit with tnear < t <= tfar, direction not normalised, no culling, no masks; (u,v) barycentric with
// p0 as base.  Ties in t resolve to the lowest flat triangle id (order independent).
#pragma once
#include "common.cuh"
#include "bvh8.cuh"

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

// ---- wide-node traversal (layout and slot order in bvh8.cuh).  Box tests use FMA and padded bounds: they only cull,
// the hit record comes from tri_test alone, so the result does not depend on the tree (ties: lowest flat id).
struct TravCount { uint32_t nodes, tris; };   // instrumented renders only (profile bit 1)

PG_HD uint32_t slab4(uint32_t meta4, uint32_t nx4, uint32_t ny4, uint32_t nz4, uint32_t fx4, uint32_t fy4, uint32_t fz4,
                     float sx, float sy, float sz, float axn, float ayn, float azn, float axf, float ayf, float azf, float tnear, float far_pad,
                     uint32_t octinv4) {
    // four children at once: meta bytes -> (bit index, child bits); inner children are re-indexed by the ray octant
    const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t inner_mask4 = (is_inner4 >> 4) * 0xFFu;
    const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1F1F1F1Fu;
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
    uint32_t hitmask = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int sh = 8 * j;
        const float tx0 = pg_fma((float)((nx4 >> sh) & 0xFFu), sx, axn), tx1 = pg_fma((float)((fx4 >> sh) & 0xFFu), sx, axf);
        const float ty0 = pg_fma((float)((ny4 >> sh) & 0xFFu), sy, ayn), ty1 = pg_fma((float)((fy4 >> sh) & 0xFFu), sy, ayf);
        const float tz0 = pg_fma((float)((nz4 >> sh) & 0xFFu), sz, azn), tz1 = pg_fma((float)((fz4 >> sh) & 0xFFu), sz, azf);
        const float tmin = fmaxf(fmaxf(tx0, ty0), fmaxf(tz0, tnear));
        const float tmax = fminf(fminf(tx1, ty1), fminf(tz1, far_pad));
        if (tmin <= tmax * 1.0000005f) hitmask |= ((child_bits4 >> sh) & 0xFFu) << ((bit_index4 >> sh) & 0xFFu);
    }
    return hitmask;
}

// F32 layout: slab tests on float planes, four slots per call; a hit ORs the slot's pre-shifted word into the mask
PG_HD uint32_t slab4f(float4 w, float4 nx, float4 ny, float4 nz, float4 fx, float4 fy, float4 fz, float idx, float idy, float idz,
                      float nxo, float nyo, float nzo, float fxo, float fyo, float fzo, float tnear, float far_pad) {
    const float nxs[4] = {nx.x, nx.y, nx.z, nx.w}, nys[4] = {ny.x, ny.y, ny.z, ny.w}, nzs[4] = {nz.x, nz.y, nz.z, nz.w};
    const float fxs[4] = {fx.x, fx.y, fx.z, fx.w}, fys[4] = {fy.x, fy.y, fy.z, fy.w}, fzs[4] = {fz.x, fz.y, fz.z, fz.w};
    const uint32_t ws[4] = {pg_f2u(w.x), pg_f2u(w.y), pg_f2u(w.z), pg_f2u(w.w)};
    uint32_t hitmask = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float tx0 = pg_fma(nxs[j], idx, nxo), tx1 = pg_fma(fxs[j], idx, fxo);
        const float ty0 = pg_fma(nys[j], idy, nyo), ty1 = pg_fma(fys[j], idy, fyo);
        const float tz0 = pg_fma(nzs[j], idz, nzo), tz1 = pg_fma(fzs[j], idz, fzo);
        const float tmin = fmaxf(fmaxf(tx0, ty0), fmaxf(tz0, tnear));
        const float tmax = fminf(fminf(tx1, ty1), fminf(tz1, far_pad));
        if (tmin <= tmax * 1.0000005f) hitmask |= ws[j];
    }
    return hitmask;
}

// ---- per-ray constants of the two layouts, the visit of one node, and a resumable traversal built from them.
// trav_advance runs a bounded number of while-while rounds, so the same code serves the run-to-completion queries
// (trace_closest8 / trace_closest8f: shadow rays, pgrt_intersect, k_secondary, the CPU emulation) and the persistent
// k_trace, which swaps finished rays for new ones between rounds.
struct RayCtxF {                       // PGRT_LAYOUT_F32
    V3 O, D; float tnear, tfar;
    bool ww;                           // loop shape, see trav_advance
    float idx, idy, idz;
    float nxo, nyo, nzo, fxo, fyo, fzo; // -O * idir per axis, lowered (near planes) / raised (far planes) by that axis' rounding pad
    float pad_far;                      // absolute slack of the comparison with the best hit so far (see ray_far_pad)
    int onx, ony, onz, ofx, ofy, ofz;  // float4 offsets of the near / far planes inside a node, by ray sign
    uint32_t octinv, sw1, sw2, sw4;    // delta-swap masks that move internal hit bits from 24 + s to 24 + (s ^ octinv)
#ifdef PGRT_SMEM_TOP
    uint32_t top;                      // builds with -DPGRT_SMEM_TOP=n (measurement): shared-memory address of a copy of nodes 0 .. n-1, 0 = none
#endif
};
struct RayCtxQ {                       // PGRT_LAYOUT_Q8
    V3 O, D; float tnear, tfar;
    bool ww;
    float idx, idy, idz;
    float pox, poy, poz;               // 2^-22 |O * idir| per axis: the rounding of O seen through that axis' reciprocal
    float pad_far;                     // absolute slack of the comparison with the best hit so far (see ray_far_pad)
    bool negx, negy, negz;
    uint32_t octinv, octinv4;
};

// The best hit's t comes from tri_test, whose own rounding starts with O - v0: an error of 2^-24 max|O| in position, i.e. about
// that over |D| in t.  A box that the hit lies on the border of (a triangle hit at the corner of its own box; the lower-id
// twin of a duplicated triangle) must still pass "entry <= best t", so that comparison gets an absolute slack of
// 2^-21 max|O_i| * min|1/D_i| besides the relative one -- small for every ray, unlike the per-axis slab pads, which grow without
// bound for a ray parallel to an axis and would switch off the culling by distance if they were used here.
#ifndef PGRT_FAR_REL
#define PGRT_FAR_REL 1.0000005f
#endif
PG_HD float ray_far_pad(V3 O, float idx, float idy, float idz) {
    return 4.7683716e-7f * fmaxf(fmaxf(fabsf(O.x), fabsf(O.y)), fabsf(O.z)) * fminf(fminf(fabsf(idx), fabsf(idy)), fabsf(idz));
}

PG_HD void ray_ctx_init(RayCtxF& r, V3 O, V3 D, float tnear, float tfar, bool ww) {
    r.O = O; r.D = D; r.tnear = tnear; r.tfar = tfar; r.ww = ww;
    const float ooeps = 8.271806e-25f;   // 2^-80
    r.idx = pg_rcp(fabsf(D.x) > ooeps ? D.x : copysignf(ooeps, D.x));
    r.idy = pg_rcp(fabsf(D.y) > ooeps ? D.y : copysignf(ooeps, D.y));
    r.idz = pg_rcp(fabsf(D.z) > ooeps ? D.z : copysignf(ooeps, D.z));
    // plane * idir - O * idir cancels when the origin is far from the box compared with t: the rounding of O * idir
    // (2^-24 of its magnitude) is an ABSOLUTE error on that axis' t.  Each axis carries its own pad, folded into the
    // constant of its FMA: a pad taken as the maximum over the axes lets a ray that is nearly parallel to one axis
    // (huge |O * idir| there) through every box that only the OTHER axes can reject (C5: 16 -> 38 ms per frame).
    const float oodx = O.x * r.idx, oody = O.y * r.idy, oodz = O.z * r.idz;
    const float px = 2.3841858e-7f * fabsf(oodx), py = 2.3841858e-7f * fabsf(oody), pz = 2.3841858e-7f * fabsf(oodz);   // 2^-22 |O * idir|
    r.nxo = -oodx - px; r.nyo = -oody - py; r.nzo = -oodz - pz;
    r.fxo = -oodx + px; r.fyo = -oody + py; r.fzo = -oodz + pz;
    r.pad_far = ray_far_pad(O, r.idx, r.idy, r.idz);
    const bool negx = r.idx < 0.0f, negy = r.idy < 0.0f, negz = r.idz < 0.0f;
    r.octinv = (negx ? 0u : 4u) | (negy ? 0u : 2u) | (negz ? 0u : 1u);
    // plane p of half h sits at float4 3 + 2*p + h
    r.onx = 3 + 2 * (negx ? 3 : 0); r.ony = 3 + 2 * (negy ? 4 : 1); r.onz = 3 + 2 * (negz ? 5 : 2);
    r.ofx = 3 + 2 * (negx ? 0 : 3); r.ofy = 3 + 2 * (negy ? 1 : 4); r.ofz = 3 + 2 * (negz ? 2 : 5);
    r.sw1 = (r.octinv & 1u) ? 0x55000000u : 0u; r.sw2 = (r.octinv & 2u) ? 0x33000000u : 0u; r.sw4 = (r.octinv & 4u) ? 0x0F000000u : 0u;
#ifdef PGRT_SMEM_TOP
    r.top = 0u;
#endif
}

PG_HD void ray_ctx_init(RayCtxQ& r, V3 O, V3 D, float tnear, float tfar, bool ww) {
    r.O = O; r.D = D; r.tnear = tnear; r.tfar = tfar; r.ww = ww;
    const float ooeps = 8.271806e-25f;
    r.idx = pg_rcp(fabsf(D.x) > ooeps ? D.x : copysignf(ooeps, D.x));
    r.idy = pg_rcp(fabsf(D.y) > ooeps ? D.y : copysignf(ooeps, D.y));
    r.idz = pg_rcp(fabsf(D.z) > ooeps ? D.z : copysignf(ooeps, D.z));
    // (origin - O) * idir carries the rounding of the subtraction, 2^-24 max(|origin|, |O|), times |idir|: an ABSOLUTE error on
    // that axis' t which a relative pad does not cover when the ray is nearly parallel to the axis (found by fuzzing
    // axis-aligned sheets hit on their border).  Per axis, never the maximum over the axes (see RayCtxF).
    r.pox = 2.3841858e-7f * fabsf(O.x * r.idx); r.poy = 2.3841858e-7f * fabsf(O.y * r.idy); r.poz = 2.3841858e-7f * fabsf(O.z * r.idz);
    r.pad_far = ray_far_pad(O, r.idx, r.idy, r.idz);
    r.negx = r.idx < 0.0f; r.negy = r.idy < 0.0f; r.negz = r.idz < 0.0f;
    r.octinv = (r.negx ? 0u : 4u) | (r.negy ? 0u : 2u) | (r.negz ? 0u : 1u);
    r.octinv4 = r.octinv * 0x01010101u;
}

// one node: returns the hit mask (bits 31..24 internal children in traversal priority order, bits 23..0 triangles)
PG_HD uint32_t node_visit(const float4* __restrict__ nodes, uint32_t ni, const RayCtxF& r, float best_t, uint32_t& child_base, uint32_t& tri_base,
                          uint32_t& imask) {
    const float4* __restrict__ nd = nodes + PGRT_NODE_F4_F32 * (size_t)ni;
    float4 f0, w0, w1, nx0, ny0, nz0, fx0, fy0, fz0, nx1, ny1, nz1, fx1, fy1, fz1;
#if defined(PGRT_SMEM_TOP) && defined(__CUDA_ARCH__)
    // the root and the nodes right behind it (the collapse numbers nodes level by level) from a copy in shared memory
    if (r.top != 0u && ni < (uint32_t)PGRT_SMEM_TOP) {
        const uint32_t a = r.top + 16u * (uint32_t)PGRT_NODE_F4_F32 * ni;
        f0 = pg_lds4(a); w0 = pg_lds4(a + 16u); w1 = pg_lds4(a + 32u);
        nx0 = pg_lds4(a + 16u * r.onx); ny0 = pg_lds4(a + 16u * r.ony); nz0 = pg_lds4(a + 16u * r.onz);
        fx0 = pg_lds4(a + 16u * r.ofx); fy0 = pg_lds4(a + 16u * r.ofy); fz0 = pg_lds4(a + 16u * r.ofz);
        nx1 = pg_lds4(a + 16u * r.onx + 16u); ny1 = pg_lds4(a + 16u * r.ony + 16u); nz1 = pg_lds4(a + 16u * r.onz + 16u);
        fx1 = pg_lds4(a + 16u * r.ofx + 16u); fy1 = pg_lds4(a + 16u * r.ofy + 16u); fz1 = pg_lds4(a + 16u * r.ofz + 16u);
    } else
#endif
    {
        f0 = pg_ldg4(nd); w0 = pg_ldg4(nd + 1); w1 = pg_ldg4(nd + 2);
        nx0 = pg_ldg4(nd + r.onx); ny0 = pg_ldg4(nd + r.ony); nz0 = pg_ldg4(nd + r.onz); fx0 = pg_ldg4(nd + r.ofx); fy0 = pg_ldg4(nd + r.ofy); fz0 = pg_ldg4(nd + r.ofz);
        nx1 = pg_ldg4(nd + r.onx + 1); ny1 = pg_ldg4(nd + r.ony + 1); nz1 = pg_ldg4(nd + r.onz + 1); fx1 = pg_ldg4(nd + r.ofx + 1); fy1 = pg_ldg4(nd + r.ofy + 1); fz1 = pg_ldg4(nd + r.ofz + 1);
    }
    const float far_pad = pg_fma(best_t, PGRT_FAR_REL, r.pad_far);
    uint32_t hitmask = slab4f(w0, nx0, ny0, nz0, fx0, fy0, fz0, r.idx, r.idy, r.idz, r.nxo, r.nyo, r.nzo, r.fxo, r.fyo, r.fzo, r.tnear, far_pad);
    hitmask |= slab4f(w1, nx1, ny1, nz1, fx1, fy1, fz1, r.idx, r.idy, r.idz, r.nxo, r.nyo, r.nzo, r.fxo, r.fyo, r.fzo, r.tnear, far_pad);
    uint32_t x;
    x = ((hitmask >> 1) ^ hitmask) & r.sw1; hitmask ^= x ^ (x << 1);
    x = ((hitmask >> 2) ^ hitmask) & r.sw2; hitmask ^= x ^ (x << 2);
    x = ((hitmask >> 4) ^ hitmask) & r.sw4; hitmask ^= x ^ (x << 4);
    child_base = pg_f2u(f0.x); tri_base = pg_f2u(f0.y); imask = pg_f2u(f0.z);
    return hitmask;
}

PG_HD uint32_t node_visit(const float4* __restrict__ nodes, uint32_t ni, const RayCtxQ& r, float best_t, uint32_t& child_base, uint32_t& tri_base,
                          uint32_t& imask) {
    const float4* __restrict__ nd = nodes + PGRT_NODE_F4_Q8 * (size_t)ni;
    const float4 n0 = pg_ldg4(nd), n1 = pg_ldg4(nd + 1), n2 = pg_ldg4(nd + 2), n3 = pg_ldg4(nd + 3), n4 = pg_ldg4(nd + 4);
    const uint32_t eb = pg_f2u(n0.w);
    const float sx = pg_u2f((eb & 0xFFu) << 23) * r.idx, sy = pg_u2f(((eb >> 8) & 0xFFu) << 23) * r.idy, sz = pg_u2f(((eb >> 16) & 0xFFu) << 23) * r.idz;
    const float ax = (n0.x - r.O.x) * r.idx, ay = (n0.y - r.O.y) * r.idy, az = (n0.z - r.O.z) * r.idz;
    // |origin * idir| <= |O * idir| + |a|: pad = 2^-22 of both
    const float pdx = pg_fma(2.3841858e-7f, fabsf(ax), r.pox), pdy = pg_fma(2.3841858e-7f, fabsf(ay), r.poy), pdz = pg_fma(2.3841858e-7f, fabsf(az), r.poz);
    const float axn = ax - pdx, ayn = ay - pdy, azn = az - pdz, axf = ax + pdx, ayf = ay + pdy, azf = az + pdz;
    const float far_pad = pg_fma(best_t, PGRT_FAR_REL, r.pad_far);
    const uint32_t lox0 = pg_f2u(n2.x), lox1 = pg_f2u(n2.y), loy0 = pg_f2u(n2.z), loy1 = pg_f2u(n2.w);
    const uint32_t loz0 = pg_f2u(n3.x), loz1 = pg_f2u(n3.y), hix0 = pg_f2u(n3.z), hix1 = pg_f2u(n3.w);
    const uint32_t hiy0 = pg_f2u(n4.x), hiy1 = pg_f2u(n4.y), hiz0 = pg_f2u(n4.z), hiz1 = pg_f2u(n4.w);
    uint32_t hitmask = slab4(pg_f2u(n1.z), r.negx ? hix0 : lox0, r.negy ? hiy0 : loy0, r.negz ? hiz0 : loz0, r.negx ? lox0 : hix0, r.negy ? loy0 : hiy0,
                             r.negz ? loz0 : hiz0, sx, sy, sz, axn, ayn, azn, axf, ayf, azf, r.tnear, far_pad, r.octinv4);
    hitmask |= slab4(pg_f2u(n1.w), r.negx ? hix1 : lox1, r.negy ? hiy1 : loy1, r.negz ? hiz1 : loz1, r.negx ? lox1 : hix1, r.negy ? loy1 : hiy1,
                     r.negz ? loz1 : hiz1, sx, sy, sz, axn, ayn, azn, axf, ayf, azf, r.tnear, far_pad, r.octinv4);
    child_base = pg_f2u(n1.x); tri_base = pg_f2u(n1.y); imask = eb >> 24;
    return hitmask;
}

struct TravState {
    uint2 ng;        // node group: x = first child node, y = hit bits (31..24) | imask (7..0)
    uint2 tg;        // triangle group: x = first triangle, y = hit bits (23..0)
    int sp;
    HitRec best;
};

PG_HD void trav_init(TravState& s, float tfar) {
    s.ng = make_uint2(0u, 0x80000000u); s.tg = make_uint2(0u, 0u); s.sp = 0;
    s.best.t = tfar; s.best.u = 0.0f; s.best.v = 0.0f; s.best.tri = PGRT_INVALID_ID;
}

// Up to `rounds` rounds; returns true once the ray is finished.  A round is: descend (while-while, r.ww: until the lane
// holds triangles to test or runs out of nodes; if-if: one node at most), then test the triangles it holds.
// While-while enters the triangle code less often per warp, but lanes that hold triangles idle while the others keep
// descending; measured with the persistent k_trace it loses on every workload (6 % on C2, 35 % on the 3 M-triangle
// grove, 50 % on the 10 M soup: profiles/r1_matrix_layout_loopshape.txt), so if-if is the default and PGRT_LOOP=ww the knob.
template <class RC, bool COUNT>
PG_HD bool trav_advance(const float4* __restrict__ nodes, const float4* __restrict__ tris, const RC& r, TravState& s, uint2* stack, TravCount& tc, int rounds) {
    for (int it = 0; it < rounds; ++it) {
        while (s.tg.y == 0u) {
            if (s.ng.y <= 0x00FFFFFFu) {
                if (s.sp == 0) return true;
                s.ng = stack[--s.sp];
            }
            const uint32_t hits = s.ng.y;
            const int bit = pg_bfind(hits);
            s.ng.y &= ~(1u << bit);
            if (s.ng.y > 0x00FFFFFFu) stack[s.sp++] = s.ng;
            const uint32_t slot = (uint32_t)(bit - 24) ^ r.octinv;
            const uint32_t rel = (uint32_t)pg_popc(hits & 0xFFu & ~(0xFFFFFFFFu << slot));
            uint32_t child_base, tri_base, imask;
            const uint32_t hitmask = node_visit(nodes, s.ng.x + rel, r, s.best.t, child_base, tri_base, imask);
            if (COUNT) tc.nodes++;
            s.ng.x = child_base;
            s.ng.y = (hitmask & 0xFF000000u) | imask;
            s.tg.x = tri_base;
            s.tg.y = hitmask & 0x00FFFFFFu;
            if (!r.ww) break;
        }
        while (s.tg.y) {
            const int bit = pg_bfind(s.tg.y);
            s.tg.y &= ~(1u << bit);
            if (COUNT) tc.tris++;
            tri_test(tris, s.tg.x + (uint32_t)bit, r.O, r.D, r.tnear, r.tfar, s.best);
        }
    }
    return false;
}

template <class RC, bool COUNT>
PG_HD HitRec trace_closest_rc(const float4* __restrict__ nodes, const float4* __restrict__ tris, uint32_t n_tris, V3 O, V3 D, float tnear, float tfar,
                              TravCount& tc, bool ww) {
    TravState s;
    trav_init(s, tfar);
    if (n_tris == 0) return s.best;
    RC r;
    ray_ctx_init(r, O, D, tnear, tfar, ww);
    uint2 stack[PGRT_STACK8];
    while (!trav_advance<RC, COUNT>(nodes, tris, r, s, stack, tc, 1 << 30)) {}
    return s.best;
}

template <bool COUNT>
PG_HD HitRec trace_closest8f(const float4* __restrict__ nodes, const float4* __restrict__ tris, uint32_t n_tris, V3 O, V3 D,
                             float tnear, float tfar, TravCount& tc, bool ww = true) {
    return trace_closest_rc<RayCtxF, COUNT>(nodes, tris, n_tris, O, D, tnear, tfar, tc, ww);
}

template <bool COUNT>
PG_HD HitRec trace_closest8(const float4* __restrict__ nodes, const float4* __restrict__ tris, uint32_t n_tris, V3 O, V3 D,
                            float tnear, float tfar, TravCount& tc, bool ww = false) {
    return trace_closest_rc<RayCtxQ, COUNT>(nodes, tris, n_tris, O, D, tnear, tfar, tc, ww);
}

// A ray that cannot reach the scene's bounding box needs no traversal: the shadow rays of the reference leave the light with
// the hit POSITION as their direction (LightSource.cpp:18-20) and nearly all of them point away from the scene, and so do
// camera rays that see only sky.  One slab test with generous absolute pads (2^-21 of every term that was rounded, four times
// what the node tests carry): like every box test of the traversal it only culls -- a hit is still decided by tri_test alone,
// and a ray that passes here walks the tree as before.
PG_HD bool ray_misses_box(const float* lo, const float* hi, V3 O, V3 D, float tnear, float tfar) {
    const float ooeps = 8.271806e-25f;   // 2^-80, as in ray_ctx_init
    const float o[3] = {O.x, O.y, O.z}, d[3] = {D.x, D.y, D.z};
    float t0 = tnear, t1 = tfar;
    for (int a = 0; a < 3; ++a) {
        const float id = pg_rcp(fabsf(d[a]) > ooeps ? d[a] : copysignf(ooeps, d[a]));
        const float oi = o[a] * id, p0 = lo[a] * id, p1 = hi[a] * id;
        const float pad = 4.7683716e-7f * (fabsf(oi) + fabsf(p0) + fabsf(p1));
        const float a0 = p0 - oi, a1 = p1 - oi;
        t0 = fmaxf(t0, fminf(a0, a1) - pad); t1 = fminf(t1, fmaxf(a0, a1) + pad);
    }
    return !(t0 <= t1);      // (NaN anywhere: not a provable miss)
}

#ifdef __CUDACC__
template <bool COUNT>
__device__ __forceinline__ HitRec trace_closest_t(const DevScene& sc, V3 O, V3 D, float tnear, float tfar, TravCount& tc) {
    if (sc.node_layout == PGRT_LAYOUT_F32) return trace_closest8f<COUNT>(sc.nodes, sc.tris, sc.n_tris, O, D, tnear, tfar, tc, sc.loop_ww != 0);
    return trace_closest8<COUNT>(sc.nodes, sc.tris, sc.n_tris, O, D, tnear, tfar, tc, sc.loop_ww != 0);
}

__device__ __forceinline__ HitRec trace_closest(const DevScene& sc, V3 O, V3 D, float tnear, float tfar) {
    TravCount tc;
    return trace_closest_t<false>(sc, O, D, tnear, tfar, tc);
}
#endif  // __CUDACC__
