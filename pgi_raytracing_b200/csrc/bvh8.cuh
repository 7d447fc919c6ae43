// bvh8.cuh -- the compact wide-node layout of the traversal: an 8-ary BVH with child boxes quantised to 8 bits
// against the node's own box (80 B per node, five 16-B loads), after Ylitie, Karras & Laine, "Efficient Incoherent
// Ray Traversal on GPUs Through Compressed Wide BVHs" (HPG 2017).  Written from the paper's description; nothing
// here exists in the reference, whose BVH lives inside the Embree binary (rtcCommitScene, pg1/raytracer.cpp:127).
//
// Everything in this file is a plain host/device function (PG_HD): the GPU build kernels call bvh8_gather /
// bvh8_emit once per wide node, and tests/emul/ compiles the same functions with g++ to check encoder + traversal
// against brute force on the CPU.
//
// Node = 5 x float4 (bit patterns):
//   n0 = (lo.x, lo.y, lo.z, ex | ey << 8 | ez << 16 | imask << 24)      e = biased exponent of the grid step 2^e
//   n1 = (child_base, tri_base, meta[0..3], meta[4..7])
//   n2 = (qlo.x[0..3], qlo.x[4..7], qlo.y[0..3], qlo.y[4..7])
//   n3 = (qlo.z[0..3], qlo.z[4..7], qhi.x[0..3], qhi.x[4..7])
//   n4 = (qhi.y[0..3], qhi.y[4..7], qhi.z[0..3], qhi.z[4..7])
// Child slot s holds the child lying furthest along (s&4 ? +x : -x, s&2 ? +y : -y, s&1 ? +z : -z), so visiting slots
// in the order s ^ (ray octant) is an approximate front-to-back order with no sorting.
//   meta[s] = 0                          empty slot
//           = 001 | 11sss                internal child: wide node child_base + (number of internal slots below s)
//           = unary(n) << 5 | offset     leaf child: n <= 3 triangles at tri_base + offset (offset <= 21)
// imask bit s is set for internal children.  Triangles: 3 x float4 each (v0 | flat id, v0 - v1, v2 - v0), 48 B.
//
// Second layout, PGRT_LAYOUT_F32 (15 x float4 = 240 B): the same node with everything the visit needs pre-decoded,
//   f0 = (child_base, tri_base, imask, 0)
//   f1, f2 = one 32-bit hit-mask word per slot: 1 << (24 + s) for an internal child, unary(n) << offset for a leaf
//   f[3 + 2*p + h] = plane p (lo.x, lo.y, lo.z, hi.x, hi.y, hi.z) of slots 4h..4h+3, as floats.
// It costs 3x the bytes and removes, per node visit, the 48 integer->float conversions (30 % of the instructions of a
// visit, profiles/r1_ncu_bvh8_q8_k_trace_k_secondary.txt) and the per-slot variable shifts that assemble the hit mask
// (another 20 %, profiles/r1_ncu_bvh8_f32_k_trace_k_shade.txt); pgrt_commit picks it while the node array fits L2 (even the
// 10 M soup turned out issue-bound, not memory-bound: profiles/r1_ncu_c5_soup_k_trace_q8.txt) and the quantised one above.
#pragma once
#include "common.cuh"

#define PGRT_LAYOUT_Q8 0
#define PGRT_LAYOUT_F32 1
#define PGRT_NODE_F4_Q8 5         // float4 per node
#define PGRT_NODE_F4_F32 15

#define PGRT_LEAF_TRIS 3          // triangles per leaf slot
#define PGRT_STACK8 40            // traversal stack entries (one pushed per level at most; commit checks the depth)

// ---- binary tree handed to the collapse: leaves 0 .. n_leaves-1 (leaf k = k-th triangle of the Morton order),
//      internal nodes above; b0 = (lo, left), b1 = (hi, right), count = triangles below.
struct Bvh2View {
    const float4* b0;
    const float4* b1;
    const uint32_t* count;
    const uint32_t* leaf_tri;     // leaf -> flat triangle id
    uint32_t n_leaves;
};

struct Wide8 { int n; uint32_t ch[8]; };

// PLOC nearest-neighbour preference.  Smaller merged area wins; equal areas are common (regular grids, instanced or
// coincident geometry) and how they resolve shapes the tree (measured: profiles/r1_sweep_ploc_ties.txt):
//   PLOC_TIES_LOWEST (default)  the lower index;
//   PLOC_TIES_BUDDY             the nearer index, then the "buddy" i^1, so that a run of identical candidates pairs up
//                               (2k, 2k+1) in one pass.  Slower trees on C2/C4, so it is only the fallback for a pass in
//                               which the default rule merges fewer than 1/16 of the clusters: with thousands of coincident
//                               triangles every cluster points at the same lowest index, one pair merges per pass and the
//                               tree becomes a chain deeper than the traversal stack.
#define PLOC_TIES_LOWEST 0
#define PLOC_TIES_BUDDY 1
struct PlocBest { float a; int j; int dist; bool buddy; };
PG_HD void ploc_best_init(PlocBest& b) { b.a = 3.402823466e+38f; b.j = -1; b.dist = 0x7fffffff; b.buddy = false; }
PG_HD void ploc_offer(PlocBest& b, float a, long i, long j, int ties) {
    const int dist = (int)(j > i ? j - i : i - j);
    const bool buddy = j == (i ^ 1L);
    const bool take = a < b.a || (ties == PLOC_TIES_BUDDY && a == b.a && (dist < b.dist || (dist == b.dist && buddy && !b.buddy)));
    if (take) { b.a = a; b.j = (int)j; b.dist = dist; b.buddy = buddy; }
}
PG_HD bool ploc_pass_stalled(uint32_t merges, uint32_t n) { return merges * 16u < n; }

PG_HD float box_half_area(float4 lo, float4 hi) {
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

// Children of the wide node rooted at binary node `root`: open the largest box first until 8 children exist;
// sub-trees of <= PGRT_LEAF_TRIS triangles stay closed while anything bigger can be opened, and are opened last
// (tighter boxes for free when slots are left over).
PG_HD void bvh8_gather(const Bvh2View& t, uint32_t root, Wide8& w) {
    w.n = 1; w.ch[0] = root;
    for (int phase = 0; phase < 2; ++phase) {
        while (w.n < 8) {
            int best = -1; float best_a = -1.0f;
            for (int k = 0; k < w.n; ++k) {
                const uint32_t c = w.ch[k];
                const uint32_t cnt = t.count[c];
                const bool openable = phase == 0 ? cnt > PGRT_LEAF_TRIS : cnt > 1;
                if (!openable) continue;
                const float a = box_half_area(t.b0[c], t.b1[c]);
                if (a > best_a) { best_a = a; best = k; }
            }
            if (best < 0) break;
            const uint32_t c = w.ch[best];
            w.ch[best] = pg_f2u(t.b0[c].w);
            w.ch[w.n++] = pg_f2u(t.b1[c].w);
        }
    }
}

PG_HD void bvh8_counts(const Bvh2View& t, const Wide8& w, int& n_internal, int& n_tris) {
    n_internal = 0; n_tris = 0;
    for (int k = 0; k < w.n; ++k) {
        const uint32_t cnt = t.count[w.ch[k]];
        if (cnt > PGRT_LEAF_TRIS) n_internal++; else n_tris += (int)cnt;
    }
}

// biased exponent eb of the smallest power of two 2^(eb-127) with lo + 255 * 2^(eb-127) >= hi
PG_HD uint32_t bvh8_exponent(float lo, float hi) {
    const float ext = hi - lo;
    uint32_t eb = 1u;
    if (ext > 0.0f) {
        const uint32_t bits = pg_f2u(ext / 255.0f);
        eb = (bits >> 23) & 0xFFu;
        if (bits & 0x7FFFFFu) eb += 1u;
        if (eb < 1u) eb = 1u;
        while (eb < 254u && !(lo + 255.0f * pg_u2f(eb << 23) >= hi)) eb += 1u;
    }
    return eb;
}

PG_HD uint32_t pack4(const uint32_t* v) { return v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24); }

PG_HD void bvh8_write_tri(const float* __restrict__ pos, uint32_t id, float4* __restrict__ out) {
    const float* p = pos + 9 * (size_t)id;
    const V3 v0 = v3(p[0], p[1], p[2]), v1 = v3(p[3], p[4], p[5]), v2 = v3(p[6], p[7], p[8]);
    const V3 e1 = v0 - v1, e2 = v2 - v0;
    out[0] = make_float4(v0.x, v0.y, v0.z, pg_u2f(id));
    out[1] = make_float4(e1.x, e1.y, e1.z, 0.0f);
    out[2] = make_float4(e2.x, e2.y, e2.z, 0.0f);
}

// Emits wide node `node_out` (5 x float4) for the gathered children, its leaf triangles at tris[tri_base ..], and the
// binary roots of its internal children in child order (int_children[r] becomes wide node child_base + r).
// Returns SAH terms of this node through sah (node term + leaf terms, un-normalised half areas).
PG_HD void bvh8_emit(const Bvh2View& t, uint32_t root, const Wide8& w, uint32_t child_base, uint32_t tri_base, const float* __restrict__ pos,
                     float4* __restrict__ node_out, float4* __restrict__ tris, uint32_t* int_children, float& sah, int layout = PGRT_LAYOUT_Q8) {
    const float4 nlo = t.b0[root], nhi = t.b1[root];
    const uint32_t ex = bvh8_exponent(nlo.x, nhi.x), ey = bvh8_exponent(nlo.y, nhi.y), ez = bvh8_exponent(nlo.z, nhi.z);
    const float sx = pg_u2f(ex << 23), sy = pg_u2f(ey << 23), sz = pg_u2f(ez << 23);
    // ---- slot assignment: greedy maximum of dot(child centre - node centre, slot direction)
    int slot_child[8]; bool child_done[8];
    for (int s = 0; s < 8; ++s) { slot_child[s] = -1; child_done[s] = false; }
    const float ncx = 0.5f * (nlo.x + nhi.x), ncy = 0.5f * (nlo.y + nhi.y), ncz = 0.5f * (nlo.z + nhi.z);
    float dx[8], dy[8], dz[8];
    for (int k = 0; k < w.n; ++k) {
        const float4 lo = t.b0[w.ch[k]], hi = t.b1[w.ch[k]];
        dx[k] = 0.5f * (lo.x + hi.x) - ncx; dy[k] = 0.5f * (lo.y + hi.y) - ncy; dz[k] = 0.5f * (lo.z + hi.z) - ncz;
    }
    for (int round = 0; round < w.n; ++round) {
        int bs = -1, bk = -1; float bc = 0.0f;
        for (int s = 0; s < 8; ++s) {
            if (slot_child[s] >= 0) continue;
            for (int k = 0; k < w.n; ++k) {
                if (child_done[k]) continue;
                const float c = ((s & 4) ? dx[k] : -dx[k]) + ((s & 2) ? dy[k] : -dy[k]) + ((s & 1) ? dz[k] : -dz[k]);
                if (bs < 0 || c > bc) { bc = c; bs = s; bk = k; }
            }
        }
        slot_child[bs] = bk; child_done[bk] = true;
    }
    // ---- per-slot encoding
    uint32_t meta[8], q[6][8];
    float fb[6][8];                     // exact child boxes (F32 layout); empty slots get an inverted box
    uint32_t imask = 0, n_int = 0, tri_off = 0;
    sah = box_half_area(nlo, nhi);
    for (int s = 0; s < 8; ++s) {
        meta[s] = 0; for (int a = 0; a < 6; ++a) { q[a][s] = 0; fb[a][s] = a < 3 ? FLT_MAX : -FLT_MAX; }
        const int k = slot_child[s];
        if (k < 0) continue;
        const uint32_t c = w.ch[k];
        const float4 lo = t.b0[c], hi = t.b1[c];
        fb[0][s] = lo.x; fb[1][s] = lo.y; fb[2][s] = lo.z; fb[3][s] = hi.x; fb[4][s] = hi.y; fb[5][s] = hi.z;
        const float clo[3] = {lo.x, lo.y, lo.z}, chi[3] = {hi.x, hi.y, hi.z}, base[3] = {nlo.x, nlo.y, nlo.z}, step[3] = {sx, sy, sz};
        for (int a = 0; a < 3; ++a) {
            float ql = floorf((clo[a] - base[a]) / step[a]);
            ql = fminf(fmaxf(ql, 0.0f), 255.0f);
            while (ql > 0.0f && base[a] + ql * step[a] > clo[a]) ql -= 1.0f;       // decoded lower bound never above the child's
            float qh = ceilf((chi[a] - base[a]) / step[a]);
            qh = fminf(fmaxf(qh, 0.0f), 255.0f);
            while (qh < 255.0f && base[a] + qh * step[a] < chi[a]) qh += 1.0f;     // decoded upper bound never below
            q[a][s] = (uint32_t)ql; q[3 + a][s] = (uint32_t)qh;
        }
        const uint32_t cnt = t.count[c];
        if (cnt > PGRT_LEAF_TRIS) {
            meta[s] = (1u << 5) | (24u + (uint32_t)s);
            imask |= 1u << s;
            int_children[n_int++] = c;
        } else {
            meta[s] = (((1u << cnt) - 1u) << 5) | tri_off;
            sah += box_half_area(lo, hi) * (float)cnt;
            // the <= 3 leaves below c
            uint32_t st[4]; int sp = 0; st[sp++] = c;
            while (sp > 0) {
                const uint32_t x = st[--sp];
                if (x < t.n_leaves) { bvh8_write_tri(pos, t.leaf_tri[x], tris + 3 * (size_t)(tri_base + tri_off)); tri_off++; }
                else { st[sp++] = pg_f2u(t.b1[x].w); st[sp++] = pg_f2u(t.b0[x].w); }
            }
        }
    }
    if (layout == PGRT_LAYOUT_F32) {
        uint32_t word[8];
        for (int s = 0; s < 8; ++s) word[s] = (imask >> s) & 1u ? 1u << (24 + s) : (meta[s] >> 5) << (meta[s] & 31u);
        node_out[0] = make_float4(pg_u2f(child_base), pg_u2f(tri_base), pg_u2f(imask), 0.0f);
        node_out[1] = make_float4(pg_u2f(word[0]), pg_u2f(word[1]), pg_u2f(word[2]), pg_u2f(word[3]));
        node_out[2] = make_float4(pg_u2f(word[4]), pg_u2f(word[5]), pg_u2f(word[6]), pg_u2f(word[7]));
        for (int a = 0; a < 6; ++a)
            for (int h = 0; h < 2; ++h) node_out[3 + 2 * a + h] = make_float4(fb[a][4 * h], fb[a][4 * h + 1], fb[a][4 * h + 2], fb[a][4 * h + 3]);
        return;
    }
    node_out[0] = make_float4(nlo.x, nlo.y, nlo.z, pg_u2f(ex | (ey << 8) | (ez << 16) | (imask << 24)));
    node_out[1] = make_float4(pg_u2f(child_base), pg_u2f(tri_base), pg_u2f(pack4(meta)), pg_u2f(pack4(meta + 4)));
    node_out[2] = make_float4(pg_u2f(pack4(q[0])), pg_u2f(pack4(q[0] + 4)), pg_u2f(pack4(q[1])), pg_u2f(pack4(q[1] + 4)));
    node_out[3] = make_float4(pg_u2f(pack4(q[2])), pg_u2f(pack4(q[2] + 4)), pg_u2f(pack4(q[3])), pg_u2f(pack4(q[3] + 4)));
    node_out[4] = make_float4(pg_u2f(pack4(q[4])), pg_u2f(pack4(q[4] + 4)), pg_u2f(pack4(q[5])), pg_u2f(pack4(q[5] + 4)));
}
