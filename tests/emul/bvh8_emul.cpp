// bvh8_emul.cpp -- CPU harness for the host/device parts of the BVH code (TEST INFRASTRUCTURE ONLY).
//
// Compiles pgi_raytracing_b200/csrc/{bvh8,traverse}.cuh with g++ (the wide-node encoder, the slot assignment, the
// quantisation and the stack traversal are plain PG_HD functions) and drives them sequentially:
//   1. a binary tree is built here by a sequential restatement of the GPU builder's PLOC pass
//      (nearest neighbour within PLOC_R in Morton order, merge mutual pairs, compact);
//   2. the tree is collapsed breadth-first with bvh8_gather / bvh8_emit, exactly as k_collapse does per wide node;
//   3. rays are traced with trace_closest8 and with a brute-force loop over tri_test.
// The product never links this file; tests/test_bvh8_emul.py loads it through ctypes.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <functional>
#include "../../pgi_raytracing_b200/csrc/traverse.cuh"

namespace {
const int PLOC_R = 8;

struct Tree {
    std::vector<float4> b0, b1; std::vector<uint32_t> count, leaf_tri;
    std::vector<float4> nodes, tris;
    std::vector<float> pos;
    uint32_t n = 0, n_nodes = 0, depth = 0, root = 0; int layout = PGRT_LAYOUT_Q8;
    double sah = 0;
};

uint64_t expand21(uint64_t v) {
    v &= 0x1FFFFFull;
    v = (v | v << 32) & 0x1F00000000FFFFull;
    v = (v | v << 16) & 0x1F0000FF0000FFull;
    v = (v | v << 8) & 0x100F00F00F00F00Full;
    v = (v | v << 4) & 0x10C30C30C30C30C3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

void tri_box(const float* p, float lo[3], float hi[3]) {
    for (int a = 0; a < 3; ++a) { lo[a] = std::min(p[a], std::min(p[3 + a], p[6 + a])); hi[a] = std::max(p[a], std::max(p[3 + a], p[6 + a])); }
}
}  // namespace

extern "C" {

void* emul_build(const float* pos, uint32_t n, int builder /*0 = PLOC, 1 = median split, 2 = binned SAH*/, int layout /*PGRT_LAYOUT_Q8 / _F32*/) {
    Tree* t = new Tree();
    t->layout = layout;
    const size_t node_f4 = layout == PGRT_LAYOUT_F32 ? PGRT_NODE_F4_F32 : PGRT_NODE_F4_Q8;
    t->n = n; t->pos.assign(pos, pos + 9 * (size_t)n);
    if (n == 0) return t;
    // Morton order (k_scene_bounds + k_morton + radix sort, restated)
    float blo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, bhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    std::vector<float> cen(3 * (size_t)n);
    for (uint32_t i = 0; i < n; ++i) {
        float lo[3], hi[3]; tri_box(pos + 9 * (size_t)i, lo, hi);
        for (int a = 0; a < 3; ++a) { const float c = 0.5f * (lo[a] + hi[a]); cen[3 * (size_t)i + a] = c; blo[a] = std::min(blo[a], c); bhi[a] = std::max(bhi[a], c); }
    }
    std::vector<std::pair<uint64_t, uint32_t>> keys(n);
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t q[3];
        for (int a = 0; a < 3; ++a) {
            const float ext = bhi[a] - blo[a];
            float f = ext > 0.0f ? (cen[3 * (size_t)i + a] - blo[a]) / ext : 0.0f;
            f = std::min(std::max(f, 0.0f), 1.0f);
            uint32_t g = (uint32_t)(f * 2097152.0f);
            q[a] = g > 2097151u ? 2097151u : g;
        }
        keys[i] = {(expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]), i};
    }
    std::stable_sort(keys.begin(), keys.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    const uint32_t total = 2 * n - 1;
    t->b0.resize(total); t->b1.resize(total); t->count.resize(total); t->leaf_tri.resize(n);
    for (uint32_t k = 0; k < n; ++k) {
        float lo[3], hi[3]; tri_box(pos + 9 * (size_t)keys[k].second, lo, hi);
        t->leaf_tri[k] = keys[k].second;
        t->b0[k] = make_float4(lo[0], lo[1], lo[2], pg_u2f(0xFFFFFFFFu));
        t->b1[k] = make_float4(hi[0], hi[1], hi[2], pg_u2f(0xFFFFFFFFu));
        t->count[k] = 1;
    }
    uint32_t next = n;
    auto merge = [&](uint32_t a, uint32_t b) {
        const float4 al = t->b0[a], ah = t->b1[a], bl = t->b0[b], bh = t->b1[b];
        t->b0[next] = make_float4(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z), pg_u2f(a));
        t->b1[next] = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), pg_u2f(b));
        t->count[next] = t->count[a] + t->count[b];
        return next++;
    };
    if (builder == 0) {
        std::vector<uint32_t> cid(n), out; std::vector<int> nn(n);
        for (uint32_t i = 0; i < n; ++i) cid[i] = i;
        uint32_t m = n;
        while (m > 1) {
            for (int mode = PLOC_TIES_LOWEST;;) {
                for (uint32_t i = 0; i < m; ++i) {
                    const float4 lo = t->b0[cid[i]], hi = t->b1[cid[i]];
                    PlocBest best; ploc_best_init(best);
                    for (int off = -PLOC_R; off <= PLOC_R; ++off) {
                        const long j = (long)i + off;
                        if (off == 0 || j < 0 || j >= (long)m) continue;
                        const float4 l2 = t->b0[cid[j]], h2 = t->b1[cid[j]];
                        const float dx = fmaxf(hi.x, h2.x) - fminf(lo.x, l2.x), dy = fmaxf(hi.y, h2.y) - fminf(lo.y, l2.y), dz = fmaxf(hi.z, h2.z) - fminf(lo.z, l2.z);
                        const float a = dx * dy + dy * dz + dz * dx;
                        ploc_offer(best, a, (long)i, j, mode);
                    }
                    nn[i] = best.j;
                }
                uint32_t merges = 0;
                for (uint32_t i = 0; i < m; ++i) merges += nn[i] > (int)i && nn[nn[i]] == (int)i;
                if (mode == PLOC_TIES_BUDDY || !ploc_pass_stalled(merges, m)) break;
                mode = PLOC_TIES_BUDDY;
            }
            out.clear();
            for (uint32_t i = 0; i < m; ++i) {
                const int j = nn[i];
                const bool mutual = j >= 0 && nn[j] == (int)i;
                if (mutual && (uint32_t)j < i) continue;
                out.push_back(mutual ? merge(cid[i], cid[j]) : cid[i]);
            }
            cid = out; m = (uint32_t)cid.size();
        }
        t->root = cid[0];
    } else if (builder == 2) {
        // Binned SAH, top-down: the family of builder behind rtcCommitScene's default MEDIUM quality (pg1/raytracer.cpp:127,
        // emb/doc/README.md:2126-2127) and the "refined by binned SAH" of the north star.  16 bins over the centroid bounds of
        // every axis, cost = area(L) * n(L) + area(R) * n(R), split down to single triangles (the collapse forms the leaves);
        // a range the bins cannot separate is cut at its median.  Here only as the yardstick the PLOC tree is measured against.
        const int NB = 16;
        std::vector<uint32_t> order(n);
        for (uint32_t i = 0; i < n; ++i) order[i] = i;
        auto cen = [&](uint32_t k, int a) { const float4 lo = t->b0[k], hi = t->b1[k]; return a == 0 ? 0.5f * (lo.x + hi.x) : (a == 1 ? 0.5f * (lo.y + hi.y) : 0.5f * (lo.z + hi.z)); };
        struct Box { float lo[3], hi[3]; void reset() { for (int a = 0; a < 3; ++a) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; } }
                     void grow(const float4& l, const float4& h) { lo[0] = fminf(lo[0], l.x); lo[1] = fminf(lo[1], l.y); lo[2] = fminf(lo[2], l.z); hi[0] = fmaxf(hi[0], h.x); hi[1] = fmaxf(hi[1], h.y); hi[2] = fmaxf(hi[2], h.z); }
                     void grow(const Box& b) { for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], b.lo[a]); hi[a] = fmaxf(hi[a], b.hi[a]); } }
                     float area() const { const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2]; return dx < 0 ? 0.0f : dx * dy + dy * dz + dz * dx; } };
        std::function<uint32_t(uint32_t, uint32_t)> rec = [&](uint32_t lo, uint32_t hi) -> uint32_t {
            if (hi - lo == 1) return order[lo];
            float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (uint32_t i = lo; i < hi; ++i) for (int a = 0; a < 3; ++a) { const float c = cen(order[i], a); clo[a] = fminf(clo[a], c); chi[a] = fmaxf(chi[a], c); }
            float best = FLT_MAX; int best_axis = -1, best_bin = 0;
            for (int a = 0; a < 3; ++a) {
                const float ext = chi[a] - clo[a];
                if (!(ext > 0.0f)) continue;
                Box bb[NB]; uint32_t bn[NB];
                for (int b = 0; b < NB; ++b) { bb[b].reset(); bn[b] = 0; }
                const float scale = NB / ext;
                for (uint32_t i = lo; i < hi; ++i) {
                    int b = (int)((cen(order[i], a) - clo[a]) * scale); b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                    bb[b].grow(t->b0[order[i]], t->b1[order[i]]); bn[b]++;
                }
                float ra[NB]; uint32_t rn[NB]; Box acc; acc.reset(); uint32_t cnt = 0;
                for (int b = NB - 1; b > 0; --b) { acc.grow(bb[b]); cnt += bn[b]; ra[b] = acc.area(); rn[b] = cnt; }
                acc.reset(); cnt = 0;
                for (int b = 0; b < NB - 1; ++b) {
                    acc.grow(bb[b]); cnt += bn[b];
                    if (cnt == 0 || rn[b + 1] == 0) continue;
                    const float cost = acc.area() * cnt + ra[b + 1] * rn[b + 1];
                    if (cost < best) { best = cost; best_axis = a; best_bin = b; }
                }
            }
            uint32_t mid = (lo + hi) / 2;
            if (best_axis >= 0) {
                const float scale = NB / (chi[best_axis] - clo[best_axis]);
                auto it = std::partition(order.begin() + lo, order.begin() + hi, [&](uint32_t k) {
                    int b = (int)((cen(k, best_axis) - clo[best_axis]) * scale); b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                    return b <= best_bin; });
                const uint32_t m2 = (uint32_t)(it - order.begin());
                if (m2 > lo && m2 < hi) mid = m2;
            }
            const uint32_t a = rec(lo, mid), b = rec(mid, hi);
            return merge(a, b);
        };
        t->root = rec(0, n);
    } else {
        // median split over the Morton order (any binary tree must give the same hits)
        struct Job { uint32_t lo, hi; };
        std::vector<uint32_t> stack_node;
        std::function<uint32_t(uint32_t, uint32_t)> rec = [&](uint32_t lo, uint32_t hi) -> uint32_t {
            if (hi - lo == 1) return lo;
            const uint32_t mid = (lo + hi) / 2;
            const uint32_t a = rec(lo, mid), b = rec(mid, hi);
            return merge(a, b);
        };
        t->root = rec(0, n);
    }
    // ---- collapse, breadth first (what k_collapse does with one thread per wide node)
    Bvh2View v{t->b0.data(), t->b1.data(), t->count.data(), t->leaf_tri.data(), n};
    t->nodes.resize(node_f4 * (size_t)std::max<uint32_t>(n, 1)); t->tris.resize(3 * (size_t)n);
    std::vector<std::pair<uint32_t, uint32_t>> cur{{t->root, 0u}}, nxt;
    uint32_t node_count = 1, tri_count = 0;
    while (!cur.empty()) {
        t->depth++;
        nxt.clear();
        for (auto [b2, wi] : cur) {
            Wide8 w; bvh8_gather(v, b2, w);
            int ni, nt; bvh8_counts(v, w, ni, nt);
            const uint32_t cb = node_count, tb = tri_count;
            node_count += ni; tri_count += nt;
            uint32_t ic[8]; float sah;
            bvh8_emit(v, b2, w, cb, tb, t->pos.data(), &t->nodes[node_f4 * (size_t)wi], t->tris.data(), ic, sah, layout);
            t->sah += sah;
            for (int r = 0; r < ni; ++r) nxt.push_back({ic[r], cb + r});
        }
        cur.swap(nxt);
    }
    t->n_nodes = node_count;
    if (tri_count != n) { fprintf(stderr, "emul_build: %u triangles emitted, %u expected\n", tri_count, n); }
    t->sah /= std::max(1e-30, (double)box_half_area(t->b0[t->root], t->b1[t->root]));
    return t;
}

void emul_free(void* h) { delete (Tree*)h; }
uint32_t emul_nodes(void* h) { return ((Tree*)h)->n_nodes; }
uint32_t emul_depth(void* h) { return ((Tree*)h)->depth; }
double emul_sah(void* h) { return ((Tree*)h)->sah; }
const float* emul_node_data(void* h) { return (const float*)((Tree*)h)->nodes.data(); }
const float* emul_tri_data(void* h) { return (const float*)((Tree*)h)->tris.data(); }

// rays: n x 8 floats (org, tnear, dir, tfar); out: n x 4 floats (t, u, v, tri id bits); stats: n x 2 uint32 (nodes, tris)
void emul_trace(void* h, const float* rays, uint64_t n, float* out, uint32_t* stats, int brute) {
    Tree* t = (Tree*)h;
    for (uint64_t i = 0; i < n; ++i) {
        const float* r = rays + 8 * i;
        const V3 O = v3(r[0], r[1], r[2]), D = v3(r[4], r[5], r[6]);
        HitRec best;
        TravCount tc; tc.nodes = 0; tc.tris = 0;
        if (brute & 1) {
            best.t = r[7]; best.u = 0; best.v = 0; best.tri = PGRT_INVALID_ID;
            for (uint32_t k = 0; k < t->n; ++k) tri_test(t->tris.data(), k, O, D, r[3], r[7], best);
        } else {
            const bool ww = (brute & 2) != 0;   // bit 1 of `brute` selects the while-while loop shape
            best = t->layout == PGRT_LAYOUT_F32 ? trace_closest8f<true>(t->nodes.data(), t->tris.data(), t->n, O, D, r[3], r[7], tc, ww)
                                                : trace_closest8<true>(t->nodes.data(), t->tris.data(), t->n, O, D, r[3], r[7], tc, ww);
        }
        out[4 * i] = best.t; out[4 * i + 1] = best.u; out[4 * i + 2] = best.v; memcpy(&out[4 * i + 3], &best.tri, 4);
        if (stats) { stats[2 * i] = tc.nodes; stats[2 * i + 1] = tc.tris; }
    }
}

// structural check: every node's decoded child boxes contain the true boxes below; returns the number of violations
uint64_t emul_check(void* h) {
    Tree* t = (Tree*)h;
    if (t->n == 0 || t->layout != PGRT_LAYOUT_Q8) return 0;   // the structural walk decodes the quantised layout
    uint64_t bad = 0;
    std::vector<uint8_t> seen(t->n, 0);
    // exact box of the subtree under wide node wi, while checking each decoded slot box against it
    std::function<void(uint32_t, float*, float*)> rec = [&](uint32_t wi, float* lo, float* hi) {
        const float4* nd = &t->nodes[5 * (size_t)wi];
        const uint32_t eb = pg_f2u(nd[0].w), imask = eb >> 24;
        const float sc[3] = {pg_u2f((eb & 0xFF) << 23), pg_u2f(((eb >> 8) & 0xFF) << 23), pg_u2f(((eb >> 16) & 0xFF) << 23)};
        const float base[3] = {nd[0].x, nd[0].y, nd[0].z};
        const uint32_t cb = pg_f2u(nd[1].x), tb = pg_f2u(nd[1].y);
        const uint32_t w2[4] = {pg_f2u(nd[2].x), pg_f2u(nd[2].y), pg_f2u(nd[2].z), pg_f2u(nd[2].w)};
        const uint32_t w3[4] = {pg_f2u(nd[3].x), pg_f2u(nd[3].y), pg_f2u(nd[3].z), pg_f2u(nd[3].w)};
        const uint32_t w4[4] = {pg_f2u(nd[4].x), pg_f2u(nd[4].y), pg_f2u(nd[4].z), pg_f2u(nd[4].w)};
        for (int a = 0; a < 3; ++a) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; }
        for (int s = 0; s < 8; ++s) {
            const uint32_t meta = (pg_f2u(s < 4 ? nd[1].z : nd[1].w) >> (8 * (s & 3))) & 0xFF;
            if (!meta) continue;
            auto byte = [&](uint32_t w0, uint32_t w1) { return (float)(((s < 4 ? w0 : w1) >> (8 * (s & 3))) & 0xFF); };
            const float qlo[3] = {byte(w2[0], w2[1]), byte(w2[2], w2[3]), byte(w3[0], w3[1])};
            const float qhi[3] = {byte(w3[2], w3[3]), byte(w4[0], w4[1]), byte(w4[2], w4[3])};
            float clo[3], chi[3];
            if ((imask >> s) & 1) {
                if ((meta >> 5) != 1 || (meta & 31) != 24u + s) bad++;
                rec(cb + pg_popc(imask & ((1u << s) - 1u)), clo, chi);
            } else {
                const uint32_t cnt = pg_popc(meta >> 5), off = meta & 31;
                for (int a = 0; a < 3; ++a) { clo[a] = FLT_MAX; chi[a] = -FLT_MAX; }
                for (uint32_t k = 0; k < cnt; ++k) {
                    const float4* q = &t->tris[3 * (size_t)(tb + off + k)];
                    const uint32_t id = pg_f2u(q[0].w);
                    if (id >= t->n || seen[id]) bad++; else seen[id] = 1;
                    float l[3], hh[3]; tri_box(&t->pos[9 * (size_t)id], l, hh);
                    for (int a = 0; a < 3; ++a) { clo[a] = std::min(clo[a], l[a]); chi[a] = std::max(chi[a], hh[a]); }
                }
            }
            for (int a = 0; a < 3; ++a) {
                if (!(base[a] + qlo[a] * sc[a] <= clo[a]) || !(base[a] + qhi[a] * sc[a] >= chi[a])) bad++;
                lo[a] = std::min(lo[a], clo[a]); hi[a] = std::max(hi[a], chi[a]);
            }
        }
    };
    float lo[3], hi[3];
    rec(0, lo, hi);
    for (uint32_t i = 0; i < t->n; ++i) if (!seen[i]) bad++;
    return bad;
}

}  // extern "C"
