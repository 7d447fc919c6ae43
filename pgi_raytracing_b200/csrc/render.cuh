// render.cuh -- GPU form of Raytracer::get_pixel / trace (pg1/raytracer.cpp:237-437).
//
// The reference recurses: trace(ray, level) calls itself for the reflection and the refraction ray of every
// dielectric hit and combines the two results with the NON-linear mix_srgb (raytracer.cpp:318-319), so no
// linear path throughput exists: every dielectric hit is a node that waits for its children.
//
// Two schedulers (pgrt_render_params.scheduler), bit-identical in every output:
//
//   0  fused (default): ONE persistent kernel per batch, k_frame, runs the whole of trace() for every sample.
//      A warp claims either 32 primary rays (regenerated from the slot index: coherent 8x4 pixel blocks) or up to 32
//      records of the frame's ray pool, runs the closest-hit query, classifies the hit, evaluates Phong with its
//      shadow query inline, and for a dielectric hit appends the reflection / refraction rays to the pool, where ANY
//      warp of the grid may trace them: records are handed out by ticket (atomicAdd on q_head, 32 records per ticket, never
//      a retry), published with a per-record epoch word, and a ticket may run ahead of the records -- its holder polls its
//      own 32 publication words.  The non-linear combine is resolved by continuation: a finished node decrements its
//      parent's pending count, and the last child to arrive evaluates mix_srgb for the parent and keeps climbing.  With one
//      sample per pixel the warp that finishes a pixel applies gamma and stores it in the frame, so neither hits,
//      directions nor colours of level 0 ever travel through HBM.  `outstanding` counts primary chunks + published records
//      not yet retired (children are added BEFORE they are published, so it is 0 only when the batch is done).  The first
//      keep_ctas CTAs stay and poll until it is 0; the others leave when they find neither a primary chunk nor a record,
//      which frees their SM slots for the next frame's kernel.  Every wait is bounded (Counters::watchdog reports, never hangs).
//   1  level-synchronous wavefront, one queue per recursion level (the cross-check of scheduler 0):
//        forward,  level = 0 .. max_depth :  k_trace -> k_shade -> k_phong
//        backward, level = max_depth-1 .. 0 : k_combine resolves each dielectric node from its two children (post-order)
//        k_resolve averages the samples of a pixel in sx-major order and applies gamma (raytracer.cpp:421-446).
//
// Pixel order: 32x8 tiles dealt round-robin to ranks; inside a tile 8x4 blocks so a warp's primary rays are coherent.
#pragma once
#include "common.cuh"
#include "traverse.cuh"
#include "shading.cuh"

struct LevelBufs {
    float4* ray_o;          // org.xyz, tnear
    float4* ray_d;          // dir.xyz, time (= IOR of the medium, raytracer.cpp:200,228,416; < 0 marks an unused slot)
    float4* hit;            // t, u, v, flat triangle id
    float4* color;          // value trace() returns for this node
    float4* dn_att;         // dielectric nodes: attenuation rgb, R
    uint2* dn_child;        // dielectric nodes: level+1 slots of the reflection / refraction child (PGRT_INVALID_ID = none)
    uint32_t* phong_list;
    uint32_t* diel_list;
    uint32_t* pending;      // dynamic scheduler, level 0 only: children of a dielectric node not yet resolved
    uint32_t cap;
};

// Pool of every ray of level >= 1 (fused scheduler).  A record is written once, by the warp that shaded its parent, and
// may be traced by any warp; everything that crosses warps (records, colours, pending counts) goes through ld.cg / st.cg,
// fences and atomics only: L1 is not coherent between SMs and a 32-B sector holds two neighbouring records.
struct RayPool {
    float4* ray_o;          // org.xyz, tnear
    float4* ray_d;          // dir.xyz, time; time < 0 marks a reserved-but-unused slot
    float4* color;          // value trace() returns for this node
    float4* att;            // dielectric nodes: attenuation rgb, R
    uint2* child;           // dielectric nodes: pool slots of the reflection / refraction child
    uint4* link;            // x = parent record, y = level | child bit << 8 | (parent is a level-0 record) << 9,
                            // z = epoch of the batch that wrote the record: stored LAST, it publishes the record
    uint32_t* pending;      // dielectric nodes: children not yet resolved
    uint32_t cap;
};

// Where the resolved pixels of a frame go (K11)
struct FrameOut {
    float4* rgba;           // float frame in the layout of tex_data_ (simpleguidx11.cpp:108-114) or compact shard buffer; may be null
    uint32_t* rgba8;        // R8G8B8A8_UNORM frame, the format the reference displays (simpleguidx11.cpp:229), or null
    int32_t compact;        // 1: addressed by shard slot, 0: by y * W + x
    int32_t direct;         // 1 (one sample per pixel): k_frame resolves the pixel itself; 0: samples go through L0.color and k_resolve
};

struct Counters {
    uint32_t n_rays[PGRT_MAX_LEVELS + 1];
    uint32_t n_phong[PGRT_MAX_LEVELS + 1];
    uint32_t n_diel[PGRT_MAX_LEVELS + 1];
    uint32_t overflow;
    uint32_t watchdog;      // fused scheduler: a bounded wait expired (internal error; wd_info says what the warp saw)
    uint32_t q_peak;        // most pool records any batch of this frame allocated (sizes the pool of the next frames)
    unsigned long long shadow, reflection, refraction;          // this batch
    unsigned long long tot_shadow, tot_reflection, tot_refraction, tot_primary;   // this frame
    // per-level frame totals; the traversal columns are filled by instrumented renders only (profile bit 1)
    unsigned long long lv_rays[PGRT_MAX_LEVELS + 1], lv_shadow[PGRT_MAX_LEVELS + 1];
    unsigned long long lv_nodes[PGRT_MAX_LEVELS + 1], lv_tris[PGRT_MAX_LEVELS + 1];
    unsigned long long lv_sh_nodes[PGRT_MAX_LEVELS + 1], lv_sh_tris[PGRT_MAX_LEVELS + 1];
    uint32_t lv_max_nodes[PGRT_MAX_LEVELS + 1], lv_sh_max_nodes[PGRT_MAX_LEVELS + 1];
    // fused scheduler: pool records allocated / claimed; epoch = batches begun on this slot (publication word of the records);
    // done_seq = frames finished without a queue overflow (the completion flag of the slot, see k_batch_end)
    // The three words a warp of k_frame looks at before every claim sit in one aligned 16-byte group (one volatile load):
    // outstanding = primary chunks + published pool records not yet fully processed (0 = batch done)
    alignas(16) uint32_t q_tail;
    uint32_t q_head, outstanding, epoch;
    uint32_t done_seq;
    uint32_t sig_value;     // what a finished frame stores in the slot's external completion flag (pgrt_slot_signal)
    uint32_t pool_iters;    // fused scheduler: warp iterations spent on pool records (rays of level >= 1 / this = lanes per iteration)
    uint32_t pad2[1];
    uint32_t wd_info[8];    // what the first warp whose bounded wait expired saw (diagnostics of the watchdog)
    unsigned long long warp_cycles[4];   // builds with -DPGRT_FRAME_TIMING only: SM cycles the frame kernel's warps spent on primary chunks, on pool records, resident in total; [3] warps
    unsigned long long t_first, t_primary_done, t_last;   // fused scheduler: %globaltimer marks (ns): first warp in, primary rays exhausted, last warp out
    unsigned long long lv_t_first[PGRT_MAX_LEVELS + 1], lv_t_last[PGRT_MAX_LEVELS + 1];   // ... and per level: first ray taken up, last ray finished
    uint32_t trace_next[PGRT_MAX_LEVELS + 1];   // k_trace: rays of the level's queue claimed so far
};

__device__ __forceinline__ void flush_trav_counts(unsigned long long nodes, unsigned long long tris, uint32_t mx,
                                                  unsigned long long* g_nodes, unsigned long long* g_tris, uint32_t* g_max) {
    for (int o = 16; o > 0; o >>= 1) {
        nodes += __shfl_xor_sync(0xffffffffu, nodes, o); tris += __shfl_xor_sync(0xffffffffu, tris, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0 && nodes) { atomicAdd(g_nodes, nodes); atomicAdd(g_tris, tris); atomicMax(g_max, mx); }
}

struct ShardInfo { int32_t rank, n_ranks, tiles_x, tiles_y; };

__device__ __forceinline__ bool slot_to_pixel(const ShardInfo& sh, int W, int H, uint32_t slot, int& x, int& y) {
    const uint32_t k = slot / PGRT_TILE_PIXELS, q = slot % PGRT_TILE_PIXELS;
    const uint32_t t = k * (uint32_t)sh.n_ranks + (uint32_t)sh.rank;
    if (t >= (uint32_t)(sh.tiles_x * sh.tiles_y)) return false;
    const uint32_t tx = t % (uint32_t)sh.tiles_x, ty = t / (uint32_t)sh.tiles_x;
    const uint32_t b = q >> 5, l = q & 31u;
    x = (int)(tx * PGRT_TILE_W + (b & 3u) * 8u + (l & 7u));
    y = (int)(ty * PGRT_TILE_H + (b >> 2) * 4u + (l >> 3));
    return x < W && y < H;
}

// ---- K6: primary rays (raytracer.cpp:405-416 + PinHoleCamera::generate_ray)
__device__ __forceinline__ RayRec primary_ray(const DevCamera& cam, const pgrt_render_params& p, int xp, int yp, int s) {
    const int w = p.sampling_width;
    const int sx = s / w, sy = s % w;
    const uint32_t pixel = (uint32_t)(yp * cam.width + xp);
    float rand1 = 0.0f, rand2 = 0.0f;
    if (p.jitter) {
        rand1 = rng_uniform(-0.5f / w, 0.5f / w, rng_u01(p.seed, pixel, (uint32_t)s, 0));
        rand2 = rng_uniform(-0.5f / w, 0.5f / w, rng_u01(p.seed, pixel, (uint32_t)s, 1));
    }
    const float new_x = (float)xp + ((float)sx * (1.0f / w)) + rand1;
    const float new_y = (float)yp + ((float)sy * (1.0f / w)) + rand2;
    RayRec r;
    if (p.camera_mode == 1) r = camera_ray_pinhole(cam, new_x, new_y);
    else {
        // aperture 0: (0 - (-0)) * u + (-0) is +0 for every u, so the two hashes are skipped
        const float l1 = p.aperture != 0.0f ? rng_uniform(-p.aperture / 2.0f, p.aperture / 2.0f, rng_u01(p.seed, pixel, (uint32_t)s, 2)) : 0.0f;
        const float l2 = p.aperture != 0.0f ? rng_uniform(-p.aperture / 2.0f, p.aperture / 2.0f, rng_u01(p.seed, pixel, (uint32_t)s, 3)) : 0.0f;
        r = camera_ray_lens(cam, new_x, new_y, p.focal_distance, l1, l2);
    }
    r.time = PGRT_IOR_AIR;   // raytracer.cpp:416
    return r;
}

__global__ void __launch_bounds__(256) k_raygen(DevCamera cam, pgrt_render_params p, ShardInfo sh, uint32_t slot0, uint32_t n_slots,
                                                LevelBufs L, Counters* cnt) {
    const int S = p.sampling_width * p.sampling_width;
    const uint32_t n = n_slots * (uint32_t)S;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) cnt->n_rays[0] = n;
    if (j >= n) return;
    const uint32_t slot = slot0 + j / (uint32_t)S;
    const int s = (int)(j % (uint32_t)S);
    int x, y;
    if (!slot_to_pixel(sh, cam.width, cam.height, slot, x, y)) {
        L.ray_o[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        L.ray_d[j] = make_float4(0.f, 0.f, 0.f, -1.0f);
        return;
    }
    const RayRec r = primary_ray(cam, p, x, y, s);
    L.ray_o[j] = make_float4(r.o.x, r.o.y, r.o.z, r.tnear);
    L.ray_d[j] = make_float4(r.d.x, r.d.y, r.d.z, r.time);
}

// Level 0 without a stored ray queue: the primary ray of slot j is a pure function of (camera, params, shard, j), so
// k_trace / k_shade / k_phong regenerate it instead of reading 32 B that k_raygen would have had to write first
// (66 MB written + 3 x 66 MB read per 1080p frame otherwise); it stores the direction (16 B) for the shading kernels,
// which recompute only the origin.  `on` = 0 reads the level's stored rays.
struct Gen0 { DevCamera cam; ShardInfo sh; uint32_t slot0, n_slots; int spp; int on; };

// origin only (the part of primary_ray that needs no normalisation): camera position plus the lens shift
__device__ __forceinline__ float4 primary_origin(const DevCamera& cam, const pgrt_render_params& p, int xp, int yp, int s) {
    if (p.camera_mode == 1) return make_float4(cam.from.x, cam.from.y, cam.from.z, 0.001f);
    const uint32_t pixel = (uint32_t)(yp * cam.width + xp);
    const float l1 = p.aperture != 0.0f ? rng_uniform(-p.aperture / 2.0f, p.aperture / 2.0f, rng_u01(p.seed, pixel, (uint32_t)s, 2)) : 0.0f;
    const float l2 = p.aperture != 0.0f ? rng_uniform(-p.aperture / 2.0f, p.aperture / 2.0f, rng_u01(p.seed, pixel, (uint32_t)s, 3)) : 0.0f;
    const V3 shift = mul3(cam.M, v3(l1, l2, 0.0f));
    return make_float4(cam.from.x + shift.x, cam.from.y + shift.y, cam.from.z + shift.z, 0.01f);
}

// k_shade / k_phong at level 0: k_trace has stored the direction it generated (16 B); only the origin is recomputed
__device__ __forceinline__ void load_ray_shade(const LevelBufs& L, const Gen0& g, const pgrt_render_params& p, uint32_t j, float4& o, float4& d) {
    d = L.ray_d[j];
    if (!g.on) { o = L.ray_o[j]; return; }
    o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d.w < 0.0f) return;
    int x, y;
    slot_to_pixel(g.sh, g.cam.width, g.cam.height, g.slot0 + j / (uint32_t)g.spp, x, y);
    o = primary_origin(g.cam, p, x, y, (int)(j % (uint32_t)g.spp));
}

__device__ __forceinline__ void load_ray(const LevelBufs& L, const Gen0& g, const pgrt_render_params& p, uint32_t j, float4& o, float4& d) {
    if (!g.on) { o = L.ray_o[j]; d = L.ray_d[j]; return; }
    const uint32_t slot = g.slot0 + j / (uint32_t)g.spp;
    int x, y;
    if (!slot_to_pixel(g.sh, g.cam.width, g.cam.height, slot, x, y)) { o = make_float4(0.f, 0.f, 0.f, 0.f); d = make_float4(0.f, 0.f, 0.f, -1.0f); return; }
    const RayRec r = primary_ray(g.cam, p, x, y, (int)(j % (uint32_t)g.spp));
    o = make_float4(r.o.x, r.o.y, r.o.z, r.tnear);
    d = make_float4(r.d.x, r.d.y, r.d.z, r.time);
}

// ---- K7: closest hit for one queue (get_ray_hit, raytracer.cpp:130-148).
// Persistent warps with ray replacement: a warp claims rays from the queue counter, traverses them a few while-while
// rounds at a time, and whenever at least `refill` of its lanes have finished it claims new rays for exactly those
// lanes (ballot + one atomicAdd per warp), so a few long traversals do not keep 31 lanes idle until they end.
// refill = 32 degenerates to "a fresh 32-ray chunk when the whole warp is done".
#define PGRT_TRACE_ROUNDS 2
__device__ __forceinline__ void set_top(RayCtxF& r, uint32_t top) {
#ifdef PGRT_SMEM_TOP
    r.top = top;
#endif
}
__device__ __forceinline__ void set_top(RayCtxQ&, uint32_t) {}

template <class RC, bool COUNT>
__device__ __forceinline__ void trace_queue(const DevScene& sc, const pgrt_render_params& p, const Gen0& g0, const LevelBufs& L, int level, Counters* cnt,
                                            uint32_t n, int refill, uint32_t top = 0u) {
    const int lane = threadIdx.x & 31;
    uint32_t* next = &cnt->trace_next[level];
    uint2 stack[PGRT_STACK8];
    RC r; TravState s; TravCount tc;
    trav_init(s, FLT_MAX);
    tc.nodes = 0; tc.tris = 0;
    uint32_t j = PGRT_INVALID_ID;
    bool more = true;
    unsigned long long my_nodes = 0, my_tris = 0; uint32_t my_max = 0;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, j == PGRT_INVALID_ID);
        if (more && ((int)__popc(idle) >= refill || idle == 0xffffffffu)) {
            const uint32_t need = (uint32_t)__popc(idle);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(next, need);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + need >= n) more = false;
            if (j == PGRT_INVALID_ID) {
                const uint32_t mine = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
                if (mine < n) {
                    float4 o, d;
                    load_ray(L, g0, p, mine, o, d);
                    if (g0.on) L.ray_d[mine] = d;      // the shading kernels read it back instead of re-normalising three times
                    if (d.w >= 0.0f && sc.n_tris != 0 && !ray_misses_box(sc.bb_lo, sc.bb_hi, v3(o.x, o.y, o.z), v3(d.x, d.y, d.z), o.w, FLT_MAX)) {
                        j = mine;
                        ray_ctx_init(r, v3(o.x, o.y, o.z), v3(d.x, d.y, d.z), o.w, FLT_MAX, sc.loop_ww != 0);
                        set_top(r, top);
                        trav_init(s, FLT_MAX);
                        tc.nodes = 0; tc.tris = 0;
                    } else {
                        L.hit[mine] = make_float4(FLT_MAX, 0.f, 0.f, __uint_as_float(PGRT_INVALID_ID));   // unused slot / empty scene
                    }
                }
            }
        }
        if (__ballot_sync(0xffffffffu, j != PGRT_INVALID_ID) == 0u) { if (!more) break; continue; }
        if (j != PGRT_INVALID_ID && trav_advance<RC, COUNT>(sc.nodes, sc.tris, r, s, stack, tc, PGRT_TRACE_ROUNDS)) {
            L.hit[j] = make_float4(s.best.t, s.best.u, s.best.v, __uint_as_float(s.best.tri));
            if (COUNT) { my_nodes += tc.nodes; my_tris += tc.tris; my_max = max(my_max, tc.nodes); }
            j = PGRT_INVALID_ID;
        }
    }
    if (COUNT) flush_trav_counts(my_nodes, my_tris, my_max, &cnt->lv_nodes[level], &cnt->lv_tris[level], &cnt->lv_max_nodes[level]);
}

#ifndef PGRT_TRACE_MIN_BLOCKS
#define PGRT_TRACE_MIN_BLOCKS 6   // register cap of k_trace = 65536 / (128 * this)
#endif
template <bool COUNT>
__global__ void __launch_bounds__(128, PGRT_TRACE_MIN_BLOCKS) k_trace(DevScene sc, pgrt_render_params p, Gen0 g0, LevelBufs L, int level, int refill, Counters* cnt) {
    const uint32_t n = min(cnt->n_rays[level], L.cap);
    uint32_t top = 0u;
#ifdef PGRT_SMEM_TOP
    // measurement build: the first PGRT_SMEM_TOP nodes (the root and what follows it) staged in shared memory
    __shared__ float4 s_top[PGRT_SMEM_TOP * PGRT_NODE_F4_F32];
    if (sc.node_layout == PGRT_LAYOUT_F32 && sc.n_tris >= 64u) {
        for (int k = threadIdx.x; k < PGRT_SMEM_TOP * PGRT_NODE_F4_F32; k += blockDim.x) s_top[k] = sc.nodes[k];
        __syncthreads();
        top = (uint32_t)__cvta_generic_to_shared(s_top);
    }
#endif
    if (sc.node_layout == PGRT_LAYOUT_F32) trace_queue<RayCtxF, COUNT>(sc, p, g0, L, level, cnt, n, refill, top);
    else trace_queue<RayCtxQ, COUNT>(sc, p, g0, L, level, cnt, n, refill);
}

// rtcInterpolate0 (raytracer.cpp:252, :344): w*a0 + u*a1 + v*a2, fused as Embree's madd chain
__device__ __forceinline__ void fetch_shading(const DevScene& sc, uint32_t tri, float u, float v, V3& n, float& tu, float& tv, uint32_t& geom) {
    const float4 q0 = __ldg(sc.shade + 4 * (size_t)tri), q1 = __ldg(sc.shade + 4 * (size_t)tri + 1);
    const float4 q2 = __ldg(sc.shade + 4 * (size_t)tri + 2), q3 = __ldg(sc.shade + 4 * (size_t)tri + 3);
    const float w = 1.0f - u - v;
    n.x = __fmaf_rn(w, q0.x, __fmaf_rn(u, q1.x, v * q2.x));
    n.y = __fmaf_rn(w, q0.y, __fmaf_rn(u, q1.y, v * q2.y));
    n.z = __fmaf_rn(w, q0.z, __fmaf_rn(u, q1.z, v * q2.z));
    // uv corners: t0 = (q0.w, q1.w), t1 = (q2.w, q3.x), t2 = (q3.y, q3.z)
    tu = __fmaf_rn(w, q0.w, __fmaf_rn(u, q2.w, v * q3.y));
    tv = __fmaf_rn(w, q1.w, __fmaf_rn(u, q3.x, v * q3.z));
    geom = __float_as_uint(q3.w);
}

// common hit frame of trace() (raytracer.cpp:244-272)
struct HitFrame { V3 dirn, n, hitp; uint32_t geom; float tu, tv; };
__device__ __forceinline__ HitFrame hit_frame(const DevScene& sc, float4 o, float4 d, float4 h) {
    HitFrame f;
    f.dirn = normalize3(v3(d.x, d.y, d.z));
    V3 n;
    fetch_shading(sc, __float_as_uint(h.w), h.y, h.z, n, f.tu, f.tv, f.geom);
    n = normalize3(n);
    f.hitp = v3(o.x, o.y, o.z) + v3(d.x, d.y, d.z) * h.x;
    if (dot3(n, f.dirn) > 0) n = -n;
    f.n = n;
    return f;
}

__device__ __forceinline__ uint32_t warp_append(uint32_t* counter, bool pred, int lane) {
    // all 32 lanes call; returns the slot of each lane with pred set
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    uint32_t base = 0;
    if (lane == 0 && m) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}

// ---- K8 (per ray): classification of one (ray, hit) pair = the head of trace() (raytracer.cpp:247-323, :390-393)
#ifndef PGRT_SHADE_MIN_BLOCKS
#define PGRT_SHADE_MIN_BLOCKS 6   // register cap of k_shade = 65536 / (256 * this) = 40.  With many frames in flight what a kernel
                                  // leaves free matters more than its own speed: fewer registers = more CTAs of OTHER frames' kernels
                                  // resident beside it (profiles/r1_sweep_register_caps_pipelined.txt: C2 6.69 -> 7.22 Grays/s)
#endif
#ifndef PGRT_PHONG_MIN_BLOCKS
#define PGRT_PHONG_MIN_BLOCKS 8   // register cap of k_phong = 65536 / (128 * this) = 64
#endif
enum ShadeKind { SK_FINAL = 0, SK_PHONG = 1, SK_DIEL = 2, SK_PATH = 3 };   // SK_PATH: Phong value + one diffuse bounce (shader_mode 3)
struct ShadeOut {
    int kind;
    float4 color;           // SK_FINAL: the value trace() returns
    HitFrame f;             // SK_PHONG / SK_DIEL
    RayRec refl, refr;      // SK_DIEL
    float4 att;             // SK_DIEL: attenuation rgb, R;  SK_PATH: diffuse albedo in output channel order, w = -1 (marks the node)
    bool has_refr;          // SK_DIEL: the refraction ray exists (no total internal reflection)
};

// m_d of the Phong sum (raytracer.cpp:338-348): Kd with the x<->z naming swap, or the texel of the diffuse map
__device__ __forceinline__ void diffuse_albedo(const DevScene& sc, const pgrt_material& mat, const HitFrame& f, float& m_d_r, float& m_d_g, float& m_d_b) {
    if (mat.diffuse_tex < 0 || mat.diffuse_tex >= sc.n_textures) {         // :338-341
        m_d_r = mat.diffuse[2]; m_d_g = mat.diffuse[1]; m_d_b = mat.diffuse[0];
    } else {                                                                // :343-348
        const Col3 texel = tex_get_texel(sc.textures[mat.diffuse_tex], f.tu, 1.0f - f.tv);
        m_d_r = texel.r; m_d_g = texel.g; m_d_b = texel.b;
    }
}

// The bulky double-precision leaves (environment look-up: atan2 / asin in double; dielectric set-up: three exp, one pow; the
// sRGB mix: six pow).  As real functions (-DPGRT_COLD_NOINLINE) they would keep the register count and the instruction-cache
// footprint of the kernels down, but measured on C2 the calls cost more than they save (0.59 vs 0.64 ms per pipelined frame,
// profiles/r2_sweep_kframe_builds.txt): inlined by default.
#ifdef PGRT_COLD_NOINLINE
#define PGRT_COLD __noinline__
#else
#define PGRT_COLD __forceinline__
#endif
__device__ PGRT_COLD float4 env_color(const DevTexture* env, float x, float y, float z) {
    const Col4 c = env_get_texel(*env, x, y, z);
    return make_float4(c.r, c.g, c.b, c.a);
}
__device__ PGRT_COLD float4 diel_attenuation(float kx, float ky, float kz, float t, float n1, float n2, float cos1, bool with_r) {
    float4 att;
    att.x = f_expf(-(1 - kx) * t); att.y = f_expf(-(1 - ky) * t); att.z = f_expf(-(1 - kz) * t); att.w = 0.0f;   // raytracer.cpp:303-305
    if (with_r) {
        const float alpha = (n1 - n2) / (n1 + n2);
        // :316  pow(1 - cos1, 5) in double: five exact-ish products (<= 2 ulp of a double, invisible after the cast to float)
        const double q1 = (double)(1 - cos1), q2 = q1 * q1;
        att.w = (float)((double)(alpha * alpha + (1 - (alpha * alpha))) * (q2 * q2 * q1));
    }
    return att;
}

// PATH: compiled with the path-tracing branch (shader_mode 3).  The default instantiation carries none of it.
template <bool PATH>
__device__ __forceinline__ void shade_classify(const DevScene& sc, const pgrt_render_params& p, int level, float4 o, float4 d, float4 h, ShadeOut& s) {
    s.kind = SK_FINAL; s.has_refr = false;
    s.color = make_float4(0.f, 0.f, 0.f, 1.f);
    s.att = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t tri = __float_as_uint(h.w);
    if (d.w < 0.0f) return;                                             // unused slot of a partial tile / dead pool slot
    if (tri == PGRT_INVALID_ID) {                                       // :390-393
        const V3 dirn = normalize3(v3(d.x, d.y, d.z));
        s.color = env_color(&sc.env, dirn.x, dirn.y, dirn.z);
        return;
    }
    s.f = hit_frame(sc, o, d, h);
    const pgrt_material& mat = sc.materials[sc.geom_material[s.f.geom]];
    if (p.shader_mode == 2) {                                           // :274-280
        s.color = make_float4((s.f.n.x + 1) / 2, (s.f.n.y + 1) / 2, (s.f.n.z + 1) / 2, 1.0f);
    } else if (level >= p.max_depth) {                                  // :282-283
        s.color = make_float4(0.f, 0.f, 0.f, 1.f);
    } else if (mat.type == 4 && (p.shader_mode == 0 || (PATH && p.shader_mode == 3))) {   // :294-323
        float n1, n2;
        if (d.w == PGRT_IOR_AIR) { n1 = PGRT_IOR_AIR; n2 = mat.ior; } else { n1 = mat.ior; n2 = PGRT_IOR_AIR; }   // :261-267
        s.kind = SK_DIEL;
        s.refl = make_reflection_ray(s.f.dirn, s.f.n, s.f.hitp, n1);
        s.refr = make_refraction_ray(s.f.dirn, s.f.n, n1, n2, s.f.hitp);
        s.has_refr = (s.refr.d.x == s.refr.d.x);                        // :309
        const V3 v = -s.f.dirn;
        s.att = diel_attenuation(mat.diffuse[0], mat.diffuse[1], mat.diffuse[2], h.x, n1, n2, fabsf(dot3(s.f.n, v)), s.has_refr);
    } else if (PATH && p.shader_mode == 3) {
        // path tracing: the Phong value of this hit plus albedo x the radiance of ONE cosine-weighted bounce; handled as
        // a node with a single child whose combine is linear (combine_node, att.w < 0)
        s.kind = SK_PATH;
        float m_d_r, m_d_g, m_d_b;
        diffuse_albedo(sc, mat, s.f, m_d_r, m_d_g, m_d_b);
        s.att = make_float4(m_d_b, m_d_g, m_d_r, -1.0f);                // phong_eval returns {blue, green, red}
        s.refl.o = s.f.hitp; s.refl.tnear = 0.01f; s.refl.time = d.w;   // stays in the medium it came through
        s.refl.d = path_bounce_direction(s.f.n, s.f.hitp, p.seed, level);
    } else {
        s.kind = SK_PHONG;
    }
}

// ---- K8b/K9 (per ray): Phong sum with the shadow query inline (raytracer.cpp:325-386, is_illuminated :150-176,
//      LightSource::GenerateRay LightSource.cpp:11-32)
struct TravAcc { unsigned long long nodes, tris; uint32_t mx; };

// closest hit as its own function: one copy of the traversal loop per node layout whatever the number of call sites,
// and the register allocation of that loop is not mixed with the shading code around it
#ifdef PGRT_TRACE_INLINE
#define PGRT_TRACE_DEV_ATTR __forceinline__
#else
#define PGRT_TRACE_DEV_ATTR __noinline__
#endif
template <bool COUNT>
__device__ PGRT_TRACE_DEV_ATTR HitRec trace_dev_walk(const DevScene& sc, V3 O, V3 D, float tnear, float tfar, TravCount& tc) {
    return trace_closest_t<COUNT>(sc, O, D, tnear, tfar, tc);
}
// the query every kernel of this file issues: the scene-box test inline (traverse.cuh), the walk only for rays that pass it
template <bool COUNT>
__device__ __forceinline__ HitRec trace_dev(const DevScene& sc, V3 O, V3 D, float tnear, float tfar, TravCount& tc) {
    if (ray_misses_box(sc.bb_lo, sc.bb_hi, O, D, tnear, tfar)) {
        HitRec miss; miss.t = tfar; miss.u = 0.0f; miss.v = 0.0f; miss.tri = PGRT_INVALID_ID;
        tc.nodes = 0; tc.tris = 0;
        return miss;
    }
    return trace_dev_walk<COUNT>(sc, O, D, tnear, tfar, tc);
}


// Raytracer::is_illuminated (raytracer.cpp:150-176) with LightSource::GenerateRay (LightSource.cpp:11-32)
template <bool COUNT>
__device__ __forceinline__ bool is_illuminated_dev(const DevScene& sc, const pgrt_render_params& p, V3 lp, V3 hitp, V3 n, unsigned long long& my_shadow, TravAcc& acc) {
    bool lit = false;
    if (p.shadow_mode == 1) {
        // README "To do: hard shadows" (README.md:20), behind a non-default flag: the ray the reference meant to
        // cast -- from the hit point towards the light, in units of that segment -- with the same rule that the
        // closest occluder decides and a dielectric one does not shadow.  Not part of the parity contract.
        const V3 to_light = v3(lp.x - hitp.x, lp.y - hitp.y, lp.z - hitp.z);
        if (!(dot3(n, to_light) < 0)) {
            my_shadow++;
            TravCount tc; tc.nodes = 0; tc.tris = 0;
            const HitRec sh = trace_dev<COUNT>(sc, hitp, to_light, 1e-3f, 1.0f, tc);
            if (COUNT) { acc.nodes += tc.nodes; acc.tris += tc.tris; acc.mx = max(acc.mx, tc.nodes); }
            if (sh.tri == PGRT_INVALID_ID) lit = true;
            else {
                const uint32_t g = __float_as_uint(__ldg(sc.shade + 4 * (size_t)sh.tri + 3).w);
                lit = sc.materials[sc.geom_material[g]].type == 4;
            }
        }
    } else if (!(dot3(n, lp) < 0)) {                                        // :155
        // the shadow ray leaves the light with the hit POSITION as its direction (sic, LightSource.cpp:18-20)
        const float tfar = l2norm3(v3(lp.x - hitp.x, lp.y - hitp.y, lp.z - hitp.z));
        my_shadow++;
        TravCount tc; tc.nodes = 0; tc.tris = 0;
        const HitRec sh = trace_dev<COUNT>(sc, lp, hitp, 0.01f, tfar, tc);
        if (COUNT) { acc.nodes += tc.nodes; acc.tris += tc.tris; acc.mx = max(acc.mx, tc.nodes); }
        if (sh.tri == PGRT_INVALID_ID) lit = true;
        else {                                                              // :166-172: a dielectric occluder does not shadow
            const uint32_t g = __float_as_uint(__ldg(sc.shade + 4 * (size_t)sh.tri + 3).w);
            lit = sc.materials[sc.geom_material[g]].type == 4;
        }
    }
    return lit;
}

template <bool COUNT>
__device__ __forceinline__ float4 phong_eval(const DevScene& sc, const pgrt_render_params& p, float4 o, float4 d, const HitFrame& f,
                                             unsigned long long& my_shadow, TravAcc& acc) {
    const pgrt_material& mat = sc.materials[sc.geom_material[f.geom]];
    float blue = 0, green = 0, red = 0;
    float m_d_r, m_d_g, m_d_b;
    diffuse_albedo(sc, mat, f, m_d_r, m_d_g, m_d_b);
    for (int li = 0; li < sc.n_lights; ++li) {                              // :351
        const pgrt_light& light = sc.lights[li];
        const V3 lp = v3(light.position[0], light.position[1], light.position[2]);
        const bool lit = is_illuminated_dev<COUNT>(sc, p, lp, f.hitp, f.n, my_shadow, acc);
        if (lit) {
            const V3 light_vector = normalize3(v3(lp.x - f.hitp.x, lp.y - f.hitp.y, lp.z - f.hitp.z));
            const V3 camera_vector = normalize3(v3(o.x - d.x, o.y - d.y, o.z - d.z));   // :360 (sic)
            const float i_d_r = light.diffuse[2], i_d_g = light.diffuse[1], i_d_b = light.diffuse[0];
            const float i_s_r = light.specular[2], i_s_g = light.specular[1], i_s_b = light.specular[0];
            const float m_s_r = mat.specular[2], m_s_g = mat.specular[1], m_s_b = mat.specular[0];
            const float ndl = dot3(f.n, light_vector);
            const V3 l_r = normalize3(2 * (ndl)*f.n - light_vector);
            if (p.shader_mode == 1) {
                blue += i_d_b * m_d_b * ndl; green += i_d_g * m_d_g * ndl; red += i_d_r * m_d_r * ndl;
            } else {
                const float spec = f_powf_whole(dot3(camera_vector, l_r), mat.shininess);
                blue += (i_d_b * m_d_b * ndl + i_s_b * m_s_b * spec);
                green += (i_d_g * m_d_g * ndl + i_s_g * m_s_g * spec);
                red += (i_d_r * m_d_r * ndl + i_s_r * m_s_r * spec);
            }
        }
    }
    return make_float4(blue, green, red, 1.0f);                            // :385
}

// ---- K10 (per node): value of a dielectric node from its children (raytracer.cpp:318-321)
//      `own`: what the node's colour slot held before (SK_PATH keeps its Phong value there; unused otherwise)
template <bool PATH>
__device__ PGRT_COLD float4 combine_node(float4 att, float4 a, bool has_b, float4 b, float4 own) {
    if (PATH && att.w < 0.0f) return make_float4(own.x + att.x * a.x, own.y + att.y * a.y, own.z + att.z * a.z, 1.0f);   // SK_PATH
    if (has_b) {
        Col4 c0, c1; c0.r = a.x; c0.g = a.y; c0.b = a.z; c0.a = a.w; c1.r = b.x; c1.g = b.y; c1.b = b.z; c1.a = b.w;
        const Col4 c = mix_srgb(c0, c1, att.w);
        return make_float4(c.r * att.x, c.g * att.y, c.b * att.z, 1.0f);
    }
    return make_float4(a.x * att.x, a.y * att.y, a.z * att.z, 1.0f);
}

// ---- K11 (per pixel): mean over the samples + gamma (raytracer.cpp:421-446), and the 8-bit form the reference displays
__device__ __forceinline__ float4 resolve_value(float fr, float fg, float fb, int S, float gamma_level) {
    Col4 in; in.r = fb / S; in.g = fg / S; in.b = fr / S; in.a = 1.0f;       // :431 (swap in)
    const Col4 g = gamma_correct(in, gamma_level);
    return make_float4(g.r, g.g, g.b, g.a);
}
// float -> R8G8B8A8_UNORM as D3D11 converts when the float texture reaches the back buffer (simpleguidx11.cpp:229,290):
// NaN -> 0, clamp to [0,1], scale by 255, round to nearest
__device__ __forceinline__ uint32_t unorm8(float c) {
    c = (c != c) ? 0.0f : fminf(fmaxf(c, 0.0f), 1.0f);
    return (uint32_t)floorf(c * 255.0f + 0.5f);
}
__device__ __forceinline__ uint32_t pack_rgba8(float4 c) { return unorm8(c.x) | (unorm8(c.y) << 8) | (unorm8(c.z) << 16) | (unorm8(c.w) << 24); }
__device__ __forceinline__ void store_pixel(const FrameOut& fo, size_t idx, float4 c) {
    if (fo.rgba) fo.rgba[idx] = c;
    if (fo.rgba8) fo.rgba8[idx] = pack_rgba8(c);
}

// a finished level-0 sample IS the pixel when there is one sample per pixel: gamma and the store at its final place
__device__ __forceinline__ void store_sample_as_pixel(const FrameOut& fo, const Gen0& g0, const pgrt_render_params& p, uint32_t i, float4 col) {
    int x, y;
    const uint32_t slot = g0.slot0 + i;
    if (slot_to_pixel(g0.sh, g0.cam.width, g0.cam.height, slot, x, y))
        store_pixel(fo, fo.compact ? (size_t)slot : (size_t)y * g0.cam.width + x, resolve_value(0.0f + col.x, 0.0f + col.y, 0.0f + col.z, 1, p.gamma_level));
    else if (fo.compact) store_pixel(fo, (size_t)slot, make_float4(0.f, 0.f, 0.f, 0.f));
}

// ---- K8 (kernel): one level of the wavefront (level-synchronous scheduler)
// POOL (hybrid scheduler, level 0 only): the children of a dielectric hit do not go to the next level's queue but to the frame's
// ray pool, as records of k_frame's own format (published by their epoch word, counted in `outstanding`), and the node waits
// in L for the continuation: the k_frame launched behind this kernel finds all of level 1 ready and runs the rest of trace().
template <bool PATH, bool POOL>
__global__ void __launch_bounds__(256, PGRT_SHADE_MIN_BLOCKS) k_shade(DevScene sc, pgrt_render_params p, Gen0 g0, int level, LevelBufs L, LevelBufs Ln, RayPool P, FrameOut fo, Counters* cnt) {
    const uint32_t n = min(cnt->n_rays[level], L.cap);
    const uint32_t epoch = POOL ? cnt->epoch : 0u;
    const int lane = threadIdx.x & 31;
    const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long my_refl = 0, my_refr = 0;
    for (uint32_t base = warp_id * 32u; base < n; base += n_warps * 32u) {
        const uint32_t i = base + lane;
        ShadeOut s; s.kind = SK_FINAL; s.has_refr = false;
        bool is_phong = false, is_diel = false;
        if (i < n) {
            float4 o, d;
            load_ray_shade(L, g0, p, i, o, d);
            shade_classify<PATH>(sc, p, level, o, d, L.hit[i], s);
            is_phong = s.kind == SK_PHONG || (PATH && s.kind == SK_PATH); is_diel = s.kind == SK_DIEL || (PATH && s.kind == SK_PATH);
            // (POOL with one sample per pixel, fo.direct: a finished sample goes straight to the frame, no colour queue, no k_resolve)
            if (s.kind == SK_FINAL) { if (POOL && fo.direct) store_sample_as_pixel(fo, g0, p, i, s.color); else L.color[i] = s.color; }
        }
        const bool has_refr = is_diel && s.has_refr;
        const uint32_t pslot = warp_append(&cnt->n_phong[level], is_phong, lane);
        if (is_phong) L.phong_list[pslot] = i;
        if (POOL) {
            const uint32_t rl = warp_append(&cnt->q_tail, is_diel, lane);
            const uint32_t rr = warp_append(&cnt->q_tail, has_refr, lane);
            int pub_w = (is_diel && rl < P.cap ? 1 : 0) + (has_refr && rr < P.cap ? 1 : 0);
            for (int o2 = 16; o2 > 0; o2 >>= 1) pub_w += __shfl_xor_sync(0xffffffffu, pub_w, o2);
            if (pub_w > 0 && lane == 0) atomicAdd(&cnt->outstanding, (uint32_t)pub_w);
            if (is_diel) {
                if (rl < P.cap && (!has_refr || rr < P.cap)) {
                    const uint32_t meta = (uint32_t)(level + 1) | (1u << 9);
                    L.dn_att[i] = s.att; L.dn_child[i] = make_uint2(rl, has_refr ? rr : PGRT_INVALID_ID); L.pending[i] = has_refr ? 2u : 1u;
                    P.ray_o[rl] = make_float4(s.refl.o.x, s.refl.o.y, s.refl.o.z, s.refl.tnear);
                    P.ray_d[rl] = make_float4(s.refl.d.x, s.refl.d.y, s.refl.d.z, s.refl.time);
                    P.link[rl] = make_uint4(i, meta, epoch, 0u);
                    my_refl++;
                    if (has_refr) {
                        P.ray_o[rr] = make_float4(s.refr.o.x, s.refr.o.y, s.refr.o.z, s.refr.tnear);
                        P.ray_d[rr] = make_float4(s.refr.d.x, s.refr.d.y, s.refr.d.z, s.refr.time);
                        P.link[rr] = make_uint4(i, meta | (1u << 8), epoch, 0u);
                        my_refr++;
                    }
                } else {
                    cnt->overflow = 1u;   // black; the frame is re-rendered with a larger pool.  A reserved record inside the pool is published as dead
                    if (rl < P.cap) { P.ray_d[rl] = make_float4(0.f, 0.f, 0.f, -1.0f); P.link[rl] = make_uint4(0u, 0u, epoch, 0u); }
                    if (has_refr && rr < P.cap) { P.ray_d[rr] = make_float4(0.f, 0.f, 0.f, -1.0f); P.link[rr] = make_uint4(0u, 0u, epoch, 0u); }
                    if (fo.direct) store_sample_as_pixel(fo, g0, p, i, make_float4(0.f, 0.f, 0.f, 1.f)); else L.color[i] = make_float4(0.f, 0.f, 0.f, 1.f);
                }
            }
            continue;
        }
        const uint32_t dslot = warp_append(&cnt->n_diel[level], is_diel, lane);
        const uint32_t rl = warp_append(&cnt->n_rays[level + 1], is_diel, lane);
        const uint32_t rr = warp_append(&cnt->n_rays[level + 1], has_refr, lane);
        if (is_diel) {
            L.diel_list[dslot] = i;
            uint2 ch = make_uint2(PGRT_INVALID_ID, PGRT_INVALID_ID);
            if (rl < Ln.cap && (!has_refr || rr < Ln.cap)) {
                ch.x = rl;
                Ln.ray_o[rl] = make_float4(s.refl.o.x, s.refl.o.y, s.refl.o.z, s.refl.tnear);
                Ln.ray_d[rl] = make_float4(s.refl.d.x, s.refl.d.y, s.refl.d.z, s.refl.time);
                my_refl++;
                if (has_refr) {
                    ch.y = rr;
                    Ln.ray_o[rr] = make_float4(s.refr.o.x, s.refr.o.y, s.refr.o.z, s.refr.tnear);
                    Ln.ray_d[rr] = make_float4(s.refr.d.x, s.refr.d.y, s.refr.d.z, s.refr.time);
                    my_refr++;
                }
            } else {
                cnt->overflow = 1u;   // the frame is re-rendered with larger queues
            }
            L.dn_att[i] = s.att;
            L.dn_child[i] = ch;
        }
    }
    for (int o = 16; o > 0; o >>= 1) { my_refl += __shfl_xor_sync(0xffffffffu, my_refl, o); my_refr += __shfl_xor_sync(0xffffffffu, my_refr, o); }
    if (lane == 0) { if (my_refl) atomicAdd(&cnt->reflection, my_refl); if (my_refr) atomicAdd(&cnt->refraction, my_refr); }
}

// ---- K8b/K9 (kernel)
template <bool COUNT>
__global__ void __launch_bounds__(128, PGRT_PHONG_MIN_BLOCKS) k_phong(DevScene sc, pgrt_render_params p, Gen0 g0, int level, LevelBufs L, FrameOut fo, Counters* cnt) {
    const uint32_t n = cnt->n_phong[level];
    unsigned long long my_shadow = 0;
    TravAcc acc; acc.nodes = 0; acc.tris = 0; acc.mx = 0;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t i = L.phong_list[k];
        float4 o, d;
        load_ray_shade(L, g0, p, i, o, d);
        const float4 h = L.hit[i];
        const HitFrame f = hit_frame(sc, o, d, h);
        const float4 col = phong_eval<COUNT>(sc, p, o, d, f, my_shadow, acc);
        if (fo.direct) store_sample_as_pixel(fo, g0, p, i, col); else L.color[i] = col;     // fo.direct: hybrid scheduler, one sample per pixel
    }
    for (int o = 16; o > 0; o >>= 1) my_shadow += __shfl_xor_sync(0xffffffffu, my_shadow, o);
    if ((threadIdx.x & 31) == 0 && my_shadow) { atomicAdd(&cnt->shadow, my_shadow); atomicAdd(&cnt->lv_shadow[level], my_shadow); }
    if (COUNT) flush_trav_counts(acc.nodes, acc.tris, acc.mx, &cnt->lv_sh_nodes[level], &cnt->lv_sh_tris[level], &cnt->lv_sh_max_nodes[level]);
}

// ---- K10 (kernel): post-order combine of one level's dielectric nodes (level-synchronous scheduler)
template <bool PATH>
__global__ void __launch_bounds__(256) k_combine(int level, LevelBufs L, LevelBufs Ln, const Counters* __restrict__ cnt) {
    const uint32_t n = cnt->n_diel[level];
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t i = L.diel_list[k];
        const float4 att = L.dn_att[i];
        const uint2 ch = L.dn_child[i];
        float4 out = make_float4(0.f, 0.f, 0.f, 1.f);
        if (ch.x != PGRT_INVALID_ID) {
            const bool has_b = ch.y != PGRT_INVALID_ID;
            out = combine_node<PATH>(att, Ln.color[ch.x], has_b, has_b ? Ln.color[ch.y] : make_float4(0.f, 0.f, 0.f, 0.f), L.color[i]);
        } else if (PATH && att.w < 0.0f) {
            out = L.color[i];             // SK_PATH whose bounce did not fit the queue (overflow: the frame is redone)
        }
        L.color[i] = out;
    }
}

// ---- fused scheduler: the whole of trace() for every sample of a batch in one persistent kernel (see the file header)
#ifndef PGRT_FRAME_THREADS
#define PGRT_FRAME_THREADS 128      // threads per CTA of k_frame (warps work on their own: the CTA is only the unit that holds an SM slot)
#endif
#ifndef PGRT_FRAME_MIN_BLOCKS
#define PGRT_FRAME_MIN_BLOCKS (5 * 128 / PGRT_FRAME_THREADS)     // register cap of k_frame = 65536 / (threads * this)
#endif

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { return *(const volatile uint32_t*)p; }
// a load the compiler can neither hoist out of a spin loop nor drop with the loop (a loop without side effects may be assumed to end)
__device__ __forceinline__ uint4 ld_volatile_u4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long global_timer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// keep_ctas: CTAs 0 .. keep_ctas-1 stay until the batch is done (they poll for work while rays are outstanding); the others
// leave as soon as they find neither a primary chunk nor a pool record, which frees their SM slots for the next frame's kernel
// while this frame's dependent chains (up to max_depth traversals in a row) are still running.
template <bool COUNT, bool PATH>
#ifdef PGRT_GRID_CONST
#define PGRT_GC const __grid_constant__
#else
#define PGRT_GC
#endif
__global__ void __launch_bounds__(PGRT_FRAME_THREADS, PGRT_FRAME_MIN_BLOCKS) k_frame(PGRT_GC DevScene sc, PGRT_GC pgrt_render_params p, PGRT_GC Gen0 g0, LevelBufs L0, RayPool P, FrameOut fo,
                                                                       int min_claim, int patience, int keep_ctas, int policy, Counters* cnt) {
    const int lane = threadIdx.x & 31;
    const uint32_t n0 = g0.n_slots * (uint32_t)g0.spp;             // primary samples of this batch
    const uint32_t epoch = cnt->epoch;
    const bool keeper = (int)blockIdx.x < keep_ctas;
    unsigned long long my_shadow = 0, my_refl = 0, my_refr = 0, my_shadow0 = 0;
    unsigned long long my_nodes0 = 0, my_tris0 = 0; uint32_t my_max0 = 0;   // COUNT: level-0 traversal statistics
    bool more_primary = true;
#ifdef PGRT_FRAME_TIMING
    const long long tm_enter = clock64(); long long tm_prim = 0, tm_sec = 0;
#endif
#ifdef PGRT_NO_PREFETCH
#define PGRT_PF_ON 0
#else
#define PGRT_PF_ON 1
#endif
    // Pool records are handed out by TICKET: atomicAdd(q_head, 32) never fails, so no warp ever retries against the others (a
    // compare-and-swap here makes hundreds of idle warps fight for one word and lets ONE of them win 32 rays per round).  A
    // ticket may run ahead of q_tail: its holder then waits for exactly its own records to be published -- each lane polls its
    // own record's publication word, 32 distinct addresses -- so idle warps queue up for the rays that are still to be
    // produced, in allocation order, and the children of a finished ray are picked up within a poll interval.
    uint32_t tk_base = 0, tk_mask = 0;          // the held ticket: records tk_base + lane, for the lanes of tk_mask that are still to be processed
    uint32_t wait_polls = 0;                    // polls since the held ticket last made progress
    uint32_t pf_base = 0xFFFFFFFFu; uint4 pf_q = make_uint4(0u, 0u, 0u, 0u);   // lane 0: the next primary chunk and a snapshot of the queue words
    if (lane == 0) {
        atomicMin(&cnt->t_first, global_timer_ns());
        pf_base = atomicAdd(&cnt->trace_next[0], 32u);
        pf_q = ld_volatile_u4(reinterpret_cast<const uint4*>(&cnt->q_tail));
    }
    for (;;) {
        uint32_t base = 0, n = 0, take = 0; int mode = 0;     // mode 1 = pool records (lanes of `take`), 2 = primary rays
        uint4 l4 = make_uint4(0u, 0u, 0u, 0u);
        // ---- a held ticket first: which of its records have been published?
        if (tk_mask) {
            const uint32_t rec = tk_base + (uint32_t)lane;
            const bool mine = (tk_mask >> lane) & 1u;
            bool ready = false;
            if (mine && rec < P.cap) { l4 = ld_volatile_u4(&P.link[rec]); ready = l4.z == epoch; }
            tk_mask &= ~__ballot_sync(0xffffffffu, mine && rec >= P.cap);          // beyond the pool: such a record never comes
            const uint32_t ready_mask = __ballot_sync(0xffffffffu, ready);
            // all of it, or -- once the rest has been waited for long enough -- what there is
            if (ready_mask && (ready_mask == tk_mask || wait_polls >= (uint32_t)(more_primary ? 4 * patience : patience))) { mode = 1; take = ready_mask; }
        }
        // ---- else new work: a full ticket's worth of waiting records, or a chunk of primary rays.  While primary rays last, the
        //      queue words and the next chunk were fetched when the previous chunk STARTED (pf_q, pf_base: lane 0), so neither
        //      round trip to L2 is waited for here; a snapshot that is one chunk old only makes a ticket run ahead or come late.
        if (mode == 0 && (more_primary || tk_mask == 0u)) {
            int m = 0;      // 2 primary chunk, 5 took a ticket, 6 nothing left anywhere, 7 primaries just ran out
            if (lane == 0) {
                if (more_primary) {
                    // While primary rays last, `policy` (PGRT_POOL_POLICY) says whether a warp looks at the pool at all.  0 (default): no
                    // -- the secondary rays wait until the primary rays of their frame are gone and are then worked off with full
                    // warps, beside the NEXT frame's primary rays when frames are pipelined: C2 0.461 ms per frame.  2: by ticket,
                    // like the idle warps later -- the chains advance all through the primary phase, but every traversal of
                    // incoherent rays holds a warp slot five to six times as long as a primary chunk: 0.592 ms.  1: a full batch
                    // that exists right now, by one compare-and-swap on a fresh read (no retry): 0.481 ms.
                    // (profiles/r2_sweep_kframe_policy.txt; the single frame takes 1.06 - 1.12 ms with all three.)
                    if (tk_mask == 0u && policy != 0 && (int)(min(pf_q.x, P.cap) - pf_q.y) >= min_claim) {
                        if (policy == 2) { base = atomicAdd(&cnt->q_head, 32u); m = 5; }
                        else {      // policy 1: only a full batch that exists right now, by one compare-and-swap on a fresh read
                            const uint4 q = ld_volatile_u4(reinterpret_cast<const uint4*>(&cnt->q_tail));
                            if ((int)(min(q.x, P.cap) - q.y) >= 32 && atomicCAS(&cnt->q_head, q.y, q.y + 32u) == q.y) { base = q.y; m = 5; }
                            pf_q = q;
                        }
                    }
                    if (m == 0) {
                        base = pf_base;
                        if (base < n0) { m = 2; n = min(32u, n0 - base); }
                        else { m = 7; atomicMax(&cnt->t_primary_done, global_timer_ns()); }
                    }
                } else {
                    // Afterwards idle warps queue up by ticket (see above): atomicAdd never fails, whoever is first in line gets the
                    // next records that are produced.
                    const uint4 q = ld_volatile_u4(reinterpret_cast<const uint4*>(&cnt->q_tail));
                    const int avail = (int)(min(q.x, P.cap) - q.y);                              // negative: tickets already run ahead of the records
                    if (tk_mask == 0u && q.z != 0u && (keeper || avail > 0)) { base = atomicAdd(&cnt->q_head, 32u); m = 5; }
                    else if (q.z == 0u || tk_mask == 0u) m = 6;     // the batch is done, or this warp may leave (not a keeper, nothing waiting)
                }
            }
            m = __shfl_sync(0xffffffffu, m, 0); base = __shfl_sync(0xffffffffu, base, 0); n = __shfl_sync(0xffffffffu, n, 0);
            if (m == 5) { tk_base = base; tk_mask = 0xffffffffu; wait_polls = 0; continue; }
            if (m == 7) { more_primary = false; continue; }
            if (m == 6) break;
            if (m == 2) {
                mode = 2;
                if (lane == 0) {      // for the NEXT iteration; the answers arrive while this chunk is traced
                    pf_base = atomicAdd(&cnt->trace_next[0], 32u);
                    pf_q = ld_volatile_u4(reinterpret_cast<const uint4*>(&cnt->q_tail));
                }
            }
        }
        if (mode == 0) {
            // holding a ticket whose records are not there yet, nothing else to do: wait (bounded), unless the batch is over
            if ((wait_polls & 3u) == 0u) {                                 // (the waiters all read this one word: not at every poll)
                uint32_t out = 1u;
                if (lane == 0) out = ld_volatile_u32(&cnt->outstanding);
                out = __shfl_sync(0xffffffffu, out, 0);
                if (out == 0u) break;                                      // nothing is in flight any more: the rest of the ticket never comes
            }
            if (++wait_polls > (1u << 23)) {                               // seconds of waiting: never expected; reported, not hung
                if (lane == 0 && atomicExch(&cnt->watchdog, 1u) == 0u) {
                    const uint4 q = ld_volatile_u4(reinterpret_cast<const uint4*>(&cnt->q_tail));
                    cnt->wd_info[0] = tk_base; cnt->wd_info[1] = tk_mask; cnt->wd_info[2] = q.x; cnt->wd_info[3] = q.y; cnt->wd_info[4] = q.z;
                    cnt->wd_info[5] = blockIdx.x | (more_primary ? 0x80000000u : 0u); cnt->wd_info[6] = ld_volatile_u32(&cnt->trace_next[0]); cnt->wd_info[7] = P.cap;
                }
                break;
            }
            __nanosleep(200u + 25u * min(wait_polls, 32u));
            continue;
        }
        const bool l0 = mode == 2;
#ifdef PGRT_FRAME_TIMING
        const long long tm_0 = clock64();
#endif

        // ---- this lane's ray
        uint32_t i = PGRT_INVALID_ID; int level = 0;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f), d = make_float4(0.f, 0.f, 0.f, -1.0f);
        uint2 lk = make_uint2(0u, 0u);
        int px = 0, py = 0; bool px_valid = false;
        if (l0) {
            if ((uint32_t)lane < n) {
                i = base + (uint32_t)lane;
                const uint32_t slot = g0.slot0 + i / (uint32_t)g0.spp;
                px_valid = slot_to_pixel(g0.sh, g0.cam.width, g0.cam.height, slot, px, py);
                if (px_valid) {
                    const RayRec r = primary_ray(g0.cam, p, px, py, (int)(i % (uint32_t)g0.spp));
                    o = make_float4(r.o.x, r.o.y, r.o.z, r.tnear); d = make_float4(r.d.x, r.d.y, r.d.z, r.time);
                }
            }
            if (tk_mask) wait_polls++;                                     // time passes for a ticket that waits meanwhile
        } else {
            n = (uint32_t)__popc(take);
            tk_mask &= ~take; wait_polls = 0;
            if (lane == 0) { atomicAdd(&cnt->pool_iters, 1u); if (more_primary) pf_q = ld_volatile_u4(reinterpret_cast<const uint4*>(&cnt->q_tail)); }
            if (COUNT && ((take >> lane) & 1u)) atomicMin(&cnt->lv_t_first[l4.y & 0xFFu], global_timer_ns());
            if ((take >> lane) & 1u) {
                i = tk_base + (uint32_t)lane;
                __threadfence();                                           // the publication word was seen: now the record itself
                o = __ldcg(&P.ray_o[i]); d = __ldcg(&P.ray_d[i]);
                lk = make_uint2(l4.x, l4.y); level = (int)(l4.y & 0xFFu);
            }
        }
        const bool live = i != PGRT_INVALID_ID && d.w >= 0.0f;

        // ---- trace() for this ray: closest hit, classification, Phong (with its shadow query)
        ShadeOut s; s.kind = SK_FINAL; s.has_refr = false;
        bool final_ = false;
        float4 col = make_float4(0.f, 0.f, 0.f, 1.f);
        uint32_t lane_shadow = 0;                                        // shadow queries this lane's ray issued (levels >= 1: counted per level)
        if (live) {
            TravCount tc; tc.nodes = 0; tc.tris = 0;
            const HitRec hr = trace_dev<COUNT>(sc, v3(o.x, o.y, o.z), v3(d.x, d.y, d.z), o.w, FLT_MAX, tc);
            if (COUNT) {
                if (l0) { my_nodes0 += tc.nodes; my_tris0 += tc.tris; my_max0 = max(my_max0, tc.nodes); }
                else {
                    atomicAdd(&cnt->lv_nodes[level], (unsigned long long)tc.nodes);
                    atomicAdd(&cnt->lv_tris[level], (unsigned long long)tc.tris); atomicMax(&cnt->lv_max_nodes[level], tc.nodes);
                }
            }
            shade_classify<PATH>(sc, p, level, o, d, make_float4(hr.t, hr.u, hr.v, __uint_as_float(hr.tri)), s);
            if (s.kind == SK_FINAL) { col = s.color; final_ = true; }
            else if (s.kind == SK_PHONG || (PATH && s.kind == SK_PATH)) {
                const unsigned long long sh0 = my_shadow;
                TravAcc a1; a1.nodes = 0; a1.tris = 0; a1.mx = 0;
                col = phong_eval<COUNT>(sc, p, o, d, s.f, my_shadow, a1);
                final_ = !PATH || s.kind == SK_PHONG;
                if (PATH && !final_) {                                   // SK_PATH: the node's own value waits in its colour slot for the bounce
                    if (l0) __stcg(&L0.color[i], col); else __stcg(&P.color[i], col);
                }
                if (l0) my_shadow0 += my_shadow - sh0;
                else lane_shadow = (uint32_t)(my_shadow - sh0);
                if (COUNT) {
                    atomicAdd(&cnt->lv_sh_nodes[level], a1.nodes); atomicAdd(&cnt->lv_sh_tris[level], a1.tris);
                    atomicMax(&cnt->lv_sh_max_nodes[level], a1.mx);
                }
            }
        } else if (i != PGRT_INVALID_ID && l0) {
            final_ = true;                                               // unused slot of a partial tile: black, as the wavefront leaves it
        }

        // ---- dielectric hits: two pool records for the children (all 32 lanes take part in the aggregated atomics)
        const bool is_diel = live && (s.kind == SK_DIEL || (PATH && s.kind == SK_PATH));
        bool has_refr = is_diel && s.has_refr;
        const uint32_t rl = warp_append(&cnt->q_tail, is_diel, lane);
        const uint32_t rr = warp_append(&cnt->q_tail, has_refr, lane);
        // records inside the pool will be claimed by somebody, alive or dead: they are outstanding from here on
        const int published = (is_diel && rl < P.cap ? 1 : 0) + (has_refr && rr < P.cap ? 1 : 0);
        // `outstanding` must never read 0 while work exists: the children are counted BEFORE their publication words are stored
        // (counted afterwards, a fast claimer could retire them first, the word would pass through 0 and a waiting keeper would
        // walk away from its ticket -- the records allocated into that ticket later were then never traced and the batch hung).
        // What this iteration retires (one primary chunk, or n pool records) leaves in the same atomic: nothing but this warp's
        // own stores is left of it, and those are ordered before the next kernel by the end of this one.
        int pub_w = published;
        for (int o2 = 16; o2 > 0; o2 >>= 1) pub_w += __shfl_xor_sync(0xffffffffu, pub_w, o2);
        const int retired = l0 ? 1 : (int)n;
        if (pub_w > 0) {
            if (lane == 0) { if (pub_w != retired) atomicAdd(&cnt->outstanding, (uint32_t)(pub_w - retired)); __threadfence(); }
            __syncwarp();
        }
        if (is_diel) {
            if (rl < P.cap && (!has_refr || rr < P.cap)) {
                const uint32_t meta = (uint32_t)(level + 1) | (l0 ? (1u << 9) : 0u);
                // the node first (its children may finish, and climb, on another SM before this warp moves on) ...
                if (l0) { __stcg(&L0.dn_att[i], s.att); __stcg(&L0.dn_child[i], make_uint2(rl, has_refr ? rr : PGRT_INVALID_ID)); __stcg(&L0.pending[i], has_refr ? 2u : 1u); }
                else { __stcg(&P.att[i], s.att); __stcg(&P.child[i], make_uint2(rl, has_refr ? rr : PGRT_INVALID_ID)); __stcg(&P.pending[i], has_refr ? 2u : 1u); }
                __stcg(&P.ray_o[rl], make_float4(s.refl.o.x, s.refl.o.y, s.refl.o.z, s.refl.tnear));
                __stcg(&P.ray_d[rl], make_float4(s.refl.d.x, s.refl.d.y, s.refl.d.z, s.refl.time));
                my_refl++;
                if (has_refr) {
                    __stcg(&P.ray_o[rr], make_float4(s.refr.o.x, s.refr.o.y, s.refr.o.z, s.refr.tnear));
                    __stcg(&P.ray_d[rr], make_float4(s.refr.d.x, s.refr.d.y, s.refr.d.z, s.refr.time));
                    my_refr++;
                }
                __threadfence();
                // ... then the publication words
                __stcg(&P.link[rl], make_uint4(i, meta, epoch, 0u));
                if (has_refr) __stcg(&P.link[rr], make_uint4(i, meta | (1u << 8), epoch, 0u));
            } else {
                cnt->overflow = 1u;   // black; the frame is re-rendered with a larger pool
                // a reserved record inside the pool will be claimed by somebody: publish it as dead
                if (rl < P.cap) { __stcg(&P.ray_d[rl], make_float4(0.f, 0.f, 0.f, -1.0f)); __threadfence(); __stcg(&P.link[rl], make_uint4(0u, 0u, epoch, 0u)); }
                if (has_refr && rr < P.cap) { __stcg(&P.ray_d[rr], make_float4(0.f, 0.f, 0.f, -1.0f)); __threadfence(); __stcg(&P.link[rr], make_uint4(0u, 0u, epoch, 0u)); }
                final_ = true; col = make_float4(0.f, 0.f, 0.f, 1.f);
            }
        }

        // ---- a finished level-0 sample is a pixel (one sample per pixel) or one addend of k_resolve
        if (final_ && l0) {
            if (fo.direct) {
                const uint32_t slot = g0.slot0 + i;
                if (px_valid) store_pixel(fo, fo.compact ? (size_t)slot : (size_t)py * g0.cam.width + px, resolve_value(0.0f + col.x, 0.0f + col.y, 0.0f + col.z, 1, p.gamma_level));
                else if (fo.compact) store_pixel(fo, (size_t)slot, make_float4(0.f, 0.f, 0.f, 0.f));
            } else L0.color[i] = col;
        }
        // ---- continuation: hand the value to the parent; the last child to arrive evaluates the parent and climbs on
        if (final_ && !l0) {
            uint32_t node = i; uint2 nlk = lk; float4 c = col;
            for (;;) {
                __stcg(&P.color[node], c);
                const uint32_t par = nlk.x; const bool par_l0 = (nlk.y >> 9) & 1u;
                __threadfence();
                const uint32_t old = atomicSub(par_l0 ? &L0.pending[par] : &P.pending[par], 1u);
                if (old != 1u) break;
                __threadfence();
                const uint2 ch = par_l0 ? __ldcg(&L0.dn_child[par]) : __ldcg(&P.child[par]);
                const float4 att = par_l0 ? __ldcg(&L0.dn_att[par]) : __ldcg(&P.att[par]);
                const bool has_b = ch.y != PGRT_INVALID_ID;
                const float4 own = PATH && att.w < 0.0f ? (par_l0 ? __ldcg(&L0.color[par]) : __ldcg(&P.color[par])) : make_float4(0.f, 0.f, 0.f, 0.f);
                c = combine_node<PATH>(att, __ldcg(&P.color[ch.x]), has_b, has_b ? __ldcg(&P.color[ch.y]) : make_float4(0.f, 0.f, 0.f, 0.f), own);
                if (par_l0) {
                    if (fo.direct) {
                        int x, y;
                        const uint32_t slot = g0.slot0 + par;
                        slot_to_pixel(g0.sh, g0.cam.width, g0.cam.height, slot, x, y);
                        store_pixel(fo, fo.compact ? (size_t)slot : (size_t)y * g0.cam.width + x, resolve_value(0.0f + c.x, 0.0f + c.y, 0.0f + c.z, 1, p.gamma_level));
                    } else L0.color[par] = c;
                    break;
                }
                const uint4 l4 = __ldcg(&P.link[par]);
                node = par; nlk = make_uint2(l4.x, l4.y);
            }
        }
        if (COUNT && !l0 && i != PGRT_INVALID_ID) atomicMax(&cnt->lv_t_last[level], global_timer_ns());
        if (!l0) {
            // per-level ray counts of the frame: one pair of atomics per warp and level (a warp's records are mostly of one level),
            // not one per ray -- a hundred thousand same-address atomics per frame are a queue of their own in L2
            const int lv = live ? level : -1;
            const unsigned grp = __match_any_sync(0xffffffffu, lv);
            uint32_t tot = 0;                                            // shadow queries of the group (every lane walks its own group's members)
            for (int src = 0; src < 32; ++src) { const uint32_t v = __shfl_sync(0xffffffffu, lane_shadow, src); if ((grp >> src) & 1u) tot += v; }
            if (lv >= 0 && lane == __ffs(grp) - 1) {
                atomicAdd(&cnt->lv_rays[lv], (unsigned long long)__popc(grp));
                if (tot) atomicAdd(&cnt->lv_shadow[lv], (unsigned long long)tot);
            }
        }
        // ---- retire what this iteration processed (one primary chunk, or n pool records), unless that left with the children above
        if (pub_w == 0 && lane == 0) atomicAdd(&cnt->outstanding, (uint32_t)(-retired));
        __syncwarp();
#ifdef PGRT_FRAME_TIMING
        if (l0) tm_prim += clock64() - tm_0; else tm_sec += clock64() - tm_0;
#endif
    }
#ifdef PGRT_FRAME_TIMING
    if (lane == 0) {
        atomicAdd(&cnt->warp_cycles[0], (unsigned long long)tm_prim); atomicAdd(&cnt->warp_cycles[1], (unsigned long long)tm_sec);
        atomicAdd(&cnt->warp_cycles[2], (unsigned long long)(clock64() - tm_enter)); atomicAdd(&cnt->warp_cycles[3], 1ull);
    }
#endif
    if (lane == 0) atomicMax(&cnt->t_last, global_timer_ns());
    for (int o = 16; o > 0; o >>= 1) {
        my_shadow += __shfl_xor_sync(0xffffffffu, my_shadow, o); my_refl += __shfl_xor_sync(0xffffffffu, my_refl, o); my_refr += __shfl_xor_sync(0xffffffffu, my_refr, o);
        my_shadow0 += __shfl_xor_sync(0xffffffffu, my_shadow0, o);
    }
    if (lane == 0) {
        if (my_shadow) atomicAdd(&cnt->shadow, my_shadow);
        if (my_shadow0) atomicAdd(&cnt->lv_shadow[0], my_shadow0);
        if (my_refl) atomicAdd(&cnt->reflection, my_refl);
        if (my_refr) atomicAdd(&cnt->refraction, my_refr);
    }
    if (COUNT) flush_trav_counts(my_nodes0, my_tris0, my_max0, &cnt->lv_nodes[0], &cnt->lv_tris[0], &cnt->lv_max_nodes[0]);
}

// ---- K11: sample resolve (raytracer.cpp:421-436) + gamma (:439-446); full-frame or compact-shard destination
__global__ void __launch_bounds__(256) k_resolve(DevCamera cam, pgrt_render_params p, ShardInfo sh, uint32_t slot0, uint32_t n_slots,
                                                 const float4* __restrict__ color0, FrameOut fo) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    const uint32_t slot = slot0 + s;
    int x, y;
    const bool valid = slot_to_pixel(sh, cam.width, cam.height, slot, x, y);
    if (!valid) { if (fo.compact) store_pixel(fo, slot, make_float4(0.f, 0.f, 0.f, 0.f)); return; }
    const int S = p.sampling_width * p.sampling_width;
    float fr = 0.0f, fg = 0.0f, fb = 0.0f;
    for (int k = 0; k < S; ++k) { const float4 c = color0[(size_t)s * S + k]; fr += c.x; fg += c.y; fb += c.z; }
    store_pixel(fo, fo.compact ? (size_t)slot : (size_t)y * cam.width + x, resolve_value(fr, fg, fb, S, p.gamma_level));
}

// compact per-rank shard buffers (rank-major) -> full frame (simpleguidx11.cpp:108-114 layout)
__global__ void __launch_bounds__(256) k_untile(DevCamera cam, int n_ranks, int tiles_x, int tiles_y, uint32_t slots_per_rank,
                                                const float4* __restrict__ gathered, float4* __restrict__ out) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= slots_per_rank * (uint32_t)n_ranks) return;
    ShardInfo sh; sh.rank = (int)(g / slots_per_rank); sh.n_ranks = n_ranks; sh.tiles_x = tiles_x; sh.tiles_y = tiles_y;
    int x, y;
    if (slot_to_pixel(sh, cam.width, cam.height, g % slots_per_rank, x, y)) out[(size_t)y * cam.width + x] = gathered[g];
}

// primary hit ids of sample 0 of every slot (what raytracer.cpp:247 sees)
__global__ void __launch_bounds__(256) k_primary_ids(DevScene sc, DevCamera cam, ShardInfo sh, uint32_t slot0, uint32_t n_slots, int S,
                                                     const float4* __restrict__ hit0, uint32_t* __restrict__ geom, uint32_t* __restrict__ prim) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    int x, y;
    if (!slot_to_pixel(sh, cam.width, cam.height, slot0 + s, x, y)) return;
    const uint32_t tri = __float_as_uint(hit0[(size_t)s * S].w);
    uint32_t g = PGRT_INVALID_ID, pr = PGRT_INVALID_ID;
    if (tri != PGRT_INVALID_ID) { g = __float_as_uint(__ldg(sc.shade + 4 * (size_t)tri + 3).w); pr = tri - sc.geom_first[g]; }
    geom[(size_t)y * cam.width + x] = g; prim[(size_t)y * cam.width + x] = pr;
}

// frame / batch bookkeeping.  `first`: first batch of a frame (resets the frame totals).
__global__ void k_batch_begin(Counters* c, uint32_t n0, int first, int pool_only) {
    const int t = threadIdx.x;
    if (t <= PGRT_MAX_LEVELS) {
        c->n_rays[t] = t == 0 ? n0 : 0u; c->n_phong[t] = 0; c->n_diel[t] = 0; c->trace_next[t] = 0;
        if (first) {
            c->lv_rays[t] = 0; c->lv_shadow[t] = 0; c->lv_nodes[t] = 0; c->lv_tris[t] = 0; c->lv_sh_nodes[t] = 0; c->lv_sh_tris[t] = 0;
            c->lv_max_nodes[t] = 0; c->lv_sh_max_nodes[t] = 0;
        }
    }
    if (t == 0) {
        c->shadow = 0; c->reflection = 0; c->refraction = 0; c->q_head = 0; c->q_tail = 0;
        c->outstanding = pool_only ? 0u : (n0 + 31u) / 32u;      // primary chunks (none when level 0 runs as wavefront kernels); pool records join as they are published
        if (first) { c->t_first = ~0ull; c->t_primary_done = 0ull; c->t_last = 0ull; c->pool_iters = 0u; }
    }
    if (first && t <= PGRT_MAX_LEVELS) { c->lv_t_first[t] = ~0ull; c->lv_t_last[t] = 0ull; }
    if (t == 0) {
        c->epoch += 1u;      // the pool's publication word: records of earlier batches (and frames) read as "not yet written"
        if (first) { c->warp_cycles[0] = 0; c->warp_cycles[1] = 0; c->warp_cycles[2] = 0; c->warp_cycles[3] = 0; c->overflow = 0; c->watchdog = 0; c->tot_shadow = 0; c->tot_reflection = 0; c->tot_refraction = 0; c->tot_primary = 0; c->q_peak = 0; }
    }
}
// `last`: last batch of the frame.  A frame that finished without a queue overflow bumps the slot's completion count (what
// pgrt_stream_wait_slot waits for) and stores the caller's tag in `done_flag` -- or adds 1 to it -- (host-visible or peer
// memory, pgrt_slot_signal / pgrt_slot_signal_add: what dist.ShardedRenderer waits for); an overflowed attempt does neither,
// so nobody is released before the retry has rendered the frame.
__global__ void k_batch_end(Counters* c, unsigned long long primary, int fused, int last, uint32_t* done_flag, int flag_add) {
    if (!fused && threadIdx.x <= PGRT_MAX_LEVELS && threadIdx.x > 0) c->lv_rays[threadIdx.x] += c->n_rays[threadIdx.x];
    if (threadIdx.x == 0) {
        c->lv_rays[0] += primary; c->tot_shadow += c->shadow; c->tot_reflection += c->reflection; c->tot_refraction += c->refraction; c->tot_primary += primary;
        c->q_peak = max(c->q_peak, c->q_tail);
        if (last && !c->overflow) {
            c->done_seq += 1u;
            // flag_add: the flag is a counter shared by all ranks (one system-scope atomic over NVLink), so that the consumer
            // needs ONE stream wait per frame instead of one per rank
            if (done_flag) { __threadfence_system(); if (flag_add) atomicAdd_system(done_flag, 1u); else *(volatile uint32_t*)done_flag = c->sig_value; }
        }
    }
}

// ---- Raytracer::trace(RTCRay, level) on caller-supplied rays (raytracer.h:31): one thread walks one ray's whole tree
//      depth-first with an explicit stack, in the reference's own order (reflection, then refraction, then the combine).
//      A third evaluation order of the same node functions, besides the two frame schedulers; not a frame path.
struct TraceFrame { float4 att; float4 other; RayRec refr; int stage; };   // other: value of the reflection child, or the node's own value (SK_PATH)
template <bool PATH>
__global__ void __launch_bounds__(128) k_trace_rays(DevScene sc, pgrt_render_params p, const pgrt_ray* __restrict__ rays, uint64_t n, int level0, float4* __restrict__ out,
                                                    TraceFrame* __restrict__ stacks, Counters* cnt) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    TraceFrame* stk = stacks + i * (uint64_t)(PGRT_MAX_LEVELS + 1);
    unsigned long long my_shadow = 0, my_refl = 0, my_refr = 0;
    TravAcc acc; acc.nodes = 0; acc.tris = 0; acc.mx = 0;
    const pgrt_ray q = rays[i];
    float4 o = make_float4(q.org_x, q.org_y, q.org_z, q.tnear), d = make_float4(q.dir_x, q.dir_y, q.dir_z, q.time);
    float tfar = q.tfar;
    int sp = 0, level = level0;
    float4 v = make_float4(0.f, 0.f, 0.f, 1.f);
    for (;;) {
        // ---- descend: evaluate the node of the current ray
        bool have_value;
        {
            TravCount tc;
            const HitRec hr = trace_dev<false>(sc, v3(o.x, o.y, o.z), v3(d.x, d.y, d.z), o.w, tfar, tc);
            ShadeOut s;
            shade_classify<PATH>(sc, p, level, o, d, make_float4(hr.t, hr.u, hr.v, __uint_as_float(hr.tri)), s);
            float4 own = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s.kind == SK_PHONG || (PATH && s.kind == SK_PATH)) own = phong_eval<false>(sc, p, o, d, s.f, my_shadow, acc);
            if (s.kind == SK_FINAL) { v = s.color; have_value = true; }
            else if (s.kind == SK_PHONG) { v = own; have_value = true; }
            else {
                TraceFrame& f = stk[sp++];
                f.att = s.att; f.refr = s.refr; f.other = own;
                f.stage = (PATH && s.kind == SK_PATH) ? 2 : (s.has_refr ? 0 : 1);   // 0: refraction still to come, 1 / 2: one child only
                o = make_float4(s.refl.o.x, s.refl.o.y, s.refl.o.z, s.refl.tnear); d = make_float4(s.refl.d.x, s.refl.d.y, s.refl.d.z, s.refl.time);
                tfar = FLT_MAX; level++; my_refl++;
                have_value = false;
            }
        }
        // ---- climb: hand the value to the parents that are complete
        bool done = false;
        while (have_value) {
            if (sp == 0) { done = true; break; }
            TraceFrame& f = stk[sp - 1];
            if (f.stage == 0) {                       // that was the reflection child: now the refraction ray (raytracer.cpp:308-311)
                f.other = v; f.stage = 3;
                o = make_float4(f.refr.o.x, f.refr.o.y, f.refr.o.z, f.refr.tnear); d = make_float4(f.refr.d.x, f.refr.d.y, f.refr.d.z, f.refr.time);
                tfar = FLT_MAX; my_refr++;            // level stays: both children live one level below the node
                have_value = false;
            } else {
                if (f.stage == 3) v = combine_node<PATH>(f.att, f.other, true, v, make_float4(0.f, 0.f, 0.f, 0.f));
                else v = combine_node<PATH>(f.att, v, false, make_float4(0.f, 0.f, 0.f, 0.f), f.other);
                --sp; --level;
            }
        }
        if (done) break;
    }
    out[i] = v;
    if (my_shadow) atomicAdd(&cnt->shadow, my_shadow);
    if (my_refl) atomicAdd(&cnt->reflection, my_refl);
    if (my_refr) atomicAdd(&cnt->refraction, my_refr);
}

// Raytracer::is_illuminated (raytracer.h:34) over a batch of (light position, hit position, normal)
__global__ void __launch_bounds__(128) k_is_illuminated(DevScene sc, pgrt_render_params p, const float* __restrict__ light, const float* __restrict__ hit,
                                                        const float* __restrict__ nrm, uint64_t n, int32_t* __restrict__ out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long sh = 0; TravAcc acc; acc.nodes = 0; acc.tris = 0; acc.mx = 0;
    out[i] = is_illuminated_dev<false>(sc, p, v3(light[3 * i], light[3 * i + 1], light[3 * i + 2]), v3(hit[3 * i], hit[3 * i + 1], hit[3 * i + 2]),
                                       v3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]), sh, acc) ? 1 : 0;
}

// ---- batch rtcIntersect1 over RTCRayHit-compatible records (device copies)
__global__ void __launch_bounds__(128) k_intersect(DevScene sc, const float* __restrict__ pos, pgrt_rayhit* __restrict__ rh, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        pgrt_rayhit q = rh[i];
        const HitRec h = trace_closest(sc, v3(q.org_x, q.org_y, q.org_z), v3(q.dir_x, q.dir_y, q.dir_z), q.tnear, q.tfar);
        if (h.tri == PGRT_INVALID_ID) continue;   // a miss leaves the record untouched
        const V3 ng = tri_ng(pos, h.tri);
        const uint32_t g = __float_as_uint(__ldg(sc.shade + 4 * (size_t)h.tri + 3).w);
        q.tfar = h.t; q.u = h.u; q.v = h.v; q.Ng_x = ng.x; q.Ng_y = ng.y; q.Ng_z = ng.z;
        q.geomID = g; q.primID = h.tri - sc.geom_first[g]; q.instID = PGRT_INVALID_ID;
        rh[i] = q;
    }
}

__global__ void __launch_bounds__(256) k_interpolate(DevScene sc, const uint32_t* __restrict__ geom, const uint32_t* __restrict__ prim,
                                                     const float* __restrict__ u, const float* __restrict__ v, uint64_t n, int slot, float* __restrict__ out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    V3 nn; float tu, tv; uint32_t g;
    fetch_shading(sc, sc.geom_first[geom[i]] + prim[i], u[i], v[i], nn, tu, tv, g);
    if (slot == 0) { out[3 * i] = nn.x; out[3 * i + 1] = nn.y; out[3 * i + 2] = nn.z; }
    else { out[2 * i] = tu; out[2 * i + 1] = tv; }
}

// ---- per-function evaluation kernels (parity tests of the leaf functions)
__global__ void k_eval_mix(const float4* c0, const float4* c1, const float* alpha, uint64_t n, float4* out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Col4 a, b; a.r = c0[i].x; a.g = c0[i].y; a.b = c0[i].z; a.a = c0[i].w; b.r = c1[i].x; b.g = c1[i].y; b.b = c1[i].z; b.a = c1[i].w;
    const Col4 o = mix_srgb(a, b, alpha[i]);
    out[i] = make_float4(o.r, o.g, o.b, o.a);
}
__global__ void k_eval_texture(DevTexture t, const float2* uv, uint64_t n, float* out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Col3 c = tex_get_texel(t, uv[i].x, uv[i].y);
    out[3 * i] = c.r; out[3 * i + 1] = c.g; out[3 * i + 2] = c.b;
}
__global__ void k_eval_env(DevTexture env, const float* dirs, uint64_t n, float4* out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Col4 c = env_get_texel(env, dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
    out[i] = make_float4(c.r, c.g, c.b, c.a);
}
__global__ void k_eval_gamma(const float4* in, float g, uint64_t n, float4* out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Col4 c; c.r = in[i].x; c.g = in[i].y; c.b = in[i].z; c.a = in[i].w;
    const Col4 o = gamma_correct(c, g);
    out[i] = make_float4(o.r, o.g, o.b, o.a);
}
__global__ void k_eval_primary(DevCamera cam, pgrt_render_params p, float* out) {
    const int S = p.sampling_width * p.sampling_width;
    const uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= (uint64_t)cam.width * cam.height * S) return;
    const int s = (int)(j % S); const uint64_t px = j / S;
    const RayRec r = primary_ray(cam, p, (int)(px % cam.width), (int)(px / cam.width), s);
    float* o = out + 9 * j;
    o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.tnear; o[4] = r.d.x; o[5] = r.d.y; o[6] = r.d.z; o[7] = r.time; o[8] = FLT_MAX;
}
__global__ void k_eval_secondary(const float* in, uint64_t n, int refraction, float* out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* q = in + 11 * i;
    const V3 d = v3(q[0], q[1], q[2]), nn = v3(q[3], q[4], q[5]), hp = v3(q[6], q[7], q[8]);
    const RayRec r = refraction ? make_refraction_ray(d, nn, q[9], q[10], hp) : make_reflection_ray(d, nn, hp, q[9]);
    float* o = out + 9 * i;
    o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.tnear; o[4] = r.d.x; o[5] = r.d.y; o[6] = r.d.z; o[7] = r.time; o[8] = FLT_MAX;
}

// flat-order shading records: 4 x float4 per triangle (normals, uv, geomID)
__global__ void __launch_bounds__(256) k_pack_shade(const float* __restrict__ nrm, const float* __restrict__ uv, const uint32_t* __restrict__ tri_geom,
                                                    uint32_t n, float4* __restrict__ shade) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* a = nrm + 9 * (size_t)i; const float* t = uv + 6 * (size_t)i;
    shade[4 * (size_t)i + 0] = make_float4(a[0], a[1], a[2], t[0]);
    shade[4 * (size_t)i + 1] = make_float4(a[3], a[4], a[5], t[1]);
    shade[4 * (size_t)i + 2] = make_float4(a[6], a[7], a[8], t[2]);
    shade[4 * (size_t)i + 3] = make_float4(t[3], t[4], t[5], __uint_as_float(tri_geom[i]));
}
