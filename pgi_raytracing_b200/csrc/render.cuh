// render.cuh -- wavefront form of Raytracer::get_pixel / trace (pg1/raytracer.cpp:237-437).
//
// The reference recurses: trace(ray, level) calls itself for the reflection and the refraction ray of every
// dielectric hit and combines the two results with the NON-linear mix_srgb (raytracer.cpp:318-319), so no
// linear path throughput exists.  Here every recursion level is a queue:
//
//   forward,  level = 0 .. max_depth :  k_trace -> k_shade -> k_phong
//       k_shade classifies each (ray, hit): miss -> env texel; depth cut-off -> black; dielectric -> pushes its
//       reflection / refraction rays on the level+1 queue (warp-aggregated atomics) and records (attenuation, R,
//       child slots); everything else -> Phong list.  k_phong evaluates the Phong sum with the shadow query inline.
//   backward, level = max_depth-1 .. 0 : k_combine resolves each dielectric node from its two children (post-order).
//   k_resolve averages the samples of a pixel in sx-major order and applies gamma (raytracer.cpp:421-446).
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
    uint32_t cap;
};

struct Counters {
    uint32_t n_rays[PGRT_MAX_LEVELS + 1];
    uint32_t n_phong[PGRT_MAX_LEVELS + 1];
    uint32_t n_diel[PGRT_MAX_LEVELS + 1];
    uint32_t overflow;
    uint32_t pad;
    unsigned long long shadow, reflection, refraction;          // this batch
    unsigned long long tot_shadow, tot_reflection, tot_refraction, tot_primary;   // this frame
    // per-level frame totals; the traversal columns are filled by instrumented renders only (profile bit 1)
    unsigned long long lv_rays[PGRT_MAX_LEVELS + 1], lv_shadow[PGRT_MAX_LEVELS + 1];
    unsigned long long lv_nodes[PGRT_MAX_LEVELS + 1], lv_tris[PGRT_MAX_LEVELS + 1];
    unsigned long long lv_sh_nodes[PGRT_MAX_LEVELS + 1], lv_sh_tris[PGRT_MAX_LEVELS + 1];
    uint32_t lv_max_nodes[PGRT_MAX_LEVELS + 1], lv_sh_max_nodes[PGRT_MAX_LEVELS + 1];
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
        const float l1 = rng_uniform(-p.aperture / 2.0f, p.aperture / 2.0f, rng_u01(p.seed, pixel, (uint32_t)s, 2));
        const float l2 = rng_uniform(-p.aperture / 2.0f, p.aperture / 2.0f, rng_u01(p.seed, pixel, (uint32_t)s, 3));
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

// ---- K7: closest hit for one queue (get_ray_hit, raytracer.cpp:130-148)
template <bool COUNT>
__global__ void __launch_bounds__(128) k_trace(DevScene sc, LevelBufs L, int level, Counters* cnt) {
    const uint32_t n = min(cnt->n_rays[level], L.cap);
    unsigned long long my_nodes = 0, my_tris = 0; uint32_t my_max = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o = L.ray_o[i], d = L.ray_d[i];
        HitRec h; h.t = FLT_MAX; h.u = 0.f; h.v = 0.f; h.tri = PGRT_INVALID_ID;
        if (d.w >= 0.0f) {
            TravCount tc; tc.nodes = 0; tc.tris = 0;
            h = trace_closest_t<COUNT>(sc, v3(o.x, o.y, o.z), v3(d.x, d.y, d.z), o.w, FLT_MAX, tc);
            if (COUNT) { my_nodes += tc.nodes; my_tris += tc.tris; my_max = max(my_max, tc.nodes); }
        }
        L.hit[i] = make_float4(h.t, h.u, h.v, __uint_as_float(h.tri));
    }
    if (COUNT) flush_trav_counts(my_nodes, my_tris, my_max, &cnt->lv_nodes[level], &cnt->lv_tris[level], &cnt->lv_max_nodes[level]);
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

// ---- K8: classification + dielectric expansion + env map (raytracer.cpp:247-323, :390-393)
__global__ void __launch_bounds__(256) k_shade(DevScene sc, pgrt_render_params p, int level, LevelBufs L, LevelBufs Ln, Counters* cnt) {
    const uint32_t n = min(cnt->n_rays[level], L.cap);
    const int lane = threadIdx.x & 31;
    const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long my_refl = 0, my_refr = 0;
    for (uint32_t base = warp_id * 32u; base < n; base += n_warps * 32u) {
        const uint32_t i = base + lane;
        bool is_phong = false, is_diel = false, has_refr = false;
        RayRec refl, refr;
        float4 att = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n) {
            const float4 o = L.ray_o[i], d = L.ray_d[i], h = L.hit[i];
            const uint32_t tri = __float_as_uint(h.w);
            if (d.w < 0.0f) {
                L.color[i] = make_float4(0.f, 0.f, 0.f, 1.f);   // unused slot of a partial tile
            } else if (tri == PGRT_INVALID_ID) {
                const V3 dirn = normalize3(v3(d.x, d.y, d.z));
                const Col4 c = env_get_texel(sc.env, dirn.x, dirn.y, dirn.z);
                L.color[i] = make_float4(c.r, c.g, c.b, c.a);
            } else {
                const HitFrame f = hit_frame(sc, o, d, h);
                const pgrt_material& mat = sc.materials[sc.geom_material[f.geom]];
                if (p.shader_mode == 2) {                                       // :274-280
                    L.color[i] = make_float4((f.n.x + 1) / 2, (f.n.y + 1) / 2, (f.n.z + 1) / 2, 1.0f);
                } else if (level >= p.max_depth) {                              // :282-283
                    L.color[i] = make_float4(0.f, 0.f, 0.f, 1.f);
                } else if (mat.type == 4 && p.shader_mode == 0) {               // :294-323
                    float n1, n2;
                    if (d.w == PGRT_IOR_AIR) { n1 = PGRT_IOR_AIR; n2 = mat.ior; } else { n1 = mat.ior; n2 = PGRT_IOR_AIR; }   // :261-267
                    is_diel = true;
                    refl = make_reflection_ray(f.dirn, f.n, f.hitp, n1);
                    att.x = f_expf(-(1 - mat.diffuse[0]) * h.x);
                    att.y = f_expf(-(1 - mat.diffuse[1]) * h.x);
                    att.z = f_expf(-(1 - mat.diffuse[2]) * h.x);
                    refr = make_refraction_ray(f.dirn, f.n, n1, n2, f.hitp);
                    has_refr = (refr.d.x == refr.d.x);                          // :309
                    if (has_refr) {
                        const V3 v = -f.dirn;
                        const float cos1 = fabsf(dot3(f.n, v));
                        const float alpha = (n1 - n2) / (n1 + n2);
                        att.w = (float)((double)(alpha * alpha + (1 - (alpha * alpha))) * pow((double)(1 - cos1), 5.0));   // :316
                    }
                } else {
                    is_phong = true;
                }
            }
        }
        const uint32_t pslot = warp_append(&cnt->n_phong[level], is_phong, lane);
        if (is_phong) L.phong_list[pslot] = i;
        const uint32_t dslot = warp_append(&cnt->n_diel[level], is_diel, lane);
        const uint32_t rl = warp_append(&cnt->n_rays[level + 1], is_diel, lane);
        const uint32_t rr = warp_append(&cnt->n_rays[level + 1], has_refr, lane);
        if (is_diel) {
            L.diel_list[dslot] = i;
            uint2 ch = make_uint2(PGRT_INVALID_ID, PGRT_INVALID_ID);
            if (rl < Ln.cap && (!has_refr || rr < Ln.cap)) {
                ch.x = rl;
                Ln.ray_o[rl] = make_float4(refl.o.x, refl.o.y, refl.o.z, refl.tnear);
                Ln.ray_d[rl] = make_float4(refl.d.x, refl.d.y, refl.d.z, refl.time);
                my_refl++;
                if (has_refr) {
                    ch.y = rr;
                    Ln.ray_o[rr] = make_float4(refr.o.x, refr.o.y, refr.o.z, refr.tnear);
                    Ln.ray_d[rr] = make_float4(refr.d.x, refr.d.y, refr.d.z, refr.time);
                    my_refr++;
                }
            } else {
                cnt->overflow = 1u;   // the frame is re-rendered in smaller batches
            }
            L.dn_att[i] = att;
            L.dn_child[i] = ch;
        }
    }
    for (int o = 16; o > 0; o >>= 1) { my_refl += __shfl_xor_sync(0xffffffffu, my_refl, o); my_refr += __shfl_xor_sync(0xffffffffu, my_refr, o); }
    if (lane == 0) { if (my_refl) atomicAdd(&cnt->reflection, my_refl); if (my_refr) atomicAdd(&cnt->refraction, my_refr); }
}

// ---- K8b/K9: Phong sum with the shadow query inline (raytracer.cpp:325-386, is_illuminated :150-176,
//      LightSource::GenerateRay LightSource.cpp:11-32)
template <bool COUNT>
__global__ void __launch_bounds__(128) k_phong(DevScene sc, pgrt_render_params p, int level, LevelBufs L, Counters* cnt) {
    const uint32_t n = cnt->n_phong[level];
    unsigned long long my_shadow = 0;
    unsigned long long my_nodes = 0, my_tris = 0; uint32_t my_max = 0;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t i = L.phong_list[k];
        const float4 o = L.ray_o[i], d = L.ray_d[i], h = L.hit[i];
        const HitFrame f = hit_frame(sc, o, d, h);
        const pgrt_material& mat = sc.materials[sc.geom_material[f.geom]];
        float blue = 0, green = 0, red = 0;
        float m_d_r, m_d_g, m_d_b;
        if (mat.diffuse_tex < 0 || mat.diffuse_tex >= sc.n_textures) {         // :338-341
            m_d_r = mat.diffuse[2]; m_d_g = mat.diffuse[1]; m_d_b = mat.diffuse[0];
        } else {                                                                // :343-348
            const Col3 texel = tex_get_texel(sc.textures[mat.diffuse_tex], f.tu, 1.0f - f.tv);
            m_d_r = texel.r; m_d_g = texel.g; m_d_b = texel.b;
        }
        for (int li = 0; li < sc.n_lights; ++li) {                              // :351
            const pgrt_light& light = sc.lights[li];
            const V3 lp = v3(light.position[0], light.position[1], light.position[2]);
            bool lit = false;
            if (!(dot3(f.n, lp) < 0)) {                                         // :155
                // the shadow ray leaves the light with the hit POSITION as its direction (sic, LightSource.cpp:18-20)
                const float tfar = l2norm3(v3(lp.x - f.hitp.x, lp.y - f.hitp.y, lp.z - f.hitp.z));
                my_shadow++;
                TravCount tc; tc.nodes = 0; tc.tris = 0;
                const HitRec sh = trace_closest_t<COUNT>(sc, lp, f.hitp, 0.01f, tfar, tc);
                if (COUNT) { my_nodes += tc.nodes; my_tris += tc.tris; my_max = max(my_max, tc.nodes); }
                if (sh.tri == PGRT_INVALID_ID) lit = true;
                else {                                                          // :166-172: a dielectric occluder does not shadow
                    const uint32_t g = __float_as_uint(__ldg(sc.shade + 4 * (size_t)sh.tri + 3).w);
                    lit = sc.materials[sc.geom_material[g]].type == 4;
                }
            }
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
                    const float spec = f_powf(dot3(camera_vector, l_r), mat.shininess);
                    blue += (i_d_b * m_d_b * ndl + i_s_b * m_s_b * spec);
                    green += (i_d_g * m_d_g * ndl + i_s_g * m_s_g * spec);
                    red += (i_d_r * m_d_r * ndl + i_s_r * m_s_r * spec);
                }
            }
        }
        L.color[i] = make_float4(blue, green, red, 1.0f);                      // :385
    }
    for (int o = 16; o > 0; o >>= 1) my_shadow += __shfl_xor_sync(0xffffffffu, my_shadow, o);
    if ((threadIdx.x & 31) == 0 && my_shadow) { atomicAdd(&cnt->shadow, my_shadow); atomicAdd(&cnt->lv_shadow[level], my_shadow); }
    if (COUNT) flush_trav_counts(my_nodes, my_tris, my_max, &cnt->lv_sh_nodes[level], &cnt->lv_sh_tris[level], &cnt->lv_sh_max_nodes[level]);
}

// ---- K10: post-order combine of one level's dielectric nodes (raytracer.cpp:318-321)
__global__ void __launch_bounds__(256) k_combine(int level, LevelBufs L, LevelBufs Ln, const Counters* __restrict__ cnt) {
    const uint32_t n = cnt->n_diel[level];
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t i = L.diel_list[k];
        const float4 att = L.dn_att[i];
        const uint2 ch = L.dn_child[i];
        float4 out = make_float4(0.f, 0.f, 0.f, 1.f);
        if (ch.x != PGRT_INVALID_ID) {
            const float4 a = Ln.color[ch.x];
            if (ch.y != PGRT_INVALID_ID) {
                const float4 b = Ln.color[ch.y];
                Col4 c0, c1; c0.r = a.x; c0.g = a.y; c0.b = a.z; c0.a = a.w; c1.r = b.x; c1.g = b.y; c1.b = b.z; c1.a = b.w;
                const Col4 c = mix_srgb(c0, c1, att.w);
                out = make_float4(c.r * att.x, c.g * att.y, c.b * att.z, 1.0f);
            } else {
                out = make_float4(a.x * att.x, a.y * att.y, a.z * att.z, 1.0f);
            }
        }
        L.color[i] = out;
    }
}

// ---- K11: sample resolve (raytracer.cpp:421-436) + gamma (:439-446); full-frame or compact-shard destination
__global__ void __launch_bounds__(256) k_resolve(DevCamera cam, pgrt_render_params p, ShardInfo sh, uint32_t slot0, uint32_t n_slots,
                                                 const float4* __restrict__ color0, float4* __restrict__ out, int compact) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    const uint32_t slot = slot0 + s;
    int x, y;
    const bool valid = slot_to_pixel(sh, cam.width, cam.height, slot, x, y);
    if (!valid) { if (compact) out[slot] = make_float4(0.f, 0.f, 0.f, 0.f); return; }
    const int S = p.sampling_width * p.sampling_width;
    float fr = 0.0f, fg = 0.0f, fb = 0.0f;
    for (int k = 0; k < S; ++k) { const float4 c = color0[(size_t)s * S + k]; fr += c.x; fg += c.y; fb += c.z; }
    Col4 in; in.r = fb / S; in.g = fg / S; in.b = fr / S; in.a = 1.0f;
    const Col4 g = gamma_correct(in, p.gamma_level);
    out[compact ? (size_t)slot : (size_t)y * cam.width + x] = make_float4(g.r, g.g, g.b, g.a);
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

__global__ void k_batch_begin(Counters* c) {
    const int t = threadIdx.x;
    if (t <= PGRT_MAX_LEVELS) { c->n_rays[t] = 0; c->n_phong[t] = 0; c->n_diel[t] = 0; }
    if (t == 0) { c->shadow = 0; c->reflection = 0; c->refraction = 0; }
}
__global__ void k_batch_end(Counters* c, unsigned long long primary) {
    if (threadIdx.x <= PGRT_MAX_LEVELS && threadIdx.x > 0) c->lv_rays[threadIdx.x] += c->n_rays[threadIdx.x];
    if (threadIdx.x == 0) { c->lv_rays[0] += primary; c->tot_shadow += c->shadow; c->tot_reflection += c->reflection; c->tot_refraction += c->refraction; c->tot_primary += primary; }
}
__global__ void k_frame_begin(Counters* c) {
    const int t = threadIdx.x;
    if (t <= PGRT_MAX_LEVELS) {
        c->lv_rays[t] = 0; c->lv_shadow[t] = 0; c->lv_nodes[t] = 0; c->lv_tris[t] = 0; c->lv_sh_nodes[t] = 0; c->lv_sh_tris[t] = 0;
        c->lv_max_nodes[t] = 0; c->lv_sh_max_nodes[t] = 0;
    }
    if (threadIdx.x == 0) { c->overflow = 0; c->tot_shadow = 0; c->tot_reflection = 0; c->tot_refraction = 0; c->tot_primary = 0; }
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
