// ref_api.cpp -- TEST INFRASTRUCTURE ONLY.  C entry points into the reference's OWN classes (Raytracer, PinHoleCamera,
// Texture, SphericalMap, LoadOBJ, mix_srgb, tutorial_1 / tutorial_2), compiled unmodified from /root/reference into
// oracle/_ref/libpg_ref.so by oracle/ref.mk.  tests/test_oracle_vs_ref.py drives them next to the oracle (pg_oracle.cpp)
// on the same inputs: that pins every shading quirk of SURVEY.md App. A to the reference's object code.  The Embree and
// FreeImage calls underneath are stubs (ref_embree.cpp, ref_freeimage.cpp): those two binaries are not vendored.
#include "stdafx.h"
#include "raytracer.h"
#include "objloader.h"
#include "tutorials.h"
#include "material.h"
#include "utils.h"
#include "texture.h"
#include "SphericalMap.h"
#include "mymath.h"
#ifdef _OPENMP
#include <omp.h>
#endif
#undef max
#undef min

namespace {
struct Ref { Raytracer* rt; PinHoleCamera cam; int w, h; };
struct Ftz {   // main() sets FTZ/DAZ before anything runs (pg1_embree.cpp:8-9); here per calling thread
    unsigned saved;
    Ftz() { saved = _mm_getcsr(); _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON); _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON); }
    ~Ftz() { _mm_setcsr(saved); }
};
RTCRay ray_from9(const float* q) {
    RTCRay r;
    r.org_x = q[0]; r.org_y = q[1]; r.org_z = q[2]; r.tnear = q[3]; r.dir_x = q[4]; r.dir_y = q[5]; r.dir_z = q[6]; r.time = q[7]; r.tfar = q[8];
    r.mask = 0; r.id = 0; r.flags = 0;
    return r;
}
void ray_to9(const RTCRay& r, float* o) { o[0] = r.org_x; o[1] = r.org_y; o[2] = r.org_z; o[3] = r.tnear; o[4] = r.dir_x; o[5] = r.dir_y; o[6] = r.dir_z; o[7] = r.time; o[8] = r.tfar; }
}

extern "C" {

// Raytracer( width, height, fov_y, view_from, view_at ) (raytracer.cpp:12-19); gamma_level as the UI leaves it (:450)
void* ref_create(int w, int h, float fov_y, const float* from, const float* at) {
    Ftz f;
    const Vector3 vf(from[0], from[1], from[2]), va(at[0], at[1], at[2]);
    Ref* r = new Ref{new Raytracer(w, h, fov_y, vf, va), PinHoleCamera(w, h, fov_y, vf, va), w, h};
    r->rt->gamma_level = 0.5f;
    return r;
}
void ref_destroy(void* h) { Ref* r = (Ref*)h; if (r) { delete r->rt; delete r; } }
// Raytracer::LoadScene (raytracer.cpp:48-128): LoadOBJ + LoadMTL + SphericalMap + the white light + the Embree upload
int ref_load_scene(void* h, const char* obj, const char* background) {
    Ftz f;
    try { ((Ref*)h)->rt->LoadScene(obj, background); } catch (const std::exception& e) { fprintf(stderr, "ref_load_scene: %s\n", e.what()); return 1; }
    return 0;
}
void ref_set_gamma(void* h, float g) { ((Ref*)h)->rt->gamma_level = g; }

// Raytracer::trace (raytracer.cpp:237-394) on caller-supplied rays: 9 floats in (org, tnear, dir, time, tfar), Color4f out
void ref_trace(void* h, const float* rays9, uint64_t n, int level, float* out4, int threads) {
    Raytracer* rt = ((Ref*)h)->rt;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads > 0 ? threads : omp_get_max_threads())
#endif
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        Ftz f;
        const Color4f c = rt->trace(ray_from9(rays9 + 9 * i), level);
        out4[4 * i] = c.r; out4[4 * i + 1] = c.g; out4[4 * i + 2] = c.b; out4[4 * i + 3] = c.a;
    }
}
// PinHoleCamera::generate_ray, both overloads (PinHoleCamera.cpp:31-63, :65-105); the lens overload draws its shift from a
// clock-seeded mt19937 (:77), so only aperture 0 is reproducible
void ref_generate_rays(void* h, const float* xy, uint64_t n, float focal, float aperture, int pinhole, float* out9) {
    Ftz f;
    const PinHoleCamera& cam = ((Ref*)h)->cam;
    for (uint64_t i = 0; i < n; ++i) ray_to9(pinhole ? cam.generate_ray(xy[2 * i], xy[2 * i + 1]) : cam.generate_ray(xy[2 * i], xy[2 * i + 1], focal, aperture), out9 + 9 * i);
}
// one frame without the clock-seeded jitter: get_pixel (raytracer.cpp:396-437) with sampling_width 1 at the pixel's own
// coordinate: generate_ray(x, y, focal, 0), time = IOR_AIR (:416), trace, sum, swap, gamma (:421-436); Producer's pixel
// order and packing (simpleguidx11.cpp:105-118)
void ref_render_unjittered(void* h, float focal, float* rgba, int threads) {
    Ref* r = (Ref*)h;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : omp_get_max_threads())
#endif
    for (int y = 0; y < r->h; ++y) {
        Ftz f;
        for (int x = 0; x < r->w; ++x) {
            RTCRay primary_ray = r->cam.generate_ray((float)x, (float)y, focal, 0.0f);
            primary_ray.time = IOR_AIR;
            const Color4f c = r->rt->trace(primary_ray, 0);
            Color4f final_color{0.0f, 0.0f, 0.0f, 1.0f};
            final_color.r += c.r; final_color.g += c.g; final_color.b += c.b;
            const Color4f pixel = r->rt->gamma(Color4f{final_color.b / 1, final_color.g / 1, final_color.r / 1, 1.0f});
            const int offset = (y * r->w + x) * 4;
            rgba[offset] = pixel.r; rgba[offset + 1] = pixel.g; rgba[offset + 2] = pixel.b; rgba[offset + 3] = pixel.a;
        }
    }
}
// Raytracer::get_pixel exactly as shipped (3x3 stratified, clock-seeded jitter and lens shift): comparable as a converged mean only
void ref_get_pixels(void* h, int x0, int y0, int x1, int y1, float* rgba, int threads) {
    Ref* r = (Ref*)h;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : omp_get_max_threads())
#endif
    for (int y = y0; y < y1; ++y) {
        Ftz f;
        for (int x = x0; x < x1; ++x) {
            const Color4f pixel = r->rt->get_pixel(x, y, 0.0f);
            const int offset = ((y - y0) * (x1 - x0) + (x - x0)) * 4;
            rgba[offset] = pixel.r; rgba[offset + 1] = pixel.g; rgba[offset + 2] = pixel.b; rgba[offset + 3] = pixel.a;
        }
    }
}
void ref_gamma(void* h, const float* in4, uint64_t n, float* out4) {
    Ftz f;
    for (uint64_t i = 0; i < n; ++i) {
        const Color4f c = ((Ref*)h)->rt->gamma(Color4f{in4[4 * i], in4[4 * i + 1], in4[4 * i + 2], in4[4 * i + 3]});
        out4[4 * i] = c.r; out4[4 * i + 1] = c.g; out4[4 * i + 2] = c.b; out4[4 * i + 3] = c.a;
    }
}
// Raytracer::is_illuminated (raytracer.cpp:150-176): light position, hit position, normal -> 0 / 1
void ref_is_illuminated(void* h, const float* light3, const float* hit3, const float* nrm3, uint64_t n, int32_t* out) {
    Ftz f;
    const Vector3 one(1, 1, 1);
    for (uint64_t i = 0; i < n; ++i) {
        LightSource l(Vector3(light3[3 * i], light3[3 * i + 1], light3[3 * i + 2]), one, one, one);
        out[i] = ((Ref*)h)->rt->is_illuminated(l, Vector3(hit3[3 * i], hit3[3 * i + 1], hit3[3 * i + 2]), Vector3(nrm3[3 * i], nrm3[3 * i + 1], nrm3[3 * i + 2])) ? 1 : 0;
    }
}
// get_reflection_ray / get_refraction_ray (raytracer.cpp:178-235): in = dir(3) normal(3) hit(3) n1 n2
void ref_secondary_rays(void* h, const float* in11, uint64_t n, int refraction, float* out9) {
    Ftz f;
    Raytracer* rt = ((Ref*)h)->rt;
    for (uint64_t i = 0; i < n; ++i) {
        const float* q = in11 + 11 * i;
        const Vector3 d(q[0], q[1], q[2]), nn(q[3], q[4], q[5]), hp(q[6], q[7], q[8]);
        ray_to9(refraction ? rt->get_refraction_ray(d, nn, q[9], q[10], hp) : rt->get_reflection_ray(d, nn, hp, q[9]), out9 + 9 * i);
    }
}
// mix_srgb (utils.cpp:238-241)
void ref_mix_srgb(const float* c0, const float* c1, const float* alpha, uint64_t n, float* out4) {
    Ftz f;
    for (uint64_t i = 0; i < n; ++i) {
        const Color4f c = mix_srgb(Color4f{c0[4 * i], c0[4 * i + 1], c0[4 * i + 2], c0[4 * i + 3]}, Color4f{c1[4 * i], c1[4 * i + 1], c1[4 * i + 2], c1[4 * i + 3]}, alpha[i]);
        out4[4 * i] = c.r; out4[4 * i + 1] = c.g; out4[4 * i + 2] = c.b; out4[4 * i + 3] = c.a;
    }
}
// Texture (texture.cpp:5-130) and SphericalMap (SphericalMap.cpp:6-29)
void* ref_texture_load(const char* file) { Texture* t = new Texture(file); if (t->width() == 0) { delete t; return nullptr; } return t; }
void ref_texture_free(void* t) { delete (Texture*)t; }
void ref_texture_size(void* t, int* wh) { wh[0] = ((Texture*)t)->width(); wh[1] = ((Texture*)t)->height(); }
void ref_texture_get_texel(void* t, const float* uv, uint64_t n, float* out3) {
    Ftz f;
    for (uint64_t i = 0; i < n; ++i) { const Color3f c = ((Texture*)t)->get_texel(uv[2 * i], uv[2 * i + 1]); out3[3 * i] = c.r; out3[3 * i + 1] = c.g; out3[3 * i + 2] = c.b; }
}
void* ref_envmap_load(const char* file) { return new SphericalMap(std::string(file)); }
void ref_envmap_free(void* e) { delete (SphericalMap*)e; }
void ref_env_get_texel(void* e, const float* dirs, uint64_t n, float* out4) {
    Ftz f;
    for (uint64_t i = 0; i < n; ++i) {
        const Color4f c = ((SphericalMap*)e)->get_texel(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
        out4[4 * i] = c.r; out4[4 * i + 1] = c.g; out4[4 * i + 2] = c.b; out4[4 * i + 3] = c.a;
    }
}
// LoadOBJ + LoadMTL (objloader.cpp:53-507): surfaces as flat arrays for comparison with the host loader of the product.
// Two calls: sizes first (tris_per_surface[n_surfaces]), then the arrays.
void* ref_obj_load(const char* file, int* n_surfaces, int* n_materials) {
    auto* pair = new std::pair<std::vector<Surface*>, std::vector<Material*>>();
    LoadOBJ(file, pair->first, pair->second);
    *n_surfaces = (int)pair->first.size(); *n_materials = (int)pair->second.size();
    return pair;
}
int ref_obj_surface_triangles(void* o, int s) { return ((std::pair<std::vector<Surface*>, std::vector<Material*>>*)o)->first[s]->no_triangles(); }
// pos / nrm: 9 floats per triangle, uv: 6 per triangle, exactly what Raytracer::LoadScene hands to Embree (raytracer.cpp:99-119)
int ref_obj_surface(void* o, int s, float* pos, float* nrm, float* uv, char* name64, int* material_index) {
    auto* pair = (std::pair<std::vector<Surface*>, std::vector<Material*>>*)o;
    Surface* surface = pair->first[s];
    for (int i = 0, k = 0; i < surface->no_triangles(); ++i) {
        Triangle& triangle = surface->get_triangle(i);
        for (int j = 0; j < 3; ++j, ++k) {
            const Vertex& vertex = triangle.vertex(j);
            pos[3 * k] = vertex.position.x; pos[3 * k + 1] = vertex.position.y; pos[3 * k + 2] = vertex.position.z;
            nrm[3 * k] = vertex.normal.x; nrm[3 * k + 1] = vertex.normal.y; nrm[3 * k + 2] = vertex.normal.z;
            uv[2 * k] = vertex.texture_coords[0].u; uv[2 * k + 1] = vertex.texture_coords[0].v;
        }
    }
    strncpy(name64, surface->get_name().c_str(), 63); name64[63] = 0;
    *material_index = -1;
    for (size_t m = 0; m < pair->second.size(); ++m) if (pair->second[m] == surface->get_material()) *material_index = (int)m;
    return 0;
}
// Material fields the path reads: out = Kd(3) Ks(3) Ns Ni type has_diffuse_texture; name64 = material name
int ref_obj_material(void* o, int m, float* out10, char* name64) {
    auto* pair = (std::pair<std::vector<Surface*>, std::vector<Material*>>*)o;
    Material* mt = pair->second[m];
    out10[0] = mt->diffuse.x; out10[1] = mt->diffuse.y; out10[2] = mt->diffuse.z; out10[3] = mt->specular.x; out10[4] = mt->specular.y; out10[5] = mt->specular.z;
    out10[6] = mt->shininess; out10[7] = mt->ior; out10[8] = (float)mt->type; out10[9] = mt->get_texture(Material::kDiffuseMapSlot) ? 1.0f : 0.0f;
    strncpy(name64, mt->get_name().c_str(), 63); name64[63] = 0;
    return 0;
}
void ref_obj_free(void* o) { delete (std::pair<std::vector<Surface*>, std::vector<Material*>>*)o; }   // (the reference never frees its surfaces either)

// the reference's own known-answer demos (tutorials.cpp:147-179); they print to stdout
int ref_tutorial_1(const char* config) { Ftz f; const int rc = tutorial_1(config); fflush(stdout); return rc; }
int ref_tutorial_2() { Ftz f; const int rc = tutorial_2(); fflush(stdout); return rc; }
float ref_deg2rad(float d) { return deg2rad(d); }
}
