// pg1_capi.cpp -- flat C entry points over the C++ host classes, so tests/ can drive LoadOBJ / LoadMTL / Texture /
// PinHoleCamera / Raytracer through ctypes exactly as a C++ caller would.  Nothing here computes pixels: Raytracer
// goes through libpgrt_b200.so.
#include <cstring>
#include <string>
#include <vector>
#include "image_io.h"
#include "raytracer.h"

namespace {
struct LoadedScene { std::vector<Surface*> surfaces; std::vector<Material*> materials; TextureCache textures; int rc = 0; };
thread_local std::string g_err;
}  // namespace

extern "C" {

const char* pg1_last_error() { return g_err.c_str(); }

// ---- loader
void* pg1_load_obj(const char* file_name, int flip_yz) {
    LoadedScene* s = new LoadedScene();
    s->rc = LoadOBJ(file_name, s->surfaces, s->materials, flip_yz != 0, Vector3(0.5f, 0.5f, 0.5f), &s->textures);
    return s;
}
void* pg1_load_mtl(const char* file_name, const char* path) {
    LoadedScene* s = new LoadedScene();
    s->rc = LoadMTL(file_name, path, s->materials, &s->textures);
    return s;
}
void pg1_free_scene(void* h) {
    LoadedScene* s = (LoadedScene*)h;
    for (Surface* x : s->surfaces) delete x;
    for (Material* m : s->materials) delete m;
    ReleaseTextureCache(s->textures);
    delete s;
}
int pg1_scene_rc(void* h) { return ((LoadedScene*)h)->rc; }
int pg1_num_surfaces(void* h) { return (int)((LoadedScene*)h)->surfaces.size(); }
int pg1_num_materials(void* h) { return (int)((LoadedScene*)h)->materials.size(); }
int pg1_surface_triangles(void* h, int i) { return ((LoadedScene*)h)->surfaces[i]->no_triangles(); }
const char* pg1_surface_name(void* h, int i) { static thread_local std::string s; s = ((LoadedScene*)h)->surfaces[i]->get_name(); return s.c_str(); }
int pg1_surface_material(void* h, int i) {
    LoadedScene* s = (LoadedScene*)h;
    for (size_t k = 0; k < s->materials.size(); ++k) if (s->materials[k] == s->surfaces[i]->get_material()) return (int)k;
    return -1;
}
void pg1_surface_data(void* h, int i, float* pos, float* nrm, float* uv) {
    const Surface* s = ((LoadedScene*)h)->surfaces[i];
    memcpy(pos, s->positions.data(), s->positions.size() * 4); memcpy(nrm, s->normals.data(), s->normals.size() * 4);
    memcpy(uv, s->tex_coords.data(), s->tex_coords.size() * 4);
}
// out16: ambient, diffuse, specular, emission (12), shininess, ior, type, has diffuse map
const char* pg1_material(void* h, int i, float* out16) {
    static thread_local std::string name;
    const Material* m = ((LoadedScene*)h)->materials[i];
    const Vector3* v[4] = {&m->ambient, &m->diffuse, &m->specular, &m->emission};
    for (int k = 0; k < 4; ++k) { out16[3 * k] = v[k]->x; out16[3 * k + 1] = v[k]->y; out16[3 * k + 2] = v[k]->z; }
    out16[12] = m->shininess; out16[13] = m->ior; out16[14] = (float)m->type;
    const Texture* t = m->get_texture(Material::kDiffuseMapSlot);
    out16[15] = (t && t->valid()) ? 1.0f : 0.0f;
    name = m->get_name();
    return name.c_str();
}

// ---- images
void* pg1_load_image(const char* file_name) {
    RawImage* im = new RawImage();
    if (!LoadImageFile(file_name, *im, &g_err)) { delete im; return nullptr; }
    return im;
}
void pg1_image_info(void* h, int* out4) { const RawImage* im = (RawImage*)h; out4[0] = im->width; out4[1] = im->height; out4[2] = im->pitch; out4[3] = im->bpp; }
void pg1_image_bytes(void* h, unsigned char* dst) { const RawImage* im = (RawImage*)h; memcpy(dst, im->bytes.data(), im->bytes.size()); }
void pg1_free_image(void* h) { delete (RawImage*)h; }
int pg1_write_ppm(const char* file_name, const float* rgba, int w, int h) { return WritePPM(file_name, rgba, w, h) ? 0 : -1; }
int pg1_write_image(const char* file_name, const float* rgba, int w, int h) { return WriteImageFile(file_name, rgba, w, h) ? 0 : -1; }

// ---- camera (host copy)
void pg1_camera_ray(int w, int h, float fov_y, const float* from, const float* at, float x, float y, int lens, float focal, float r1, float r2, float* out9) {
    PinHoleCamera cam(w, h, fov_y, Vector3(from[0], from[1], from[2]), Vector3(at[0], at[1], at[2]));
    const RTCRay r = lens ? cam.generate_ray(x, y, focal, r1, r2) : cam.generate_ray(x, y);
    out9[0] = r.org_x; out9[1] = r.org_y; out9[2] = r.org_z; out9[3] = r.tnear; out9[4] = r.dir_x; out9[5] = r.dir_y; out9[6] = r.dir_z; out9[7] = r.time; out9[8] = r.tfar;
}

// ---- Raytracer
void* pg1_raytracer_create(int w, int h, float fov_y, const float* from, const float* at, const char* config) {
    try { return new Raytracer(w, h, fov_y, Vector3(from[0], from[1], from[2]), Vector3(at[0], at[1], at[2]), config); }
    catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void pg1_raytracer_destroy(void* h) { delete (Raytracer*)h; }
int pg1_raytracer_load_scene(void* h, const char* obj, const char* bg) {
    try { ((Raytracer*)h)->LoadScene(obj, bg); return 0; } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
void pg1_raytracer_set(void* h, int sampling_width, int jitter, float focal, float aperture, int max_depth, float gamma_level, unsigned seed) {
    Raytracer* r = (Raytracer*)h;
    r->sampling_width = sampling_width; r->jitter = jitter != 0; r->focal_distance = focal; r->aperture = aperture; r->max_depth = max_depth;
    r->gamma_level = gamma_level; r->seed = seed;
}
int pg1_raytracer_render(void* h, float* rgba, unsigned long long* rays4) {
    try {
        pgrt_render_stats st;
        ((Raytracer*)h)->RenderFrame(rgba, &st);
        if (rays4) { rays4[0] = st.rays_primary; rays4[1] = st.rays_shadow; rays4[2] = st.rays_reflection; rays4[3] = st.rays_refraction; }
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int pg1_raytracer_get_pixel(void* h, int x, int y, float* rgba) {
    try { const Color4f c = ((Raytracer*)h)->get_pixel(x, y); rgba[0] = c.r; rgba[1] = c.g; rgba[2] = c.b; rgba[3] = c.a; return 0; }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int pg1_raytracer_trace(void* h, const float* rays12, unsigned long long n, int level, float* rgba) {
    try {
        if (n == 1) { const Color4f c = ((Raytracer*)h)->trace(*reinterpret_cast<const RTCRay*>(rays12), level); rgba[0] = c.r; rgba[1] = c.g; rgba[2] = c.b; rgba[3] = c.a; }
        else ((Raytracer*)h)->TraceRays(reinterpret_cast<const RTCRay*>(rays12), (size_t)n, level, rgba);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int pg1_raytracer_is_illuminated(void* h, const float* light3, const float* hit3, const float* nrm3) {
    try {
        const Vector3 one(1, 1, 1);
        return ((Raytracer*)h)->is_illuminated(LightSource(Vector3(light3[0], light3[1], light3[2]), one, one, one), Vector3(hit3[0], hit3[1], hit3[2]),
                                               Vector3(nrm3[0], nrm3[1], nrm3[2])) ? 1 : 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int pg1_raytracer_counts(void* h, int* out2) { out2[0] = (int)((Raytracer*)h)->no_surfaces(); out2[1] = (int)((Raytracer*)h)->no_materials(); return 0; }

}  // extern "C"
