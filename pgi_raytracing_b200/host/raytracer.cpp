#include "raytracer.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

Raytracer::Raytracer(const int width, const int height, const float fov_y, const Vector3 view_from, const Vector3 view_at, const char* config)
    : camera_(width, height, fov_y, view_from, view_at) {
    if (InitDeviceAndScene(config) != 0) throw std::runtime_error("Raytracer: no CUDA device (the render loop has no CPU fallback)");
    const float from[3] = {view_from.x, view_from.y, view_from.z}, at[3] = {view_at.x, view_at.y, view_at.z};
    try { check(pgrt_set_camera(ctx_, width, height, fov_y, from, at)); }
    catch (...) { ReleaseDeviceAndScene(); throw; }      // the destructor does not run for a constructor that throws
}

Raytracer::~Raytracer() {
    ReleaseDeviceAndScene();
    if (owns_scene_) {
        for (Surface* s : surfaces_) delete s;
        for (Material* m : materials_) delete m;
    }
    ReleaseTextureCache(textures_);
}

void Raytracer::check(int rc) const {
    if (rc != PGRT_OK) throw std::runtime_error(std::string("pgrt error ") + std::to_string(rc) + ": " + pgrt_last_error(ctx_));
}

int Raytracer::InitDeviceAndScene(const char* config) {
    int device = 0;
    if (config) if (const char* p = strstr(config, "device=")) device = atoi(p + 7);
    return pgrt_create(&ctx_, device) == PGRT_OK ? 0 : -1;
}

int Raytracer::ReleaseDeviceAndScene() {
    if (ctx_) { pgrt_destroy(ctx_); ctx_ = nullptr; }
    return 0;
}

void Raytracer::LoadScene(const std::string object_file_name, const std::string background_file_name) {
    LoadOBJ(object_file_name.c_str(), surfaces_, materials_, false, Vector3(0.5f, 0.5f, 0.5f), &textures_);
    owns_scene_ = true;
    background_ = SphericalMap(background_file_name);
    LoadSurfaces(surfaces_, materials_, background_.texture());
}

void Raytracer::LoadSurfaces(const std::vector<Surface*>& surfaces, const std::vector<Material*>& materials, const Texture* background) {
    if (&surfaces != &surfaces_) surfaces_ = surfaces;
    if (&materials != &materials_) materials_ = materials;
    check(pgrt_clear_scene(ctx_));
    // one white light (pg1/raytracer.cpp:66-68)
    lights_.clear();
    lights_.push_back(LightSource(Vector3(-500, -100, 500), Vector3(1, 1, 1), Vector3(1, 1, 1), Vector3(1, 1, 1)));
    // materials and their diffuse maps (the only slot the path reads, pg1/raytracer.cpp:337)
    std::map<const Texture*, int> tex_id;
    std::vector<pgrt_material> mats(materials_.size() + 1);
    for (size_t i = 0; i < materials_.size(); ++i) {
        const Material* m = materials_[i];
        pgrt_material& o = mats[i];
        o.diffuse[0] = m->diffuse.x; o.diffuse[1] = m->diffuse.y; o.diffuse[2] = m->diffuse.z;
        o.specular[0] = m->specular.x; o.specular[1] = m->specular.y; o.specular[2] = m->specular.z;
        o.shininess = m->shininess; o.ior = m->ior; o.type = m->type; o.diffuse_tex = -1;
        const Texture* t = m->get_texture(Material::kDiffuseMapSlot);
        if (t && t->valid()) {
            auto it = tex_id.find(t);
            if (it == tex_id.end()) {
                const int id = (int)tex_id.size();
                check(pgrt_set_texture(ctx_, id, t->data(), t->width(), t->height(), t->scan_width(), t->pixel_size()));
                it = tex_id.emplace(t, id).first;
            }
            o.diffuse_tex = it->second;
        }
    }
    {   // a surface without a material (no matching usemtl) dereferences NULL in the reference; here it gets Material()
        const Material dflt; pgrt_material& o = mats[materials_.size()];
        o.diffuse[0] = dflt.diffuse.x; o.diffuse[1] = dflt.diffuse.y; o.diffuse[2] = dflt.diffuse.z;
        o.specular[0] = dflt.specular.x; o.specular[1] = dflt.specular.y; o.specular[2] = dflt.specular.z;
        o.shininess = dflt.shininess; o.ior = dflt.ior; o.type = dflt.type; o.diffuse_tex = -1;
    }
    check(pgrt_set_materials(ctx_, mats.data(), (int)mats.size()));
    if (background && background->valid())
        check(pgrt_set_envmap(ctx_, background->data(), background->width(), background->height(), background->scan_width(), background->pixel_size()));
    std::vector<pgrt_light> ls(lights_.size());
    for (size_t i = 0; i < lights_.size(); ++i) {
        const LightSource& l = lights_[i];
        const Vector3* src[4] = {&l.position_, &l.ambient_, &l.diffuse_, &l.spectular_};
        float* dst[4] = {ls[i].position, ls[i].ambient, ls[i].diffuse, ls[i].specular};
        for (int k = 0; k < 4; ++k) { dst[k][0] = src[k]->x; dst[k][1] = src[k]->y; dst[k][2] = src[k]->z; }
    }
    check(pgrt_set_lights(ctx_, ls.data(), (int)ls.size()));
    // surfaces loop (pg1/raytracer.cpp:71-125): geomID = attach order
    for (const Surface* s : surfaces_) {
        int mat = (int)materials_.size();
        for (size_t i = 0; i < materials_.size(); ++i) if (materials_[i] == s->get_material()) { mat = (int)i; break; }
        uint32_t geom_id = 0;
        check(pgrt_add_mesh(ctx_, s->positions.data(), s->normals.data(), s->tex_coords.data(), (uint32_t)s->no_triangles(), mat, &geom_id));
    }
    check(pgrt_commit(ctx_, &build_stats_));   // rtcCommitScene (:127)
}

pgrt_render_params Raytracer::params() const {
    pgrt_render_params p;
    pgrt_default_params(&p);
    p.sampling_width = sampling_width; p.jitter = jitter ? 1 : 0; p.focal_distance = focal_distance; p.aperture = aperture;
    p.max_depth = max_depth; p.gamma_level = gamma_level; p.seed = seed;
    p.shadow_mode = hard_shadows ? 1 : 0;
    if (path_tracing) p.shader_mode = 3;
    return p;
}

void Raytracer::RenderFrame(float* rgba, pgrt_render_stats* stats) {
    const pgrt_render_params p = params();
    check(pgrt_render(ctx_, &p, rgba, stats, 0));
}

void Raytracer::RenderAccumulated(int n_frames, float* rgba, pgrt_render_stats* stats) {
    const pgrt_render_params p = params();
    check(pgrt_render_accumulate(ctx_, &p, n_frames, rgba, stats));
}

Color4f Raytracer::get_pixel(const int x, const int y, const float /*t: ignored, as in the reference*/) {
    const pgrt_render_params p = params();
    float px[4];
    check(pgrt_get_pixel(ctx_, &p, x, y, px));
    return Color4f{px[0], px[1], px[2], px[3]};
}

Color4f Raytracer::gamma(Color4f input) {
    float in[4] = {input.r, input.g, input.b, input.a}, out[4];
    check(pgrt_eval_gamma(ctx_, in, gamma_level, 1, out));
    return Color4f{out[0], out[1], out[2], out[3]};
}

static RTCRay unpack_ray(const float* o) {
    RTCRay r;
    r.org_x = o[0]; r.org_y = o[1]; r.org_z = o[2]; r.tnear = o[3]; r.dir_x = o[4]; r.dir_y = o[5]; r.dir_z = o[6]; r.time = o[7]; r.tfar = o[8];
    r.mask = 0; r.id = 0; r.flags = 0;
    return r;
}

RTCRay Raytracer::get_refraction_ray(Vector3 direction, Vector3 normal, float iorFrom, float iorTo, Vector3 hit_point) {
    const float in[11] = {direction.x, direction.y, direction.z, normal.x, normal.y, normal.z, hit_point.x, hit_point.y, hit_point.z, iorFrom, iorTo};
    float out[9];
    check(pgrt_eval_secondary_rays(ctx_, in, 1, 1, out));
    return unpack_ray(out);
}

RTCRay Raytracer::get_reflection_ray(Vector3 direction, Vector3 normal, Vector3 hit_point, float ior) {
    const float in[11] = {direction.x, direction.y, direction.z, normal.x, normal.y, normal.z, hit_point.x, hit_point.y, hit_point.z, ior, 0.0f};
    float out[9];
    check(pgrt_eval_secondary_rays(ctx_, in, 1, 0, out));
    return unpack_ray(out);
}

void Raytracer::Intersect(pgrt_rayhit* rayhits, size_t n) { check(pgrt_intersect(ctx_, rayhits, n)); }

// Raytracer::trace (pg1/raytracer.cpp:237-394): the recursion runs on the device (pgrt_trace), in the reference's own order
void Raytracer::TraceRays(const RTCRay* rays, size_t n, int level, float* rgba) {
    static_assert(sizeof(RTCRay) == sizeof(pgrt_ray), "RTCRay layout");
    const pgrt_render_params p = params();
    check(pgrt_trace(ctx_, &p, reinterpret_cast<const pgrt_ray*>(rays), n, level, rgba));
}

Color4f Raytracer::trace(RTCRay ray, int level) {
    float px[4];
    TraceRays(&ray, 1, level, px);
    return Color4f{px[0], px[1], px[2], px[3]};
}

// Raytracer::is_illuminated (pg1/raytracer.cpp:150-176)
bool Raytracer::is_illuminated(LightSource light, Vector3 hit_position, Vector3 normal) {
    const pgrt_render_params p = params();
    const float l[3] = {light.position_.x, light.position_.y, light.position_.z}, h[3] = {hit_position.x, hit_position.y, hit_position.z},
                n[3] = {normal.x, normal.y, normal.z};
    int32_t lit = 0;
    check(pgrt_is_illuminated(ctx_, &p, l, h, n, 1, &lit));
    return lit != 0;
}
