// raytracer.h -- Raytracer (pg1/raytracer.h:15-52) over the C ABI of include/pgrt.h.
//
// Same construction arguments, same public members and method roles as the reference class; what changed underneath:
//   * the SimpleGuiDX11 base (Win32 window, D3D11, ImGui, the Producer thread) is gone: RenderFrame() is one Producer
//     iteration (pg1/simpleguidx11.cpp:95-118) and get_pixel() serves pixels of that frame;
//   * RTCDevice / RTCScene are one pgrt_context (CUDA, sm_100a); InitDeviceAndScene / ReleaseDeviceAndScene create and
//     destroy it; LoadScene uploads the surfaces and commits (GPU BVH build) instead of rtcCommitScene;
//   * trace() recursion lives on the device; the constants get_pixel / trace hard-code are public members here
//     (sampling_width, focal_distance, aperture, max_depth) with the reference's values as defaults.
// Errors: the reference's Embree error callback throws std::runtime_error (pg1/tutorials.cpp:9-26); so does check().
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/pgrt.h"
#include "camera.h"
#include "objloader.h"

class Raytracer {
public:
    Raytracer(const int width, const int height, const float fov_y, const Vector3 view_from, const Vector3 view_at,
              const char* config = "threads=0,verbose=3");
    ~Raytracer();
    Raytracer(const Raytracer&) = delete;
    Raytracer& operator=(const Raytracer&) = delete;

    float gamma_level = 0.5f;            // pg1/raytracer.h:23; UI slider default pg1/raytracer.cpp:450
    int sampling_width = 3;              // pg1/raytracer.cpp:398
    float focal_distance = 200.0f;       // :399
    float aperture = 5.0f;               // :400
    int max_depth = 7;                   // :282
    unsigned int seed = 1;               // replaces the clock seed of :407
    bool jitter = true;
    // the README's to-do list (README.md:20-21), off by default: hard shadows (hit point -> light) and path tracing
    bool hard_shadows = false;
    bool path_tracing = false;           // shader_mode 3 of pgrt.h; converge with RenderAccumulated

    int InitDeviceAndScene(const char* config);      // config: "device=N" selects the CUDA ordinal; Embree keys are ignored
    int ReleaseDeviceAndScene();
    void LoadScene(const std::string object_file_name, const std::string background_file_name);
    // same, from surfaces already in memory (what LoadScene does after LoadOBJ, pg1/raytracer.cpp:66-127)
    void LoadSurfaces(const std::vector<Surface*>& surfaces, const std::vector<Material*>& materials, const Texture* background);

    Color4f trace(RTCRay ray, int level);                                                 // pg1/raytracer.h:31 (one ray; TraceRays for batches)
    Color4f get_pixel(const int x, const int y, const float t = 0.0f);
    Color4f gamma(Color4f input);
    bool is_illuminated(LightSource light, Vector3 hit_position, Vector3 normal);        // pg1/raytracer.h:34
    RTCRay get_refraction_ray(Vector3 direction, Vector3 normal, float iorFrom, float iorTo, Vector3 hit_point);
    RTCRay get_reflection_ray(Vector3 direction, Vector3 normal, Vector3 hit_point, float ior);

    // one Producer iteration into a caller-owned frame of width*height*4 floats (row 0 = top, a = 1)
    void RenderFrame(float* rgba, pgrt_render_stats* stats = nullptr);
    // progressive loop: the mean of n_frames Producer iterations with seeds seed, seed+1, ... (the reference starts over every iteration)
    void RenderAccumulated(int n_frames, float* rgba, pgrt_render_stats* stats = nullptr);
    // rtcIntersect1 on caller-owned RTCRayHit-compatible records
    void Intersect(pgrt_rayhit* rayhits, size_t n);
    // trace() over a batch of caller-owned rays, all at recursion level `level`; rgba = n x 4 floats
    void TraceRays(const RTCRay* rays, size_t n, int level, float* rgba);

    int width() const { return camera_.width(); }
    int height() const { return camera_.height(); }
    const pgrt_build_stats& build_stats() const { return build_stats_; }
    size_t no_surfaces() const { return surfaces_.size(); }
    size_t no_materials() const { return materials_.size(); }
    const std::vector<Surface*>& surfaces() const { return surfaces_; }
    const std::vector<Material*>& materials() const { return materials_; }
    pgrt_context* context() const { return ctx_; }
    pgrt_render_params params() const;

private:
    void check(int rc) const;
    std::vector<Surface*> surfaces_;
    std::vector<Material*> materials_;
    std::vector<LightSource> lights_;
    SphericalMap background_;
    TextureCache textures_;
    pgrt_context* ctx_ = nullptr;
    PinHoleCamera camera_;
    pgrt_build_stats build_stats_ = {};
    bool owns_scene_ = false;
};

// pg1/tutorials.h:8 -- headless: renders `frames` frames and writes the last one (see pg1_main.cpp for the options)
int raytrace_loop(const std::string object_file_name, const std::string background_file_name, const char* config = "threads=0,verbose=0");
