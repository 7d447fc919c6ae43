// pg1_main.cpp -- headless counterpart of the reference's entry point (pg1/pg1_embree.cpp:4-22 + raytrace_loop,
// pg1/tutorials.cpp:181-200): same scene files, same camera, same per-pixel settings; the window and the endless
// Producer loop become `--frames N` iterations and an image file.
//
//   pg1_b200 [obj] [background] [--width W --height H] [--frames N] [--spp-width S] [--aperture A] [--focal F] [--depth D]
//            [--gamma G] [--seed K] [--no-jitter] [--device N] [--accumulate N] [--path-tracing] [--hard-shadows]
//            [--out frame.ppm|.png|.pfm] [--pfm frame.pfm]
// Defaults = the reference's hard-coded values: ../../../data/6887_allied_avenger.obj, ../../../data/spherical_map_lakeside.jpg,
// 640x480, fov_y 42.185 deg, eye (-140,-175,80) -> (0,0,40), 3x3 jittered samples, thin lens f=200 a=5, depth 7, gamma 0.5.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "image_io.h"
#include "raytracer.h"

namespace {
struct Options {
    std::string obj = "../../../data/6887_allied_avenger.obj", bg = "../../../data/spherical_map_lakeside.jpg";
    int width = 640, height = 480, frames = 3, spp = 3, depth = 7, device = 0, accumulate = 0;
    float aperture = 5.0f, focal = 200.0f, gamma = 0.5f;
    unsigned seed = 1; bool jitter = true, path_tracing = false, hard_shadows = false;
    std::string out = "frame.ppm", pfm;
} g_opt;
}  // namespace

int raytrace_loop(const std::string object_file_name, const std::string background_file_name, const char* config) {
    std::string cfg = std::string(config ? config : "") + ",device=" + std::to_string(g_opt.device);
    Raytracer raytracer(g_opt.width, g_opt.height, deg2rad(42.185f), Vector3(-140, -175, 80), Vector3(0, 0, 40), cfg.c_str());
    raytracer.sampling_width = g_opt.spp; raytracer.aperture = g_opt.aperture; raytracer.focal_distance = g_opt.focal;
    raytracer.max_depth = g_opt.depth; raytracer.gamma_level = g_opt.gamma; raytracer.seed = g_opt.seed; raytracer.jitter = g_opt.jitter;
    raytracer.path_tracing = g_opt.path_tracing; raytracer.hard_shadows = g_opt.hard_shadows;
    raytracer.LoadScene(object_file_name, background_file_name);
    const pgrt_build_stats& bs = raytracer.build_stats();
    printf("Surfaces = %zu\nMaterials = %zu\n", raytracer.no_surfaces(), raytracer.no_materials());   // pg1/raytracer.cpp:456-457
    printf("BVH: %u triangles, %u wide nodes, depth %u, SAH %.2f, built in %.3f ms on the GPU\n", bs.triangles, bs.nodes, bs.depth, bs.sah_cost, bs.build_ms);
    std::vector<float> frame((size_t)g_opt.width * g_opt.height * 4);
    for (int f = 0; f < g_opt.frames; ++f) {   // MainLoop / Producer (pg1/simpleguidx11.cpp:95-125), `frames` iterations
        pgrt_render_stats st;
        const auto t0 = std::chrono::steady_clock::now();
        raytracer.RenderFrame(frame.data(), &st);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        const unsigned long long rays = st.rays_primary + st.rays_shadow + st.rays_reflection + st.rays_refraction;
        printf("frame %d: %.3f ms host (%.3f ms device), %llu rays (primary %llu, shadow %llu, reflection %llu, refraction %llu), %.1f Mrays/s\n", f, ms,
               st.frame_ms, rays, (unsigned long long)st.rays_primary, (unsigned long long)st.rays_shadow, (unsigned long long)st.rays_reflection,
               (unsigned long long)st.rays_refraction, rays / (ms * 1e3));
    }
    if (g_opt.accumulate > 0) {                // progressive: the mean of N iterations with consecutive seeds
        pgrt_render_stats st;
        raytracer.RenderAccumulated(g_opt.accumulate, frame.data(), &st);
        printf("accumulated %d frames: %.3f ms device, %llu rays\n", g_opt.accumulate, st.frame_ms,
               (unsigned long long)(st.rays_primary + st.rays_shadow + st.rays_reflection + st.rays_refraction));
    }
    if (!g_opt.out.empty() && !WriteImageFile(g_opt.out.c_str(), frame.data(), g_opt.width, g_opt.height)) printf("cannot write %s\n", g_opt.out.c_str());
    if (!g_opt.pfm.empty()) WritePFM(g_opt.pfm.c_str(), frame.data(), g_opt.width, g_opt.height);
    return EXIT_SUCCESS;
}

#ifndef PG1_NO_MAIN
int main(int argc, char** argv) {
    printf("PG1 ray tracer, B200 render loop (%s)\n", pgrt_version());
    int positional = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "%s needs a value\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--width") g_opt.width = atoi(val());
        else if (a == "--height") g_opt.height = atoi(val());
        else if (a == "--frames") g_opt.frames = atoi(val());
        else if (a == "--spp-width") g_opt.spp = atoi(val());
        else if (a == "--aperture") g_opt.aperture = (float)atof(val());
        else if (a == "--focal") g_opt.focal = (float)atof(val());
        else if (a == "--depth") g_opt.depth = atoi(val());
        else if (a == "--gamma") g_opt.gamma = (float)atof(val());
        else if (a == "--seed") g_opt.seed = (unsigned)strtoul(val(), nullptr, 10);
        else if (a == "--no-jitter") g_opt.jitter = false;
        else if (a == "--device") g_opt.device = atoi(val());
        else if (a == "--accumulate") g_opt.accumulate = atoi(val());
        else if (a == "--path-tracing") g_opt.path_tracing = true;
        else if (a == "--hard-shadows") g_opt.hard_shadows = true;
        else if (a == "--out") g_opt.out = val();
        else if (a == "--pfm") g_opt.pfm = val();
        else if (positional == 0) { g_opt.obj = a; positional++; }
        else if (positional == 1) { g_opt.bg = a; positional++; }
        else { fprintf(stderr, "unexpected argument %s\n", a.c_str()); return 2; }
    }
    try {
        return raytrace_loop(g_opt.obj, g_opt.bg, "threads=0,verbose=0");
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return EXIT_FAILURE;
    }
}
#endif
