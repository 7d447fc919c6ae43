/* pgrt.h -- C ABI of the B200-native render loop that replaces, in rddrdhd/PGI_RayTracing,
 *   (B1) the per-pixel loop  SimpleGuiDX11::Producer -> Raytracer::get_pixel -> Raytracer::trace
 *        (src/pg/pg1_embree/simpleguidx11.cpp:86-128, raytracer.cpp:396-437, :237-394), and
 *   (B2) the Embree 3 calls underneath it (rtcNewGeometry ... rtcCommitScene, rtcIntersect1, rtcInterpolate0;
 *        src/pg/pg1_embree/raytracer.cpp:26-47, :71-127, :130-148, :249-253, :344).
 *
 * Plain C: opaque context, POD structs, pointers + sizes, int status (0 = PGRT_OK), no exceptions across the
 * boundary, no torch / CUDA types in any signature (streams and device buffers travel as void*).
 * All reference citations below are relative to src/pg/pg1_embree/ of the reference repository.
 * The library is CUDA-only: there is no CPU fallback; pgrt_create fails when no sm_100 device is present.
 */
#ifndef PGRT_H_
#define PGRT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGRT_OK 0
#define PGRT_ERR_INVALID 1      /* bad argument / call order                               */
#define PGRT_ERR_CUDA 2         /* a CUDA runtime call failed; see pgrt_last_error         */
#define PGRT_ERR_NO_DEVICE 3    /* no usable GPU                                           */
#define PGRT_ERR_OVERFLOW 4     /* secondary-ray queues overflowed even at the minimum batch */

#define PGRT_MAX_INFLIGHT 32           /* frame slots of a context (pipelined frames)             */

#define PGRT_INVALID_ID 0xFFFFFFFFu   /* = RTC_INVALID_GEOMETRY_ID, embree3/rtcore_common.h:45 */
#define PGRT_IOR_AIR 1.000293f        /* material.h:15 */

typedef struct pgrt_context pgrt_context;

/* Fields of Material the path reads (material.h:96-108); filled by LoadMTL (objloader.cpp:53-208). */
typedef struct pgrt_material {
    float diffuse[3];     /* Kd as parsed (x,y,z)                      */
    float specular[3];    /* Ks as parsed                              */
    float shininess;      /* Ns                                        */
    float ior;            /* Ni                                        */
    int32_t type;         /* MTL "shader N": 4 = dielectric, else Phong (raytracer.cpp:293-325) */
    int32_t diffuse_tex;  /* texture id given to pgrt_set_texture, -1 = none (Material::kDiffuseMapSlot) */
} pgrt_material;

/* LightSource (LightSource.h:11-15). */
typedef struct pgrt_light {
    float position[3];
    float ambient[3];
    float diffuse[3];
    float specular[3];
} pgrt_light;

/* Everything get_pixel / trace hard-code (raytracer.cpp:398-400, :282, :450); defaults = pgrt_default_params. */
typedef struct pgrt_render_params {
    int32_t sampling_width;   /* 3   stratified grid per pixel, S = w*w samples (raytracer.cpp:398)          */
    int32_t jitter;           /* 1   U[-0.5/w, 0.5/w) jitter (raytracer.cpp:408-410); 0 = sample at the cell corner */
    float focal_distance;     /* 200 (raytracer.cpp:399)                                                    */
    float aperture;           /* 5   (raytracer.cpp:400); 0 = thin-lens code path with no lens shift          */
    int32_t max_depth;        /* 7   "level >= max_depth -> black" (raytracer.cpp:282)                       */
    float gamma_level;        /* 0.5 UI slider default (raytracer.cpp:450); pixel = c^(2*gamma_level)         */
    uint32_t seed;            /* counter-based RNG seed (replaces the clock-seeded mt19937, raytracer.cpp:407) */
    int32_t camera_mode;      /* 0 thin lens generate_ray(x,y,f,a) PinHoleCamera.cpp:65-105; 1 pinhole :31-63 */
    int32_t shader_mode;      /* 0 Whitted (trace as shipped); 1 Lambert (diffuse addend of :377 only);
                                 2 normal shader (the commented block raytracer.cpp:274-280);
                                 3 path tracing (README.md:21 "To do"; no reference counterpart): dielectrics as in 0, every other
                                   hit = its Phong value + albedo x the radiance of one cosine-weighted bounce (the environment
                                   map lights the scene); converge with pgrt_render_accumulate.  Bounces count as reflection rays. */
    int32_t scheduler;        /* 0 fused: one persistent kernel runs trace() for every sample, secondary rays in a shared pool;
                                 1 level-synchronous wavefront (one queue per recursion level);
                                 2 hybrid: level 0 as three lean wavefront kernels, every ray of level >= 1 through the pool;
                                 3 automatic (what pgrt_default_params sets): hybrid for batches of PGRT_AUTO_HYBRID_MIN
                                   (400 000) samples or more, fused below -- the faster one on either side.
                                 Same image, bit for bit, whichever runs. */
    int32_t shadow_mode;      /* 0 is_illuminated as shipped (shadow rays that leave the light towards the hit POSITION,
                                 raytracer.cpp:150-176, LightSource.cpp:11-32: the parity contract); 1 the hard shadows of the
                                 README's to-do list (README.md:20): hit point -> light.  Non-default, no reference counterpart. */
    int32_t reserved[5];
} pgrt_render_params;

typedef struct pgrt_build_stats {
    uint32_t triangles;
    uint32_t nodes;           /* 8-wide nodes emitted (80 B each)                                */
    float build_ms;           /* device time of the whole commit (Morton .. collapse)            */
    float sort_ms;            /* radix sort of the Morton keys                                   */
    float sah_cost;           /* SAH cost of the emitted wide tree: sum(area(node)) + sum(area(leaf slot) * triangles),
                                 over area(root)                                                  */
    float tree_ms;            /* binary tree over the Morton order (PLOC passes)                 */
    float collapse_ms;        /* collapse to the wide layout + triangle re-layout                */
    uint32_t depth;           /* levels of the wide tree                                         */
    uint32_t ploc_passes;     /* multi-block PLOC passes (the last <= 512 clusters finish in one block) */
    uint32_t node_bytes;      /* 80 = quantised planes, 240 = float planes (chosen by scene size; PGRT_NODE_LAYOUT=q8|f32) */
    uint32_t reserved[2];
} pgrt_build_stats;

typedef struct pgrt_render_stats {
    uint64_t rays_primary;    /* one per get_ray_hit call, by caller (raytracer.cpp:241, :164, :300, :310) */
    uint64_t rays_shadow;
    uint64_t rays_reflection;
    uint64_t rays_refraction;
    uint64_t nodes_visited;   /* BVH nodes fetched / triangles tested over all closest-hit queries (profile & 2 renders only) */
    uint64_t tris_tested;
    float frame_ms;           /* device time, first launch .. framebuffer ready                 */
    float trace_ms;           /* sum of closest-hit traversal kernel launches (profiled renders only) */
    float shade_ms;           /* sum of the other kernels (profiled renders only)               */
    uint32_t trace_launches;
    uint32_t launches;        /* kernels launched for this frame                                 */
    uint32_t batches;         /* sample batches the frame was cut into                           */
    uint32_t overflow_retries;
    uint32_t max_nodes_per_ray;   /* worst single query (profile & 2 renders only)             */
    uint32_t reserved[4];     /* [0] pool records of the largest batch; fused scheduler, last batch: [1] frame kernel first warp in ..
                                 last warp out (us), [2] .. until the primary rays ran out (us), [3] warp iterations spent on secondary rays */
} pgrt_render_stats;

/* One recursion level of trace() (raytracer.cpp:237, `level`) in the last rendered frame. */
typedef struct pgrt_level_stats {
    uint64_t rays;            /* closest-hit queries issued by trace() at this level (level 0 = primary)   */
    uint64_t shadow_rays;     /* is_illuminated queries issued by Phong hits of this level                 */
    uint64_t nodes, tris;     /* BVH nodes fetched / triangles tested by `rays` (profile bit 1 only)       */
    uint64_t shadow_nodes, shadow_tris;
    uint32_t max_nodes, shadow_max_nodes;   /* worst single query                                            */
    float trace_ms, shade_ms; /* profile bit 0 only                                                         */
} pgrt_level_stats;

/* Layout-compatible with RTCRayHit (embree3/rtcore_ray.h:11-49), 80 bytes. */
typedef struct pgrt_rayhit {
    float org_x, org_y, org_z, tnear;
    float dir_x, dir_y, dir_z, time;
    float tfar;
    uint32_t mask, id, flags;
    float Ng_x, Ng_y, Ng_z, u, v;
    uint32_t primID, geomID, instID;
} pgrt_rayhit;

/* Layout-compatible with RTCRay (embree3/rtcore_ray.h:11-27), 48 bytes. */
typedef struct pgrt_ray {
    float org_x, org_y, org_z, tnear;
    float dir_x, dir_y, dir_z, time;   /* time carries the IOR of the medium the ray travels in (raytracer.cpp:200,228,416) */
    float tfar;
    uint32_t mask, id, flags;
} pgrt_ray;

/* ---- lifetime: replaces Raytracer::InitDeviceAndScene / ReleaseDeviceAndScene (raytracer.cpp:26-46),
 *      i.e. rtcNewDevice + rtcNewScene / rtcReleaseScene + rtcReleaseDevice.  `device` = CUDA ordinal. */
int pgrt_create(pgrt_context** ctx, int device);
void pgrt_destroy(pgrt_context* ctx);
/* replaces the rtcSetDeviceErrorFunction callback (raytracer.cpp:29-30, tutorials.cpp:9-26): last message, never NULL */
const char* pgrt_last_error(const pgrt_context* ctx);
/* optional: run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL restores the context's own */
int pgrt_set_stream(pgrt_context* ctx, void* cuda_stream);

/* ---- scene upload: replaces the per-surface block of Raytracer::LoadScene (raytracer.cpp:71-125):
 *      rtcNewGeometry + 4x rtcSetNewGeometryBuffer + rtcSetGeometryUserData + rtcCommitGeometry + rtcAttachGeometry.
 *      pos / nrm: 9 floats per triangle (3 un-indexed corners, :105-111), uv: 6 floats per triangle (:113-114);
 *      host pointers, borrowed for the call only.  *geom_id receives what rtcAttachGeometry would return (:123). */
int pgrt_add_mesh(pgrt_context* ctx, const float* pos, const float* nrm, const float* uv, uint32_t n_triangles,
                  int32_t material_id, uint32_t* geom_id);
int pgrt_set_materials(pgrt_context* ctx, const pgrt_material* materials, int32_t n);
/* raw top-down BGR(A) bytes exactly as Texture::Texture leaves them (texture.cpp:36-47): pitch bytes per row, bpp 3 or 4 */
int pgrt_set_texture(pgrt_context* ctx, int32_t id, const uint8_t* bytes, int32_t width, int32_t height, int32_t pitch, int32_t bpp);
/* SphericalMap's texture (SphericalMap.cpp:12-15) */
int pgrt_set_envmap(pgrt_context* ctx, const uint8_t* bytes, int32_t width, int32_t height, int32_t pitch, int32_t bpp);
int pgrt_set_lights(pgrt_context* ctx, const pgrt_light* lights, int32_t n);   /* lights_ (raytracer.cpp:66-68) */
/* replaces rtcCommitScene (raytracer.cpp:127): GPU BVH build over everything added so far */
int pgrt_commit(pgrt_context* ctx, pgrt_build_stats* stats);
int pgrt_clear_scene(pgrt_context* ctx);

/* ---- camera: replaces PinHoleCamera::PinHoleCamera (PinHoleCamera.cpp:5-29), up = +Z (PinHoleCamera.h:36) */
int pgrt_set_camera(pgrt_context* ctx, int32_t width, int32_t height, float fov_y, const float from[3], const float at[3]);
void pgrt_default_params(pgrt_render_params* p);

/* ---- the path: one Producer iteration (simpleguidx11.cpp:95-118).
 *      rgba = width*height*4 floats, pixel (x,y) at (y*width+x)*4, row 0 = top, a = 1 (simpleguidx11.cpp:108-114).
 *      pgrt_render          : host destination (device->host copy inside the call, as memcpy :121-124).
 *      pgrt_render_device   : device destination, asynchronous on the context stream; stats may be NULL.
 *      `profile` bit 0 brackets every launch with events to fill trace_ms / shade_ms (serialises nothing);
 *      bit 1 runs the instrumented traversal that counts nodes / triangles per query (slower; same image). */
int pgrt_render(pgrt_context* ctx, const pgrt_render_params* p, float* rgba_host, pgrt_render_stats* stats, int32_t profile);
int pgrt_render_device(pgrt_context* ctx, const pgrt_render_params* p, void* rgba_device, pgrt_render_stats* stats, int32_t profile);
/* the same frame as R8G8B8A8_UNORM, the format the reference's swap chain presents (simpleguidx11.cpp:229; the float texture is
 * converted on the way to the back buffer, :290): byte = round(clamp(c,0,1)*255), NaN -> 0; pixel (x,y) at (y*width+x)*4, a = 255.
 * A quarter of the bytes of the float frame when the frame has to reach host memory. */
int pgrt_render_rgba8(pgrt_context* ctx, const pgrt_render_params* p, uint8_t* rgba8_host, pgrt_render_stats* stats, int32_t profile);
/* ---- pipelined frames.  Producer renders frames forever (simpleguidx11.cpp:95-125); a context can keep up to
 *      PGRT_MAX_INFLIGHT of them in flight, each in its own slot (own CUDA stream, queues and counters), so that the
 *      latency-bound tail of one frame and its device->host copy overlap the next frame's primary rays.
 *      *_begin enqueues the whole frame and returns without waiting for the GPU (host destinations should be pinned);
 *      pgrt_render_end waits for that slot, handles the rare queue-overflow retry and fills the stats.
 *      pgrt_render / pgrt_render_device / pgrt_render_shard_device are begin + end on slot 0.
 *      The scene, camera and shard may only change while no frame is in flight. */
int pgrt_render_begin(pgrt_context* ctx, const pgrt_render_params* p, float* rgba_host, int32_t slot, int32_t profile);
int pgrt_render_device_begin(pgrt_context* ctx, const pgrt_render_params* p, void* rgba_device, int32_t slot, int32_t profile);
int pgrt_render_shard_device_begin(pgrt_context* ctx, const pgrt_render_params* p, void* shard_rgba_device, int32_t slot, int32_t profile);
/* sharded context, full-frame destination: this rank's tiles are stored at their final place (y*width + x) of a frame
 * that may live on ANOTHER GPU (a peer-mapped pointer: CUDA IPC / NVLink).  The resolve kernel then IS the gather: no
 * collective moves pixels; the caller only needs a completion barrier before rank 0 reads the frame. */
/* frames shared between the processes of one box: rank 0 allocates and exports, the others import (cudaIpc*; the
 * mapping is opened from the importing context's own device, so its kernels reach the memory over NVLink) */
int pgrt_frame_alloc(pgrt_context* ctx, uint64_t bytes, void** device_ptr);
int pgrt_frame_free(pgrt_context* ctx, void* device_ptr);
int pgrt_frame_export(pgrt_context* ctx, void* device_ptr, uint8_t handle[64]);
int pgrt_frame_import(pgrt_context* ctx, const uint8_t handle[64], void** device_ptr);
int pgrt_frame_unmap(pgrt_context* ctx, void* device_ptr);
/* host frames shared between the processes of one box: `host` is page-aligned host memory that every rank has mapped
 * (a memfd / POSIX shared-memory mapping).  Registered with this context's device, *device_ptr is the address its
 * kernels store through: passed as `frame_device` to pgrt_render_shard_to_frame_begin, every rank's resolve kernel writes
 * its own tiles into the one host frame through its own PCIe link (zero-copy; no device-to-host copy on rank 0). */
int pgrt_host_frame_register(pgrt_context* ctx, void* host, uint64_t bytes, void** device_ptr);
int pgrt_host_frame_unregister(pgrt_context* ctx, void* host);
int pgrt_enable_peer_access(pgrt_context* ctx, int32_t peer_device);   /* let this context's kernels store into memory of `peer_device` */
int pgrt_render_shard_to_frame_begin(pgrt_context* ctx, const pgrt_render_params* p, void* frame_device, int32_t slot, int32_t profile);
int pgrt_render_rgba8_begin(pgrt_context* ctx, const pgrt_render_params* p, uint8_t* rgba8_host, int32_t slot, int32_t profile);   /* pgrt_render_rgba8, pipelined */
int pgrt_render_shard_to_frame_rgba8_begin(pgrt_context* ctx, const pgrt_render_params* p, void* frame_rgba8_device, int32_t slot, int32_t profile);
int pgrt_render_end(pgrt_context* ctx, int32_t slot, pgrt_render_stats* stats);
void* pgrt_slot_stream(pgrt_context* ctx, int32_t slot);                          /* cudaStream_t of a slot (slot 0 = pgrt_set_stream's) */
/* make `cuda_stream` wait for the frames begun on the slot so far.  The wait is on the slot's completion count, which only a
 * frame that finished WITHOUT a secondary-ray queue overflow advances: a consumer ordered here is never released by an attempt
 * that pgrt_render_end is still going to repeat with larger queues. */
int pgrt_stream_wait_slot(pgrt_context* ctx, int32_t slot, void* cuda_stream);
/* completion flags between the processes of one box, replacing a per-frame collective: the frames begun on `slot` from now on
 * store `value` in *flag_device (own memory, a peer-mapped word, registered host memory) when they finish without a queue
 * overflow; pgrt_stream_wait_value32 makes a stream wait until *flag_device >= value (cuStreamWaitValue32),
 * pgrt_stream_write_value32 stores a value in stream order (the "frame consumed" signal in the other direction). */
int pgrt_slot_signal(pgrt_context* ctx, int32_t slot, void* flag_device, uint32_t value);
/* the same with a counter shared by all ranks: a finished frame adds 1 to *counter_device (system-scope atomic; the counter may live on
 * another GPU behind NVLink), so the consumer waits once per frame for n_ranks * frames instead of once per rank */
int pgrt_slot_signal_add(pgrt_context* ctx, int32_t slot, void* counter_device);
int pgrt_stream_wait_value32(pgrt_context* ctx, void* cuda_stream, void* flag_device, uint32_t value);
int pgrt_stream_write_value32(pgrt_context* ctx, void* cuda_stream, void* flag_device, uint32_t value);
/* cross-frame accumulation (the step after the path; the reference's Producer re-renders from scratch every iteration,
 * simpleguidx11.cpp:95-118): n_frames finished frames with seeds p->seed, p->seed + 1, ... are summed on the device in
 * frame order (float) and divided by n_frames; rgba_host as pgrt_render.  stats = totals over the frames. */
int pgrt_render_accumulate(pgrt_context* ctx, const pgrt_render_params* p, int32_t n_frames, float* rgba_host, pgrt_render_stats* stats);
/* single-pixel hook with the signature of SimpleGuiDX11::get_pixel (simpleguidx11.h:27): serves pixel (x,y) of
 * the frame rendered by the last pgrt_render* call (re-renders when the camera, scene or params changed). */
int pgrt_get_pixel(pgrt_context* ctx, const pgrt_render_params* p, int32_t x, int32_t y, float rgba[4]);
/* primary hit of sample (0,0) of every pixel: what ray_hit.hit.geomID / primID hold at raytracer.cpp:247 */
int pgrt_primary_ids(pgrt_context* ctx, const pgrt_render_params* p, uint32_t* geom_id_host, uint32_t* prim_id_host);

/* ---- image-tile sharding (one process per GPU): this context renders tiles t with t % n_ranks == rank of the
 *      32x8-pixel tile grid.  pgrt_render_shard_device writes the compact per-rank buffer
 *      (pgrt_shard_pixels * 4 floats); pgrt_untile scatters n_ranks such buffers (concatenated, rank-major, each
 *      pgrt_shard_pixels long) into the full frame.  No data-path collective lives here: the gather between is NCCL.
 *      pgrt_render_shard_to_frame_begin (below) skips both: tiles go straight into rank 0's frame over NVLink. */
int pgrt_set_shard(pgrt_context* ctx, int32_t rank, int32_t n_ranks);
uint64_t pgrt_shard_pixels(const pgrt_context* ctx);   /* padded pixel slots per rank (same on every rank) */
int pgrt_render_shard_device(pgrt_context* ctx, const pgrt_render_params* p, void* shard_rgba_device, pgrt_render_stats* stats, int32_t profile);
int pgrt_untile(pgrt_context* ctx, const void* gathered_device, int32_t n_ranks, void* rgba_device);
int pgrt_untile_on_stream(pgrt_context* ctx, const void* gathered_device, int32_t n_ranks, void* rgba_device, void* cuda_stream);   /* same, on the stream the gather ran on */

/* ---- rtcIntersect1 over a batch (raytracer.cpp:130-148; semantics emb/doc/README.md:6331-6415):
 *      closest hit in (tnear, tfar]; on hit writes tfar, u, v, Ng, primID, geomID; a miss leaves the record
 *      untouched.  Host array of n RTCRayHit-compatible records. */
/* Ray directions: a component whose magnitude is below 2^-80 is treated as 2^-80 of the same sign when the reciprocal
 * is formed (Embree clamps in the same way, at a larger threshold); a direction ALL of whose components are that small is outside
 * the supported domain. */
int pgrt_intersect(pgrt_context* ctx, pgrt_rayhit* rayhits_host, uint64_t n);
/* rtcInterpolate0 (raytracer.cpp:252, :344): slot 0 -> 3 floats (normal), slot 1 -> 2 floats (uv), per query */
int pgrt_interpolate(pgrt_context* ctx, const uint32_t* geom_id, const uint32_t* prim_id, const float* u, const float* v,
                     uint64_t n, int32_t slot, float* out_host);

/* ---- the two public query methods of the class (raytracer.h:31, :34) over batches.
 *      pgrt_trace: Color4f Raytracer::trace(RTCRay ray, int level) (raytracer.cpp:237-394) for n caller-supplied rays, all at
 *      recursion level `level`; rgba_host receives n x 4 floats.  Only shader_mode, max_depth, shadow_mode and seed of p apply.
 *      pgrt_is_illuminated: bool Raytracer::is_illuminated(LightSource, Vector3 hit_position, Vector3 normal) (:150-176) for n
 *      (light position, hit position, normal) triples of 3 floats each; lit_host receives 0 / 1. */
int pgrt_trace(pgrt_context* ctx, const pgrt_render_params* p, const pgrt_ray* rays_host, uint64_t n, int32_t level, float* rgba_host);
int pgrt_is_illuminated(pgrt_context* ctx, const pgrt_render_params* p, const float* light_pos3, const float* hit_pos3, const float* normal3,
                        uint64_t n, int32_t* lit_host);

/* ---- TESTING: device implementations of the path's leaf functions, exposed for per-function parity tests (not part of the
 *      drop-in surface; a binding may ignore everything from here to "introspection") */
int pgrt_eval_mix_srgb(pgrt_context* ctx, const float* c0, const float* c1, const float* alpha, uint64_t n, float* out);  /* utils.cpp:238-241 */
int pgrt_eval_texture(pgrt_context* ctx, int32_t tex_id, const float* uv, uint64_t n, float* out3);  /* Texture::get_texel texture.cpp:77-130; id -1 = env texture */
int pgrt_eval_envmap(pgrt_context* ctx, const float* dirs, uint64_t n, float* out4);                 /* SphericalMap::get_texel SphericalMap.cpp:17-29 */
int pgrt_eval_gamma(pgrt_context* ctx, const float* in4, float gamma_level, uint64_t n, float* out4); /* Raytracer::gamma raytracer.cpp:439-446 */
int pgrt_eval_primary_rays(pgrt_context* ctx, const pgrt_render_params* p, float* out9);             /* get_pixel :405-416 + generate_ray; 9 floats/ray */
int pgrt_eval_secondary_rays(pgrt_context* ctx, const float* in11, uint64_t n, int32_t refraction, float* out9);  /* raytracer.cpp:178-235 */

/* ---- TESTING / measurement helper: stream `bytes` of writes through a scratch buffer on the slot's stream (evicts the 126 MB L2
 *      when bytes exceeds it); benchmarks call it before every timed frame */
int pgrt_debug_flush_l2(pgrt_context* ctx, int32_t slot, uint64_t bytes, uint32_t value);
/* read bandwidth (GB/s) of a `bytes`-sized buffer that stays L2-resident, from all SMs with L1 bypassed, `iters` passes:
 * the memory roof that applies while a scene's nodes and triangles fit L2 */
int pgrt_debug_l2_bandwidth(pgrt_context* ctx, uint64_t bytes, int32_t iters, float* gb_per_s);
/* builds with -DPGRT_FRAME_TIMING only (zeros otherwise): SM cycles the warps of the slot's last frame kernel spent on
 * [0] primary chunks, [1] pool records (secondary rays), [2] resident in total; [3] = warps that ran */
int pgrt_debug_frame_cycles(pgrt_context* ctx, int32_t slot, uint64_t out[4]);

/* ---- introspection */
uint32_t pgrt_num_triangles(const pgrt_context* ctx);
uint32_t pgrt_num_geometries(const pgrt_context* ctx);
int pgrt_last_level_stats(const pgrt_context* ctx, int32_t level, pgrt_level_stats* out);   /* levels 0 .. max_depth of the last frame */
uint64_t pgrt_kernel_launches(const pgrt_context* ctx);   /* kernels launched by this context since creation */
const char* pgrt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PGRT_H_ */
