// pgrt_api.cu -- C-ABI of libpgrt_b200.so (include/pgrt.h): context, scene upload, GPU BVH commit, frame orchestration.
// No CPU fallback lives here: every entry point that computes launches CUDA kernels or fails.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <mutex>
#include "common.cuh"
#include "bvh_build.cuh"
#include "render.cuh"

namespace {

// cudaFree synchronises the whole device.  A buffer that grows while frames are in flight (a queue-overflow retry inside
// pgrt_render_end, the first frame of another slot) must not do that: a consumer stream may be parked on a completion flag
// (pgrt_stream_wait_slot, dist.ShardedRenderer) that only the work we are about to enqueue will set, and the free would wait
// for that stream for ever.  Outgrown buffers go to a graveyard that is emptied where everything is synchronised anyway
// (pgrt_commit, pgrt_destroy).
std::mutex g_graveyard_lock;
std::vector<void*> g_graveyard;
void deferred_free(void* p) { std::lock_guard<std::mutex> g(g_graveyard_lock); g_graveyard.push_back(p); }
void drain_graveyard() {
    std::vector<void*> v;
    { std::lock_guard<std::mutex> g(g_graveyard_lock); v.swap(g_graveyard); }
    for (void* p : v) cudaFree(p);
}

template <typename T>
struct DevBuf {
    T* p = nullptr; size_t n = 0;
    cudaError_t ensure(size_t count) {
        if (count <= n) return cudaSuccess;
        if (p) deferred_free(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

struct HostTex { DevBuf<uint8_t> bytes; int w = 0, h = 0, pitch = 0, bpp = 0; bool set = false; };

enum KClass { KC_TRACE = 0, KC_SHADE = 1 };

// everything one frame in flight owns
struct FrameSlot {
    cudaStream_t stream = nullptr, own_stream = nullptr;
    LevelBufs levels[PGRT_MAX_LEVELS + 1] = {};
    DevBuf<float4> lv_f4[PGRT_MAX_LEVELS + 1][5];
    DevBuf<uint2> lv_child[PGRT_MAX_LEVELS + 1];
    DevBuf<uint32_t> lv_list[PGRT_MAX_LEVELS + 1][2];
    DevBuf<uint32_t> l0_pending;
    RayPool pool = {};                // every ray of level >= 1 (fused scheduler)
    DevBuf<float4> pool_f4[4]; DevBuf<uint2> pool_u2[1]; DevBuf<uint4> pool_u4[1]; DevBuf<uint32_t> pool_u32[1];
    DevBuf<Counters> d_counters;
    DevBuf<float4> d_frame;           // device frame behind a host destination
    Counters* h_counters = nullptr;   // pinned
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_class, ev_level;
    size_t ev_used = 0;
    cudaEvent_t ev_frame0 = nullptr, ev_frame1 = nullptr, ev_done = nullptr;
    int frame_grid = 0, keep_ctas = 0; // CTAs of this frame's k_frame, and how many of them stay until the batch is done (sized in frame_begin)
    // the frame in flight
    bool busy = false;
    pgrt_render_params params = {};
    void* dest = nullptr; int dest_mode = 0; int fmt = 0; void* host_dst = nullptr; int profile = 0;   // fmt: 0 float4, 1 R8G8B8A8_UNORM
    uint32_t* sig_flag = nullptr; uint32_t sig_value = 0; bool sig_add = false;   // external completion flag (pgrt_slot_signal / pgrt_slot_signal_add)
    uint32_t seq_expected = 0;        // frames begun on this slot = value of Counters::done_seq once the last of them has finished
    uint64_t batch_slots = 0; int n_levels = 0; bool fused = false, hybrid = false;   // hybrid: level 0 as wavefront kernels, levels >= 1 through the pool (k_frame)
    pgrt_render_stats rs = {};
    // the frame as a CUDA graph (re-used while nothing it depends on changes)
    cudaGraphExec_t gexec = nullptr; uint64_t gkey = 0, gseen = 0; uint32_t g_launches = 0, g_trace_launches = 0, g_batches = 0;

    void release() {
        if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; }
        for (int l = 0; l <= PGRT_MAX_LEVELS; ++l) {
            for (int k = 0; k < 5; ++k) lv_f4[l][k].release();
            lv_child[l].release(); lv_list[l][0].release(); lv_list[l][1].release();
        }
        l0_pending.release();
        for (auto& b : pool_f4) b.release(); for (auto& b : pool_u2) b.release(); for (auto& b : pool_u4) b.release(); for (auto& b : pool_u32) b.release();
        d_counters.release(); d_frame.release();
        for (auto e : ev_pool) cudaEventDestroy(e);
        ev_pool.clear();
        if (ev_frame0) cudaEventDestroy(ev_frame0);
        if (ev_frame1) cudaEventDestroy(ev_frame1);
        if (ev_done) cudaEventDestroy(ev_done);
        if (h_counters) cudaFreeHost(h_counters);
        if (own_stream) cudaStreamDestroy(own_stream);
        ev_frame0 = ev_frame1 = ev_done = nullptr; h_counters = nullptr; own_stream = stream = nullptr;
    }
};

}  // namespace

struct pgrt_context {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;    // = slots[0].stream: scene uploads, commit, eval entry points, synchronous renders
    std::string err = "";
    uint64_t launches = 0;

    // host staging of the scene (what LoadScene hands over, pg1/raytracer.cpp:71-125)
    std::vector<uint32_t> h_geom_first;
    uint32_t n_added = 0;             // triangles uploaded by pgrt_add_mesh since pgrt_clear_scene
    // pgrt_add_mesh streams the caller's arrays to the device as they arrive: two pinned staging buffers, the copy of one
    // chunk to pinned memory overlaps the DMA of the previous one; nothing of the geometry stays on the host
    uint8_t* stage[2] = {nullptr, nullptr}; cudaEvent_t stage_done[2] = {nullptr, nullptr}; int stage_next = 0;
    std::vector<int32_t> h_geom_material;
    std::vector<pgrt_material> h_materials;
    std::vector<pgrt_light> h_lights;
    std::vector<HostTex> textures;
    HostTex env;

    // device scene
    DevBuf<float> d_pos, d_nrm, d_uv;
    DevBuf<uint32_t> d_tri_geom, d_geom_first;
    DevBuf<int32_t> d_geom_material;
    DevBuf<pgrt_material> d_materials;
    DevBuf<pgrt_light> d_lights;
    DevBuf<DevTexture> d_textures;
    DevBuf<float4> d_shade, d_tris, d_nodes;
    uint32_t n_tris = 0;
    float bb_lo[3] = {0.f, 0.f, 0.f}, bb_hi[3] = {0.f, 0.f, 0.f};     // scene bounds (root of the binary tree), set by pgrt_commit
    int node_layout = PGRT_LAYOUT_Q8, loop_ww = 0;
    bool committed = false, tables_dirty = true;
    pgrt_build_stats last_build = {};

    DevCamera cam = {};
    bool cam_set = false;
    ShardInfo shard = {0, 1, 0, 0};

    // frame state: PGRT_MAX_INFLIGHT independent frame slots, each with its own stream, queues and counters, so the
    // latency-bound tail of one frame (k_secondary) and its device->host copy overlap the next frame's primary work
    FrameSlot slots[PGRT_MAX_INFLIGHT];
    int frame_per_sm_max = 0, frame_per_sm_env = 0;   // occupancy bound of k_frame; PGRT_FRAME_CTAS_PER_SM
    int min_claim = 32;               // k_frame takes pool records ahead of primary rays once this many wait (PGRT_MIN_CLAIM, 1..32)
    int keep_ctas = 2;                // CTAs of k_frame that stay until the batch is done (PGRT_KEEP_CTAS)
    int pool_policy = 0;              // how k_frame takes secondary rays while primary rays last: 0 not at all, 1 full batches by compare-and-swap, 2 tickets (PGRT_POOL_POLICY)
    int claim_patience = 4;           // polls after which an idle warp of k_frame halves the number of pool records it waits for (PGRT_CLAIM_PATIENCE)
    double pool_scale = 1.0;          // grown after a pool overflow; never shrinks before the next pgrt_commit
    // driver entry points for stream memory operations (completion flags without a collective); null = not available
    void* fn_wait32 = nullptr; void* fn_write32 = nullptr;
    bool fuse_raygen = true;          // PGRT_FUSE_RAYGEN=0 restores the stored level-0 ray queue (k_raygen)
    bool use_graphs = true;           // PGRT_GRAPHS=0: every frame as individual launches
    int trace_refill = 32;            // k_trace claims new rays once this many lanes of a warp are idle (PGRT_TRACE_REFILL, 1..32); 32 = whole-warp chunks, the fastest for coherent primary rays (profiles/r1_sweep_trace_refill.txt)
    int trace_ctas_per_sm = 6;        // persistent k_trace grid (PGRT_TRACE_CTAS_PER_SM)
    unsigned trace_rays_per_cta = 1024;   // ... but no more CTAs than one per this many rays of level 0 (PGRT_TRACE_RAYS_PER_CTA)
    DevBuf<uint32_t> d_ids;
    DevBuf<uint4> flush_buf;          // pgrt_debug_flush_l2
    DevBuf<float4> acc_sum, acc_frames[4]; cudaEvent_t acc_event = nullptr;   // pgrt_render_accumulate
    uint32_t* h_pin = nullptr;        // pinned scratch for small read-backs of the build
    size_t max_batch_samples = (size_t)1 << 23;
    uint64_t batch_limit = 0;         // pixel slots per batch that last survived a queue overflow (0 = none yet); reset by pgrt_commit
    size_t auto_hybrid_min = 400000;  // scheduler 3: batches of this many samples or more run the hybrid scheduler (PGRT_AUTO_HYBRID_MIN)
    size_t min_level_cap = (size_t)1 << 18;
    double level_cap_factor = 0.5;   // queue capacity of the deeper levels per primary sample (x4 for the one pool of the dynamic scheduler); overflow = retry in smaller batches
    pgrt_level_stats level_stats[PGRT_MAX_LEVELS + 1] = {};
    int level_stats_n = 0;

    // get_pixel cache
    std::vector<float> px_cache; pgrt_render_params px_params = {}; bool px_valid = false;

    int fail(int code, const std::string& msg) { err = msg; return code; }
    int fail_cuda(cudaError_t e, const char* what, const char* file, int line) {
        char buf[512];
        snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
        err = buf;
        return PGRT_ERR_CUDA;
    }
    DevScene dev_scene() const {
        DevScene s = {};
        s.nodes = d_nodes.p; s.tris = d_tris.p; s.shade = d_shade.p; s.geom_first = d_geom_first.p; s.geom_material = d_geom_material.p;
        s.materials = d_materials.p; s.textures = d_textures.p; s.n_textures = (int32_t)textures.size();
        s.env.data = env.set ? env.bytes.p : nullptr; s.env.width = env.w; s.env.height = env.h; s.env.pitch = env.pitch; s.env.bpp = env.bpp;
        s.lights = d_lights.p; s.n_lights = (int32_t)h_lights.size(); s.n_tris = n_tris; s.node_layout = node_layout; s.loop_ww = loop_ww;
        for (int a = 0; a < 3; ++a) { s.bb_lo[a] = bb_lo[a]; s.bb_hi[a] = bb_hi[a]; }
        return s;
    }
};

#define CHECK_CTX(ctx) do { if (!(ctx)) return PGRT_ERR_INVALID; } while (0)
#define LAUNCH_OK() CUDA_TRY(cudaGetLastError())

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// --------------------------------------------------------------------------------------------------- lifetime
extern "C" const char* pgrt_version(void) { return "pgrt-b200 0.1 (sm_100a)"; }

extern "C" void pgrt_destroy(pgrt_context* ctx);

extern "C" int pgrt_create(pgrt_context** out, int device) {
    if (!out) return PGRT_ERR_INVALID;
    *out = nullptr;
    // one hardware channel per frame-slot stream: with the default of 8, streams share channels and one that waits (completion
    // flags, events) holds back its channel-mates.  Read when the CUDA context is created: a no-op if the process has one already.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return PGRT_ERR_NO_DEVICE;
    pgrt_context* ctx = new pgrt_context();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return PGRT_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    for (FrameSlot& S : ctx->slots) {
        bool ok = cudaStreamCreateWithFlags(&S.own_stream, cudaStreamNonBlocking) == cudaSuccess;
        S.stream = S.own_stream;
        ok = ok && cudaEventCreate(&S.ev_frame0) == cudaSuccess && cudaEventCreate(&S.ev_frame1) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&S.ev_done, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaMallocHost((void**)&S.h_counters, sizeof(Counters)) == cudaSuccess && S.d_counters.ensure(1) == cudaSuccess;
        ok = ok && cudaMemset(S.d_counters.p, 0, sizeof(Counters)) == cudaSuccess;
        if (!ok) { pgrt_destroy(ctx); return PGRT_ERR_CUDA; }
    }
    ctx->stream = ctx->slots[0].stream;
    if (cudaMallocHost((void**)&ctx->h_pin, 256) != cudaSuccess) { pgrt_destroy(ctx); return PGRT_ERR_CUDA; }
    if (const char* e = getenv("PGRT_FUSE_RAYGEN")) ctx->fuse_raygen = atoi(e) != 0;
    if (const char* e = getenv("PGRT_GRAPHS")) ctx->use_graphs = atoi(e) != 0;
    if (const char* e = getenv("PGRT_TRACE_REFILL")) ctx->trace_refill = std::min(32, std::max(1, atoi(e)));
    if (const char* e = getenv("PGRT_TRACE_RAYS_PER_CTA")) ctx->trace_rays_per_cta = (unsigned)std::min(1 << 20, std::max(128, atoi(e)));
    if (const char* e = getenv("PGRT_TRACE_CTAS_PER_SM")) ctx->trace_ctas_per_sm = std::min(16, std::max(1, atoi(e)));
    if (const char* e = getenv("PGRT_MIN_CLAIM")) ctx->min_claim = std::min(32, std::max(1, atoi(e)));
    if (const char* e = getenv("PGRT_KEEP_CTAS")) ctx->keep_ctas = std::min(1 << 16, std::max(1, atoi(e)));
    if (const char* e = getenv("PGRT_POOL_POLICY")) ctx->pool_policy = std::min(2, std::max(0, atoi(e)));
    if (const char* e = getenv("PGRT_CLAIM_PATIENCE")) ctx->claim_patience = std::min(1 << 20, std::max(1, atoi(e)));
    {   // cuStreamWaitValue32 / cuStreamWriteValue32 through the runtime (no link-time dependency on libcuda)
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &ctx->fn_wait32, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) ctx->fn_wait32 = nullptr;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &ctx->fn_write32, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) ctx->fn_write32 = nullptr;
        cudaGetLastError();
    }
    if (const char* e = getenv("PGRT_MAX_BATCH_SAMPLES")) ctx->max_batch_samples = std::max<size_t>(256, strtoull(e, nullptr, 10));
    if (const char* e = getenv("PGRT_AUTO_HYBRID_MIN")) ctx->auto_hybrid_min = (size_t)strtoull(e, nullptr, 10);
    if (const char* e = getenv("PGRT_MIN_LEVEL_CAP")) ctx->min_level_cap = std::max<size_t>(64, strtoull(e, nullptr, 10));
    if (const char* e = getenv("PGRT_LEVEL_CAP_FACTOR")) ctx->level_cap_factor = std::max(0.01, atof(e));
    *out = ctx;
    return PGRT_OK;
}

static void free_scene_device(pgrt_context* ctx) {
    ctx->d_pos.release(); ctx->d_nrm.release(); ctx->d_uv.release(); ctx->d_tri_geom.release(); ctx->d_geom_first.release();
    ctx->d_geom_material.release(); ctx->d_shade.release(); ctx->d_tris.release(); ctx->d_nodes.release();
}

static void sync_all_slots(pgrt_context* ctx) {
    for (FrameSlot& S : ctx->slots) if (S.stream) cudaStreamSynchronize(S.stream);
}

extern "C" void pgrt_destroy(pgrt_context* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    sync_all_slots(ctx);
    free_scene_device(ctx);
    ctx->d_materials.release(); ctx->d_lights.release(); ctx->d_textures.release();
    for (auto& t : ctx->textures) t.bytes.release();
    ctx->env.bytes.release();
    for (FrameSlot& S : ctx->slots) S.release();
    ctx->d_ids.release(); ctx->flush_buf.release();
    ctx->acc_sum.release(); for (auto& b : ctx->acc_frames) b.release(); if (ctx->acc_event) cudaEventDestroy(ctx->acc_event);
    for (int k = 0; k < 2; ++k) { if (ctx->stage[k]) cudaFreeHost(ctx->stage[k]); if (ctx->stage_done[k]) cudaEventDestroy(ctx->stage_done[k]); }
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    drain_graveyard();
    delete ctx;
}

extern "C" const char* pgrt_last_error(const pgrt_context* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int pgrt_set_stream(pgrt_context* ctx, void* s) {
    CHECK_CTX(ctx);
    cudaSetDevice(ctx->device);
    sync_all_slots(ctx);
    ctx->slots[0].stream = s ? (cudaStream_t)s : ctx->slots[0].own_stream;
    ctx->stream = ctx->slots[0].stream;
    return PGRT_OK;
}

extern "C" void* pgrt_slot_stream(pgrt_context* ctx, int32_t slot) {
    if (!ctx || slot < 0 || slot >= PGRT_MAX_INFLIGHT) return nullptr;
    return (void*)ctx->slots[slot].stream;
}

// --------------------------------------------------------------------------------------------------- scene
extern "C" int pgrt_clear_scene(pgrt_context* ctx) {
    CHECK_CTX(ctx);
    ctx->h_geom_first.clear(); ctx->h_geom_material.clear();
    ctx->n_added = 0; ctx->n_tris = 0; ctx->committed = false; ctx->px_valid = false;
    return PGRT_OK;
}

#define PGRT_STAGE_BYTES ((size_t)32 << 20)

__global__ void __launch_bounds__(256) k_fill_u32(uint32_t* __restrict__ dst, uint32_t n, uint32_t v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}

// grows a device array to `count` elements keeping the first `keep` (geometric growth; scene uploads only)
template <typename T>
static cudaError_t grow_keep(DevBuf<T>& b, size_t count, size_t keep, cudaStream_t st) {
    if (count <= b.n) return cudaSuccess;
    const size_t cap = std::max(count, b.n + b.n / 2);
    T* q = nullptr;
    cudaError_t e = cudaMalloc((void**)&q, cap * sizeof(T));
    if (e != cudaSuccess) return e;
    if (keep) e = cudaMemcpyAsync(q, b.p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (b.p) cudaFree(b.p);
    b.p = q; b.n = cap;
    return e;
}

// host array -> device through the pinned ring
static int upload_stream(pgrt_context* ctx, void* dst, const void* src, size_t bytes) {
    cudaStream_t st = ctx->stream;
    for (size_t off = 0; off < bytes; off += PGRT_STAGE_BYTES) {
        const size_t n = std::min(PGRT_STAGE_BYTES, bytes - off);
        const int k = ctx->stage_next; ctx->stage_next ^= 1;
        if (!ctx->stage[k]) {
            CUDA_TRY(cudaMallocHost((void**)&ctx->stage[k], PGRT_STAGE_BYTES));
            CUDA_TRY(cudaEventCreateWithFlags(&ctx->stage_done[k], cudaEventDisableTiming));
        } else CUDA_TRY(cudaEventSynchronize(ctx->stage_done[k]));
        memcpy(ctx->stage[k], (const uint8_t*)src + off, n);
        CUDA_TRY(cudaMemcpyAsync((uint8_t*)dst + off, ctx->stage[k], n, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaEventRecord(ctx->stage_done[k], st));
    }
    return PGRT_OK;
}

extern "C" int pgrt_add_mesh(pgrt_context* ctx, const float* pos, const float* nrm, const float* uv, uint32_t T, int32_t material_id, uint32_t* geom_id) {
    CHECK_CTX(ctx);
    if (T && (!pos || !nrm || !uv)) return ctx->fail(PGRT_ERR_INVALID, "pgrt_add_mesh: null buffer");
    if ((uint64_t)ctx->n_added + T >= (1ull << 29)) return ctx->fail(PGRT_ERR_INVALID, "pgrt_add_mesh: more than 2^29 triangles");
    cudaSetDevice(ctx->device);
    if (ctx->committed || ctx->n_added == 0) sync_all_slots(ctx);   // frames in flight still read the arrays that may move
    const uint32_t g = (uint32_t)ctx->h_geom_first.size();
    const size_t first = ctx->n_added, total = first + T;
    cudaStream_t st = ctx->stream;
    if (T) {
        CUDA_TRY(grow_keep(ctx->d_pos, 9 * total, 9 * first, st)); CUDA_TRY(grow_keep(ctx->d_nrm, 9 * total, 9 * first, st));
        CUDA_TRY(grow_keep(ctx->d_uv, 6 * total, 6 * first, st)); CUDA_TRY(grow_keep(ctx->d_tri_geom, total, first, st));
        int rc = upload_stream(ctx, ctx->d_pos.p + 9 * first, pos, 9 * (size_t)T * 4); if (rc) return rc;
        rc = upload_stream(ctx, ctx->d_nrm.p + 9 * first, nrm, 9 * (size_t)T * 4); if (rc) return rc;
        rc = upload_stream(ctx, ctx->d_uv.p + 6 * first, uv, 6 * (size_t)T * 4); if (rc) return rc;
        k_fill_u32<<<div_up(T, 256), 256, 0, st>>>(ctx->d_tri_geom.p + first, T, g);
        ctx->launches++;
        LAUNCH_OK();
    }
    ctx->h_geom_first.push_back((uint32_t)first);
    ctx->h_geom_material.push_back(material_id);
    ctx->n_added = (uint32_t)total;
    if (geom_id) *geom_id = g;
    ctx->committed = false; ctx->px_valid = false;
    return PGRT_OK;
}

extern "C" int pgrt_set_materials(pgrt_context* ctx, const pgrt_material* m, int32_t n) {
    CHECK_CTX(ctx);
    if (n < 0 || (n && !m)) return ctx->fail(PGRT_ERR_INVALID, "pgrt_set_materials: bad arguments");
    ctx->h_materials.assign(m, m + n); ctx->tables_dirty = true; ctx->px_valid = false;
    return PGRT_OK;
}

extern "C" int pgrt_set_lights(pgrt_context* ctx, const pgrt_light* l, int32_t n) {
    CHECK_CTX(ctx);
    if (n < 0 || (n && !l)) return ctx->fail(PGRT_ERR_INVALID, "pgrt_set_lights: bad arguments");
    ctx->h_lights.assign(l, l + n); ctx->tables_dirty = true; ctx->px_valid = false;
    return PGRT_OK;
}

static int upload_tex(pgrt_context* ctx, HostTex& t, const uint8_t* bytes, int w, int h, int pitch, int bpp) {
    if (!bytes || w <= 0 || h <= 0 || (bpp != 3 && bpp != 4) || pitch < w * bpp) return ctx->fail(PGRT_ERR_INVALID, "texture: bad arguments");
    cudaSetDevice(ctx->device);
    sync_all_slots(ctx);
    CUDA_TRY(t.bytes.ensure((size_t)pitch * h + 4));
    CUDA_TRY(cudaMemcpyAsync(t.bytes.p, bytes, (size_t)pitch * h, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    t.w = w; t.h = h; t.pitch = pitch; t.bpp = bpp; t.set = true;
    ctx->tables_dirty = true; ctx->px_valid = false;
    return PGRT_OK;
}

extern "C" int pgrt_set_texture(pgrt_context* ctx, int32_t id, const uint8_t* bytes, int32_t w, int32_t h, int32_t pitch, int32_t bpp) {
    CHECK_CTX(ctx);
    if (id < 0 || id > 4096) return ctx->fail(PGRT_ERR_INVALID, "pgrt_set_texture: bad id");
    if ((size_t)id >= ctx->textures.size()) ctx->textures.resize(id + 1);
    return upload_tex(ctx, ctx->textures[id], bytes, w, h, pitch, bpp);
}

extern "C" int pgrt_set_envmap(pgrt_context* ctx, const uint8_t* bytes, int32_t w, int32_t h, int32_t pitch, int32_t bpp) {
    CHECK_CTX(ctx);
    return upload_tex(ctx, ctx->env, bytes, w, h, pitch, bpp);
}

static int upload_tables(pgrt_context* ctx) {
    if (!ctx->tables_dirty) return PGRT_OK;
    cudaSetDevice(ctx->device);
    sync_all_slots(ctx);
    for (const auto& m : ctx->h_geom_material)
        if (m < 0 || (size_t)m >= ctx->h_materials.size()) return ctx->fail(PGRT_ERR_INVALID, "a mesh references a material id that was never set");
    CUDA_TRY(ctx->d_materials.ensure(ctx->h_materials.size()));
    if (!ctx->h_materials.empty()) CUDA_TRY(cudaMemcpyAsync(ctx->d_materials.p, ctx->h_materials.data(), ctx->h_materials.size() * sizeof(pgrt_material), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx->d_lights.ensure(ctx->h_lights.size()));
    if (!ctx->h_lights.empty()) CUDA_TRY(cudaMemcpyAsync(ctx->d_lights.p, ctx->h_lights.data(), ctx->h_lights.size() * sizeof(pgrt_light), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<DevTexture> dt(ctx->textures.size());
    for (size_t i = 0; i < dt.size(); ++i) {
        const HostTex& t = ctx->textures[i];
        dt[i].data = t.set ? t.bytes.p : nullptr; dt[i].width = t.w; dt[i].height = t.h; dt[i].pitch = t.pitch; dt[i].bpp = t.bpp;
    }
    for (const auto& m : ctx->h_materials)
        if (m.diffuse_tex >= 0 && ((size_t)m.diffuse_tex >= dt.size() || !dt[m.diffuse_tex].data)) return ctx->fail(PGRT_ERR_INVALID, "a material references a texture id that was never set");
    CUDA_TRY(ctx->d_textures.ensure(dt.size()));
    if (!dt.empty()) CUDA_TRY(cudaMemcpyAsync(ctx->d_textures.p, dt.data(), dt.size() * sizeof(DevTexture), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->tables_dirty = false;
    return PGRT_OK;
}

// --------------------------------------------------------------------------------------------------- commit (GPU BVH build)
extern "C" int pgrt_commit(pgrt_context* ctx, pgrt_build_stats* stats) {
    CHECK_CTX(ctx);
    cudaSetDevice(ctx->device);
    sync_all_slots(ctx);
    drain_graveyard();
    cudaStream_t st = ctx->stream;
    const uint32_t N = ctx->n_added;   // the geometry is on the device already (pgrt_add_mesh streams it)
    ctx->n_tris = N; ctx->px_valid = false; ctx->batch_limit = 0; ctx->pool_scale = 1.0;
    pgrt_build_stats bs = {}; bs.triangles = N;
    const uint32_t G = (uint32_t)ctx->h_geom_first.size();
    CUDA_TRY(ctx->d_geom_first.ensure(G)); CUDA_TRY(ctx->d_geom_material.ensure(G));
    if (G) {
        CUDA_TRY(cudaMemcpyAsync(ctx->d_geom_first.p, ctx->h_geom_first.data(), G * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(ctx->d_geom_material.p, ctx->h_geom_material.data(), G * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    }
    ctx->tables_dirty = true;
    if (N == 0) {
        ctx->committed = true; ctx->last_build = bs;
        if (stats) *stats = bs;
        CUDA_TRY(cudaStreamSynchronize(st));
        return PGRT_OK;
    }
    // node layout: un-quantised planes while the node array fits L2 (the traversal is issue-bound there and the decode is
    // 30 % of a node visit: C4 with 3.08 M triangles gains 8 %), the 80-B quantised node otherwise; PGRT_NODE_LAYOUT=q8|f32 overrides
    int layout = ((double)N * 0.15 * PGRT_NODE_F4_F32 * 16.0 <= 128e6) ? PGRT_LAYOUT_F32 : PGRT_LAYOUT_Q8;   // about the size of L2
    if (const char* e = getenv("PGRT_NODE_LAYOUT")) layout = !strcmp(e, "f32") ? PGRT_LAYOUT_F32 : (!strcmp(e, "q8") ? PGRT_LAYOUT_Q8 : layout);
    const size_t node_f4 = layout == PGRT_LAYOUT_F32 ? PGRT_NODE_F4_F32 : PGRT_NODE_F4_Q8;
    ctx->node_layout = layout;
    ctx->loop_ww = 0;   // loop shape (traverse.cuh): if-if wins on every workload measured (profiles/r1_matrix_layout_loopshape.txt); PGRT_LOOP=ww|ii overrides
    if (const char* e = getenv("PGRT_LOOP")) ctx->loop_ww = !strcmp(e, "ww") ? 1 : (!strcmp(e, "ii") ? 0 : ctx->loop_ww);
    CUDA_TRY(ctx->d_shade.ensure(4 * (size_t)N)); CUDA_TRY(ctx->d_tris.ensure(3 * (size_t)N)); CUDA_TRY(ctx->d_nodes.ensure(node_f4 * (size_t)N));

    cudaEvent_t e0, e1, e2, e3, e4;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3); cudaEventCreate(&e4);
    // build temporaries
    DevBuf<uint64_t> keys[2]; DevBuf<uint32_t> vals[2]; DevBuf<uint32_t> hist; DevBuf<SceneBounds> sb;
    DevBuf<float4> b0, b1; DevBuf<uint32_t> bcount, cid[2]; DevBuf<int> nn; DevBuf<uint2> bsums, items[2]; DevBuf<CollapseCounters> cc;
    DevBuf<int> ti[6]; DevBuf<float> tf[2];
    const uint32_t n_tiles = div_up(N, RS_TILE);
    const int ploc_ties = getenv("PGRT_PLOC_TIES") ? atoi(getenv("PGRT_PLOC_TIES")) : PLOC_TIES_LOWEST;
    const bool use_lbvh = getenv("PGRT_BUILDER") && !strcmp(getenv("PGRT_BUILDER"), "lbvh");
    int rc = PGRT_OK;
    auto cleanup = [&]() {
        keys[0].release(); keys[1].release(); vals[0].release(); vals[1].release(); hist.release(); sb.release();
        b0.release(); b1.release(); bcount.release(); cid[0].release(); cid[1].release(); nn.release(); bsums.release(); items[0].release(); items[1].release(); cc.release();
        for (auto& b : ti) b.release(); for (auto& b : tf) b.release();
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaEventDestroy(e3); cudaEventDestroy(e4);
    };
#define BUILD_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = ctx->fail_cuda(e__, #call, __FILE__, __LINE__); cleanup(); return rc; } } while (0)
    BUILD_TRY(keys[0].ensure(N)); BUILD_TRY(keys[1].ensure(N)); BUILD_TRY(vals[0].ensure(N)); BUILD_TRY(vals[1].ensure(N));
    BUILD_TRY(hist.ensure(256 * (size_t)n_tiles + 256)); BUILD_TRY(sb.ensure(1));
    BUILD_TRY(b0.ensure(2 * (size_t)N)); BUILD_TRY(b1.ensure(2 * (size_t)N)); BUILD_TRY(bcount.ensure(2 * (size_t)N));
    BUILD_TRY(items[0].ensure(N)); BUILD_TRY(items[1].ensure(N)); BUILD_TRY(cc.ensure(1));
    if (use_lbvh) {
        BUILD_TRY(ti[0].ensure(N)); BUILD_TRY(ti[1].ensure(N)); BUILD_TRY(ti[2].ensure(2 * (size_t)N)); BUILD_TRY(ti[3].ensure(N)); BUILD_TRY(ti[4].ensure(N)); BUILD_TRY(ti[5].ensure(N));
        BUILD_TRY(tf[0].ensure(6 * (size_t)N)); BUILD_TRY(tf[1].ensure(6 * (size_t)N));
        BUILD_TRY(cid[0].ensure(N));
    } else {
        BUILD_TRY(cid[0].ensure(N)); BUILD_TRY(cid[1].ensure(N)); BUILD_TRY(nn.ensure(N)); BUILD_TRY(bsums.ensure(div_up(N, PLOC_THREADS) + 1));
    }

    // ---- K1: Morton keys of the centroids
    BUILD_TRY(cudaEventRecord(e0, st));
    k_pack_shade<<<div_up(N, 256), 256, 0, st>>>(ctx->d_nrm.p, ctx->d_uv.p, ctx->d_tri_geom.p, N, ctx->d_shade.p);
    k_init_bounds<<<1, 32, 0, st>>>(sb.p);
    k_scene_bounds<<<std::min<unsigned>(div_up(N, 256), ctx->sm_count * 8), 256, 0, st>>>(ctx->d_pos.p, N, sb.p);
    k_morton<<<div_up(N, 256), 256, 0, st>>>(ctx->d_pos.p, N, sb.p, keys[0].p, vals[0].p);
    ctx->launches += 4;
    // ---- K2: radix sort
    BUILD_TRY(cudaEventRecord(e1, st));
    int cur = 0;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = pass * 8;
        k_rs_hist<<<n_tiles, RS_THREADS, 0, st>>>(keys[cur].p, N, shift, hist.p, n_tiles);
        k_rs_scan_rows<<<256, 1024, 0, st>>>(hist.p, n_tiles, hist.p + 256 * (size_t)n_tiles);
        k_rs_scatter<<<n_tiles, RS_THREADS, 0, st>>>(keys[cur].p, vals[cur].p, keys[cur ^ 1].p, vals[cur ^ 1].p, N, shift, hist.p, hist.p + 256 * (size_t)n_tiles, n_tiles);
        cur ^= 1; ctx->launches += 3;
    }
    BUILD_TRY(cudaEventRecord(e2, st));
    // ---- K3/K4: binary tree over the Morton order
    uint32_t root2 = 0, passes = 0;
    if (use_lbvh) {
        if (N >= 2) {
            BinTree t; t.left = ti[0].p; t.right = ti[1].p; t.parent = ti[2].p; t.first = ti[3].p; t.last = ti[4].p; t.flag = ti[5].p; t.lo = tf[0].p; t.hi = tf[1].p;
            k_karras<<<div_up(N - 1, 256), 256, 0, st>>>(keys[cur].p, (int)N, t);
            k_refit<<<div_up(N, 256), 256, 0, st>>>(ctx->d_pos.p, vals[cur].p, (int)N, t);
            k_lbvh_to_b2<<<div_up(2 * (size_t)N - 1, 256), 256, 0, st>>>((int)N, t, vals[cur].p, b0.p, b1.p, bcount.p);
            ctx->launches += 3;
            root2 = N;
        } else {
            k_ploc_init<<<1, 256, 0, st>>>(ctx->d_pos.p, vals[cur].p, N, b0.p, b1.p, bcount.p, cid[0].p);
            ctx->launches += 1;
        }
    } else {
        k_ploc_init<<<div_up(N, 256), 256, 0, st>>>(ctx->d_pos.p, vals[cur].p, N, b0.p, b1.p, bcount.p, cid[0].p);
        ctx->launches += 1;
        uint32_t n = N, next = N; int c = 0;
        while (n > PLOC_FINISH) {
            const uint32_t nb = div_up(n, PLOC_THREADS);
            uint32_t survivors = 0, merges = 0;
            for (int mode = ploc_ties;;) {
                k_ploc_nn<<<nb, PLOC_THREADS, 0, st>>>(b0.p, b1.p, cid[c].p, n, nn.p, mode);
                k_ploc_count<<<nb, PLOC_THREADS, 0, st>>>(nn.p, n, bsums.p);
                k_ploc_scan<<<1, 1024, 0, st>>>(bsums.p, nb, bsums.p + nb);
                ctx->launches += 3;
                BUILD_TRY(cudaMemcpyAsync(ctx->h_pin, bsums.p + nb, sizeof(uint2), cudaMemcpyDeviceToHost, st));
                BUILD_TRY(cudaStreamSynchronize(st));
                survivors = ctx->h_pin[0]; merges = ctx->h_pin[1];
                if (mode == PLOC_TIES_BUDDY || !ploc_pass_stalled(merges, n)) break;
                mode = PLOC_TIES_BUDDY;   // a run of ties pointed every cluster at the same neighbour: redo the pass pairing buddies (bvh8.cuh)
            }
            if (merges == 0 || survivors >= n) { rc = ctx->fail(PGRT_ERR_INVALID, "pgrt_commit: the triangles do not cluster (non-finite vertex coordinates?)"); cleanup(); return rc; }
            k_ploc_apply<<<nb, PLOC_THREADS, 0, st>>>(nn.p, cid[c].p, n, bsums.p, next, b0.p, b1.p, bcount.p, cid[c ^ 1].p);
            ctx->launches += 1; passes++;
            n = survivors; next += merges; c ^= 1;
        }
        if (n > 1) {
            k_ploc_finish<<<1, PLOC_FINISH, 0, st>>>(cid[c].p, n, next, b0.p, b1.p, bcount.p, (uint32_t*)(bsums.p), ploc_ties);
            ctx->launches += 1;
            BUILD_TRY(cudaMemcpyAsync(ctx->h_pin, bsums.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            BUILD_TRY(cudaStreamSynchronize(st));
            if (ctx->h_pin[0] != 1u) { rc = ctx->fail(PGRT_ERR_INVALID, "pgrt_commit: the triangles do not cluster (non-finite vertex coordinates?)"); cleanup(); return rc; }
        }
        root2 = N >= 2 ? 2 * N - 2 : 0;
    }
    BUILD_TRY(cudaEventRecord(e3, st));
    // ---- K5: collapse to the 8-wide quantised layout, one level per launch
    Bvh2View view; view.b0 = b0.p; view.b1 = b1.p; view.count = bcount.p; view.leaf_tri = vals[cur].p; view.n_leaves = N;
    {
        CollapseCounters init = {}; init.nodes = 1;
        const uint2 first = make_uint2(root2, 0u);
        BUILD_TRY(cudaMemcpyAsync(cc.p, &init, sizeof init, cudaMemcpyHostToDevice, st));
        BUILD_TRY(cudaMemcpyAsync(items[0].p, &first, sizeof first, cudaMemcpyHostToDevice, st));
    }
    uint32_t n_items = 1, depth = 0; int ic = 0;
    CollapseCounters hcc = {};
    while (n_items) {
        k_collapse<<<div_up(n_items, 128), 128, 0, st>>>(view, ctx->d_pos.p, items[ic].p, n_items, items[ic ^ 1].p, cc.p, ctx->d_nodes.p, ctx->d_tris.p, layout);
        ctx->launches += 1; depth++;
        BUILD_TRY(cudaMemcpyAsync(ctx->h_pin, cc.p, sizeof(CollapseCounters), cudaMemcpyDeviceToHost, st));
        BUILD_TRY(cudaMemsetAsync(&cc.p->next_items, 0, sizeof(uint32_t), st));
        BUILD_TRY(cudaStreamSynchronize(st));
        memcpy(&hcc, ctx->h_pin, sizeof hcc);
        n_items = hcc.next_items; ic ^= 1;
        if (depth > 4096) break;
    }
    BUILD_TRY(cudaEventRecord(e4, st));
    BUILD_TRY(cudaGetLastError());
    float4 rb[2];
    BUILD_TRY(cudaMemcpyAsync(&rb[0], b0.p + root2, sizeof(float4), cudaMemcpyDeviceToHost, st));
    BUILD_TRY(cudaMemcpyAsync(&rb[1], b1.p + root2, sizeof(float4), cudaMemcpyDeviceToHost, st));
    BUILD_TRY(cudaStreamSynchronize(st));
    if (hcc.tris != N) { rc = ctx->fail(PGRT_ERR_CUDA, "pgrt_commit: collapse emitted a wrong number of triangles"); cleanup(); return rc; }
    if (depth + 2 > PGRT_STACK8) { rc = ctx->fail(PGRT_ERR_INVALID, "pgrt_commit: the tree is deeper than the traversal stack (degenerate input)"); cleanup(); return rc; }
    ctx->bb_lo[0] = rb[0].x; ctx->bb_lo[1] = rb[0].y; ctx->bb_lo[2] = rb[0].z; ctx->bb_hi[0] = rb[1].x; ctx->bb_hi[1] = rb[1].y; ctx->bb_hi[2] = rb[1].z;
    if (getenv("PGRT_NO_SCENE_BOX")) for (int a = 0; a < 3; ++a) { ctx->bb_lo[a] = -FLT_MAX; ctx->bb_hi[a] = FLT_MAX; }   // A/B: the test never culls
    const float ra = box_half_area(rb[0], rb[1]);
    bs.nodes = hcc.nodes; bs.sah_cost = ra > 0.0f ? hcc.sah / ra : (float)N;
    bs.depth = depth; bs.ploc_passes = passes; bs.node_bytes = (uint32_t)(node_f4 * 16);
    cudaEventElapsedTime(&bs.build_ms, e0, e4);
    cudaEventElapsedTime(&bs.sort_ms, e1, e2);
    cudaEventElapsedTime(&bs.tree_ms, e2, e3);
    cudaEventElapsedTime(&bs.collapse_ms, e3, e4);
#undef BUILD_TRY
    cleanup();
    ctx->committed = true; ctx->last_build = bs;
    if (stats) *stats = bs;
    // PGRT_L2_PERSIST=1 (measurement, off by default): an access-policy window that marks the node array as persisting in L2
    // on every slot stream (the frame kernels are captured with it), so that streaming traffic -- the frame, the bench's L2
    // flush -- does not evict the tree.  profiles/r2_l2_persist.txt says what it buys.
    if (const char* e = getenv("PGRT_L2_PERSIST")) {
        cudaDeviceProp prop;
        if (atoi(e) != 0 && cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
            const size_t bytes = (size_t)bs.nodes * node_f4 * 16;
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, bytes + (bytes >> 2)));
            cudaStreamAttrValue v = {};
            v.accessPolicyWindow.base_ptr = ctx->d_nodes.p;
            v.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)prop.accessPolicyMaxWindowSize);
            v.accessPolicyWindow.hitRatio = 1.0f;
            v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            for (FrameSlot& S : ctx->slots) { cudaStreamSetAttribute(S.stream, cudaStreamAttributeAccessPolicyWindow, &v); S.gkey = 0; }
            cudaGetLastError();
        }
    }
    return PGRT_OK;
}

// --------------------------------------------------------------------------------------------------- camera / params
extern "C" int pgrt_set_camera(pgrt_context* ctx, int32_t width, int32_t height, float fov_y, const float from[3], const float at[3]) {
    CHECK_CTX(ctx);
    if (width <= 0 || height <= 0 || !from || !at) return ctx->fail(PGRT_ERR_INVALID, "pgrt_set_camera: bad arguments");
    // PinHoleCamera::PinHoleCamera (PinHoleCamera.cpp:5-29); this file is compiled without FMA contraction on the host too
    DevCamera c;
    c.width = width; c.height = height;
    c.f_y = height / (2 * tanf(fov_y / 2));
    c.from = v3(from[0], from[1], from[2]);
    const V3 up = v3(0.0f, 0.0f, 1.0f);
    V3 z_c = c.from - v3(at[0], at[1], at[2]);
    V3 x_c = cross3(up, z_c);
    V3 y_c = cross3(z_c, x_c);
    z_c = normalize3(z_c); x_c = normalize3(x_c); y_c = normalize3(y_c);
    c.M.m00 = x_c.x; c.M.m01 = y_c.x; c.M.m02 = z_c.x;
    c.M.m10 = x_c.y; c.M.m11 = y_c.y; c.M.m12 = z_c.y;
    c.M.m20 = x_c.z; c.M.m21 = y_c.z; c.M.m22 = z_c.z;
    ctx->cam = c; ctx->cam_set = true; ctx->px_valid = false;
    ctx->shard.tiles_x = (width + PGRT_TILE_W - 1) / PGRT_TILE_W;
    ctx->shard.tiles_y = (height + PGRT_TILE_H - 1) / PGRT_TILE_H;
    return PGRT_OK;
}

extern "C" void pgrt_default_params(pgrt_render_params* p) {
    if (!p) return;
    memset(p, 0, sizeof *p);
    p->sampling_width = 3; p->jitter = 1; p->focal_distance = 200.0f; p->aperture = 5.0f; p->max_depth = 7; p->gamma_level = 0.5f;
    p->seed = 1; p->camera_mode = 0; p->shader_mode = 0; p->scheduler = 3;
}

extern "C" int pgrt_set_shard(pgrt_context* ctx, int32_t rank, int32_t n_ranks) {
    CHECK_CTX(ctx);
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return ctx->fail(PGRT_ERR_INVALID, "pgrt_set_shard: bad rank");
    ctx->shard.rank = rank; ctx->shard.n_ranks = n_ranks; ctx->px_valid = false;
    return PGRT_OK;
}

static uint64_t shard_slots(const pgrt_context* ctx) {
    const uint64_t tiles = (uint64_t)ctx->shard.tiles_x * ctx->shard.tiles_y;
    return (tiles + ctx->shard.n_ranks - 1) / ctx->shard.n_ranks * PGRT_TILE_PIXELS;
}
extern "C" uint64_t pgrt_shard_pixels(const pgrt_context* ctx) { return ctx ? shard_slots(ctx) : 0; }

// --------------------------------------------------------------------------------------------------- frame
static int ensure_pool(pgrt_context* ctx, FrameSlot& S, size_t cap) {
    for (auto& b : S.pool_f4) CUDA_TRY(b.ensure(cap));
    for (auto& b : S.pool_u2) CUDA_TRY(b.ensure(cap));
    for (auto& b : S.pool_u32) CUDA_TRY(b.ensure(cap));
    if (cap > S.pool_u4[0].n) {
        // the link array carries the publication words: fresh memory must not look like a record of a recent batch
        CUDA_TRY(S.pool_u4[0].ensure(cap));
        CUDA_TRY(cudaMemsetAsync(S.pool_u4[0].p, 0, S.pool_u4[0].n * sizeof(uint4), S.stream));
    }
    RayPool& P = S.pool;
    P.ray_o = S.pool_f4[0].p; P.ray_d = S.pool_f4[1].p; P.color = S.pool_f4[2].p; P.att = S.pool_f4[3].p;
    P.child = S.pool_u2[0].p; P.link = S.pool_u4[0].p; P.pending = S.pool_u32[0].p;
    P.cap = (uint32_t)std::min<size_t>(cap, 0xFFFFFFF0u);
    return PGRT_OK;
}

static int ensure_levels(pgrt_context* ctx, FrameSlot& S, int n_levels, size_t cap0, size_t capn, bool fused) {
    for (int l = 0; l <= n_levels; ++l) {   // one spare level so k_shade always has a (never written) "next"
        const size_t cap = l == 0 ? cap0 : (l == n_levels ? 1 : capn);
        // the fused scheduler keeps rays and hits in registers: level 0 holds colours and dielectric nodes only
        for (int k = 0; k < 5; ++k) if (!fused || k >= 3) CUDA_TRY(S.lv_f4[l][k].ensure(cap));
        CUDA_TRY(S.lv_child[l].ensure(cap));
        if (!fused) { CUDA_TRY(S.lv_list[l][0].ensure(cap)); CUDA_TRY(S.lv_list[l][1].ensure(cap)); }
        LevelBufs& L = S.levels[l];
        L.ray_o = S.lv_f4[l][0].p; L.ray_d = S.lv_f4[l][1].p; L.hit = S.lv_f4[l][2].p; L.color = S.lv_f4[l][3].p; L.dn_att = S.lv_f4[l][4].p;
        L.dn_child = S.lv_child[l].p; L.phong_list = S.lv_list[l][0].p; L.diel_list = S.lv_list[l][1].p;
        L.pending = nullptr;
        L.cap = (uint32_t)cap;
    }
    CUDA_TRY(S.l0_pending.ensure(cap0));
    S.levels[0].pending = S.l0_pending.p;
    return PGRT_OK;
}

struct FrameTimer {
    FrameSlot* S; bool on;
    void begin(int cls, int level = 0) {
        if (!on) return;
        if (S->ev_used + 2 > S->ev_pool.size()) {
            if (S->ev_pool.size() >= 16384) { on = false; return; }
            for (int k = 0; k < 256; ++k) { cudaEvent_t e; cudaEventCreate(&e); S->ev_pool.push_back(e); }
            S->ev_class.resize(S->ev_pool.size() / 2); S->ev_level.resize(S->ev_pool.size() / 2);
        }
        S->ev_class[S->ev_used / 2] = cls; S->ev_level[S->ev_used / 2] = level;
        cudaEventRecord(S->ev_pool[S->ev_used], S->stream);
    }
    void end() { if (!on) return; cudaEventRecord(S->ev_pool[S->ev_used + 1], S->stream); S->ev_used += 2; }
};

static int validate_frame(pgrt_context* ctx, const pgrt_render_params* p) {
    if (!p) return ctx->fail(PGRT_ERR_INVALID, "render: null params");
    if (!ctx->committed) return ctx->fail(PGRT_ERR_INVALID, "render: pgrt_commit has not been called since the scene changed");
    if (!ctx->cam_set) return ctx->fail(PGRT_ERR_INVALID, "render: pgrt_set_camera has not been called");
    if (p->sampling_width < 1 || p->sampling_width > 64) return ctx->fail(PGRT_ERR_INVALID, "render: sampling_width out of range [1,64]");
    if (p->max_depth < 0 || p->max_depth >= PGRT_MAX_LEVELS) return ctx->fail(PGRT_ERR_INVALID, "render: max_depth out of range [0,32]");
    if (p->scheduler < 0 || p->scheduler > 3) return ctx->fail(PGRT_ERR_INVALID, "render: scheduler must be 0 (fused), 1 (level-synchronous), 2 (hybrid) or 3 (automatic)");
    if (p->shadow_mode != 0 && p->shadow_mode != 1) return ctx->fail(PGRT_ERR_INVALID, "render: shadow_mode must be 0 (as shipped) or 1 (hit point -> light)");
    return upload_tables(ctx);
}

// dest_mode: 0 = full frame (W*H pixels), 1 = compact shard buffer, 2 = ids only (geom/prim in d_ids);  fmt: 0 float4, 1 RGBA8
static size_t pixel_bytes(const FrameSlot& S) { return S.fmt == 1 ? 4 : 16; }

static int prepare_frame(pgrt_context* ctx, FrameSlot& S) {
    const int SPP = S.params.sampling_width * S.params.sampling_width;
    const size_t cap0 = (size_t)S.batch_slots * SPP;
    // deeper levels: Whitted spawns rays at dielectric hits only (a fraction of the samples); the path-tracing mode spawns
    // one at every hit of every level, so its queues are sized for a full level each.  pool_scale grows after an overflow.
    const bool path = S.params.shader_mode == 3;
    const size_t capn = std::max<size_t>(ctx->min_level_cap, (size_t)((path ? std::max(1.0, ctx->level_cap_factor) : ctx->level_cap_factor) * ctx->pool_scale * (double)cap0));
    int rc = ensure_levels(ctx, S, (S.fused || S.hybrid) ? 1 : S.n_levels, cap0, capn, S.fused);
    if (rc) return rc;
    if (S.fused || S.hybrid) { rc = ensure_pool(ctx, S, capn * (size_t)std::min(S.n_levels - 1, path ? 8 : 4)); if (rc) return rc; }   // one pool replaces the per-level queues
    if (S.host_dst) CUDA_TRY(S.d_frame.ensure(((size_t)ctx->cam.width * ctx->cam.height * pixel_bytes(S) + 15) / 16));
    return PGRT_OK;
}

// the launches of one frame, in order, on the slot's stream (directly, or into a stream capture)
static int record_frame(pgrt_context* ctx, FrameSlot& S) {
    cudaStream_t st = S.stream;
    const pgrt_render_params* p = &S.params;
    const int dest_mode = S.dest_mode, profile = S.profile;
    const int SPP = p->sampling_width * p->sampling_width;
    const int n_levels = S.n_levels;
    const uint64_t total_slots = shard_slots(ctx);
    const uint64_t batch_slots = S.batch_slots;
    const DevScene sc = ctx->dev_scene();
    const unsigned trace_grid = ctx->sm_count * ctx->trace_ctas_per_sm, phong_grid = ctx->sm_count * 16, shade_grid = ctx->sm_count * 8;
    const bool fused = S.fused, hybrid = S.hybrid;
    FrameTimer tm{&S, (profile & 1) != 0};
    const bool count = (profile & 2) != 0;
    const bool path = p->shader_mode == 3;   // path-tracing instantiations of k_shade / k_frame / k_combine
    pgrt_render_stats& rs = S.rs;
    S.ev_used = 0; rs.launches = 0; rs.trace_launches = 0; rs.batches = 0;
    Counters* cnt = S.d_counters.p;
    FrameOut fo = {};
    void* out = S.host_dst ? (void*)S.d_frame.p : S.dest;
    if (S.fmt == 1) fo.rgba8 = (uint32_t*)out; else fo.rgba = (float4*)out;
    fo.compact = dest_mode == 1; fo.direct = ((fused || (hybrid && !path)) && SPP == 1) ? 1 : 0;   // (path tracing keeps a node's own value in its colour slot)
    FrameOut fo_q = fo; fo_q.direct = 0;                                                            // level-synchronous kernels: samples go through the colour queue
    for (uint64_t slot0 = 0; slot0 < total_slots; slot0 += batch_slots) {
        const uint32_t n_slots = (uint32_t)std::min<uint64_t>(batch_slots, total_slots - slot0);
        const uint32_t n0 = n_slots * (uint32_t)SPP;
        const bool first = slot0 == 0, last = slot0 + batch_slots >= total_slots;
        rs.batches++;
        k_batch_begin<<<1, 64, 0, st>>>(cnt, n0, first ? 1 : 0, hybrid ? 1 : 0); rs.launches++;
        Gen0 g0; g0.cam = ctx->cam; g0.sh = ctx->shard; g0.slot0 = (uint32_t)slot0; g0.n_slots = n_slots; g0.spp = SPP; g0.on = (fused || ctx->fuse_raygen) ? 1 : 0;
        Gen0 gN = g0; gN.on = 0;
        if (fused) {
            tm.begin(KC_TRACE, 0);
            if (path) {
                if (count) k_frame<true, true><<<S.frame_grid, PGRT_FRAME_THREADS, 0, st>>>(sc, *p, g0, S.levels[0], S.pool, fo, ctx->min_claim, ctx->claim_patience, S.keep_ctas, ctx->pool_policy, cnt);
                else k_frame<false, true><<<S.frame_grid, PGRT_FRAME_THREADS, 0, st>>>(sc, *p, g0, S.levels[0], S.pool, fo, ctx->min_claim, ctx->claim_patience, S.keep_ctas, ctx->pool_policy, cnt);
            } else {
                if (count) k_frame<true, false><<<S.frame_grid, PGRT_FRAME_THREADS, 0, st>>>(sc, *p, g0, S.levels[0], S.pool, fo, ctx->min_claim, ctx->claim_patience, S.keep_ctas, ctx->pool_policy, cnt);
                else k_frame<false, false><<<S.frame_grid, PGRT_FRAME_THREADS, 0, st>>>(sc, *p, g0, S.levels[0], S.pool, fo, ctx->min_claim, ctx->claim_patience, S.keep_ctas, ctx->pool_policy, cnt);
            }
            rs.launches++; rs.trace_launches++;
            tm.end();
        } else {
            if (!g0.on) {
                tm.begin(KC_SHADE);
                k_raygen<<<div_up(n0, 256), 256, 0, st>>>(ctx->cam, *p, ctx->shard, (uint32_t)slot0, n_slots, S.levels[0], cnt); rs.launches++;
                tm.end();
            }
            for (int l = 0; l < (hybrid ? 1 : n_levels); ++l) {
                tm.begin(KC_TRACE, l);
                // level 0 of a small batch (a shard of a frame): a grid sized to the rays -- one CTA per PGRT_TRACE_RAYS_PER_CTA (1024:
                // eight 32-ray chunks per warp) -- so that warps refill instead of hundreds of CTAs starting for one chunk each (an
                // eighth of C2 through the hybrid: 0.1117 ms with 888 CTAs, 0.1028 with 296: profiles/r2_sweep_shard_depth.txt)
                const unsigned tg = l == 0 ? std::max(std::min(trace_grid, (unsigned)ctx->sm_count), std::min(trace_grid, div_up(n0, ctx->trace_rays_per_cta))) : trace_grid;
                if (count) k_trace<true><<<tg, 128, 0, st>>>(sc, *p, l == 0 ? g0 : gN, S.levels[l], l, ctx->trace_refill, cnt);
                else k_trace<false><<<tg, 128, 0, st>>>(sc, *p, l == 0 ? g0 : gN, S.levels[l], l, ctx->trace_refill, cnt);
                rs.launches++; rs.trace_launches++;
                tm.end();
                if (dest_mode == 2) break;
                tm.begin(KC_SHADE, l);
                if (hybrid) {
                    if (path) k_shade<true, true><<<shade_grid, 256, 0, st>>>(sc, *p, g0, l, S.levels[l], S.levels[l + 1], S.pool, fo, cnt);
                    else k_shade<false, true><<<shade_grid, 256, 0, st>>>(sc, *p, g0, l, S.levels[l], S.levels[l + 1], S.pool, fo, cnt);
                }
                else if (path) k_shade<true, false><<<shade_grid, 256, 0, st>>>(sc, *p, l == 0 ? g0 : gN, l, S.levels[l], S.levels[l + 1], S.pool, fo_q, cnt);
                else k_shade<false, false><<<shade_grid, 256, 0, st>>>(sc, *p, l == 0 ? g0 : gN, l, S.levels[l], S.levels[l + 1], S.pool, fo_q, cnt);
                rs.launches++;
                tm.end();
                tm.begin(KC_TRACE, l);   // Phong = shading preamble + one inline shadow traversal per light
                if (count) k_phong<true><<<phong_grid, 128, 0, st>>>(sc, *p, l == 0 ? g0 : gN, l, S.levels[l], hybrid ? fo : fo_q, cnt);
                else k_phong<false><<<phong_grid, 128, 0, st>>>(sc, *p, l == 0 ? g0 : gN, l, S.levels[l], hybrid ? fo : fo_q, cnt);
                rs.launches++; rs.trace_launches++;
                tm.end();
            }
            if (hybrid && p->max_depth >= 1) {
                // every ray of level >= 1: the pool half of k_frame (its claim of primary chunks finds trace_next[0] used up by k_trace above)
                tm.begin(KC_TRACE, 1);
                if (path) {
                    if (count) k_frame<true, true><<<S.frame_grid, PGRT_FRAME_THREADS, 0, st>>>(sc, *p, g0, S.levels[0], S.pool, fo, ctx->min_claim, ctx->claim_patience, S.keep_ctas, ctx->pool_policy, cnt);
                    else k_frame<false, true><<<S.frame_grid, PGRT_FRAME_THREADS, 0, st>>>(sc, *p, g0, S.levels[0], S.pool, fo, ctx->min_claim, ctx->claim_patience, S.keep_ctas, ctx->pool_policy, cnt);
                } else {
                    if (count) k_frame<true, false><<<S.frame_grid, PGRT_FRAME_THREADS, 0, st>>>(sc, *p, g0, S.levels[0], S.pool, fo, ctx->min_claim, ctx->claim_patience, S.keep_ctas, ctx->pool_policy, cnt);
                    else k_frame<false, false><<<S.frame_grid, PGRT_FRAME_THREADS, 0, st>>>(sc, *p, g0, S.levels[0], S.pool, fo, ctx->min_claim, ctx->claim_patience, S.keep_ctas, ctx->pool_policy, cnt);
                }
                rs.launches++; rs.trace_launches++;
                tm.end();
            }
        }
        if (dest_mode == 2) {
            uint32_t* geom = ctx->d_ids.p; uint32_t* prim = geom + (size_t)ctx->cam.width * ctx->cam.height;
            k_primary_ids<<<div_up(n_slots, 256), 256, 0, st>>>(sc, ctx->cam, ctx->shard, (uint32_t)slot0, n_slots, SPP, S.levels[0].hit, geom, prim); rs.launches++;
        } else if (!fo.direct) {
            tm.begin(KC_SHADE);
            if (!fused && !hybrid) for (int l = n_levels - 2; l >= 0; --l) {
                if (path) k_combine<true><<<shade_grid, 256, 0, st>>>(l, S.levels[l], S.levels[l + 1], cnt);
                else k_combine<false><<<shade_grid, 256, 0, st>>>(l, S.levels[l], S.levels[l + 1], cnt);
                rs.launches++;
            }
            k_resolve<<<div_up(n_slots, 256), 256, 0, st>>>(ctx->cam, *p, ctx->shard, (uint32_t)slot0, n_slots, S.levels[0].color, fo); rs.launches++;
            tm.end();
        }
        unsigned long long valid_px;
        {   // primary rays = valid pixels of this batch * S (host-side count; partial tiles hold unused slots)
            uint64_t v = 0;
            const uint32_t W = ctx->cam.width, H = ctx->cam.height;
            for (uint64_t k = slot0 / PGRT_TILE_PIXELS; k < (slot0 + n_slots) / PGRT_TILE_PIXELS; ++k) {
                const uint64_t t = k * ctx->shard.n_ranks + ctx->shard.rank;
                if (t >= (uint64_t)ctx->shard.tiles_x * ctx->shard.tiles_y) continue;
                const uint32_t tx = (uint32_t)(t % ctx->shard.tiles_x), ty = (uint32_t)(t / ctx->shard.tiles_x);
                const uint32_t w = std::min<uint32_t>(PGRT_TILE_W, W - tx * PGRT_TILE_W), h = std::min<uint32_t>(PGRT_TILE_H, H - ty * PGRT_TILE_H);
                v += (uint64_t)w * h;
            }
            valid_px = v * SPP;
        }
        k_batch_end<<<1, 64, 0, st>>>(cnt, valid_px, (fused || hybrid) ? 1 : 0, last ? 1 : 0, S.sig_flag, S.sig_add ? 1 : 0); rs.launches++;
    }
    CUDA_TRY(cudaMemcpyAsync(S.h_counters, cnt, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    if (S.host_dst)   // the memcpy of simpleguidx11.cpp:121-124; overlaps the next frame when the destination is pinned
        CUDA_TRY(cudaMemcpyAsync(S.host_dst, S.d_frame.p, (size_t)ctx->cam.width * ctx->cam.height * pixel_bytes(S), cudaMemcpyDeviceToHost, st));
    LAUNCH_OK();
    return PGRT_OK;
}

static uint64_t fnv1a(const void* data, size_t n, uint64_t h = 1469598103934665603ull) {
    const uint8_t* p = (const uint8_t*)data;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

// Everything the recorded launches depend on: a frame whose key equals the slot's cached one is re-submitted as ONE
// cudaGraphLaunch instead of its launches + copies (55 -> ~10 us of host time per frame; it matters once a GPU's share of
// the frame is tens of microseconds, i.e. with many GPUs on a small frame).
static uint64_t frame_key(pgrt_context* ctx, const FrameSlot& S) {
    const DevScene sc = ctx->dev_scene();
    uint64_t h = fnv1a(&S.params, sizeof S.params);
    h = fnv1a(&ctx->cam, sizeof ctx->cam, h); h = fnv1a(&ctx->shard, sizeof ctx->shard, h); h = fnv1a(&sc, sizeof sc, h);
    h = fnv1a(&S.dest, sizeof S.dest, h); h = fnv1a(&S.host_dst, sizeof S.host_dst, h); h = fnv1a(&S.dest_mode, sizeof S.dest_mode, h);
    h = fnv1a(&S.batch_slots, sizeof S.batch_slots, h); h = fnv1a(&S.n_levels, sizeof S.n_levels, h);
    h = fnv1a(S.levels, sizeof(LevelBufs) * (size_t)(S.n_levels + 1), h); h = fnv1a(&S.pool, sizeof S.pool, h);
    const void* extra[5] = {S.d_frame.p, S.d_counters.p, S.stream, S.sig_flag, S.sig_add ? (const void*)1 : nullptr};
    h = fnv1a(extra, sizeof extra, h);
    const int flags[10] = {(S.fused ? 1 : 0) | (S.hybrid ? 2 : 0), (ctx->fuse_raygen ? 1 : 0) | (int)(ctx->trace_rays_per_cta << 1), S.frame_grid, ctx->trace_refill, ctx->trace_ctas_per_sm, S.fmt, ctx->min_claim, S.keep_ctas, ctx->claim_patience, ctx->pool_policy};
    return fnv1a(flags, sizeof flags, h) | 1ull;
}

static bool host_ptr_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

typedef int (*pgrt_cu_memop32)(cudaStream_t, unsigned long long, uint32_t, unsigned int);   // CUresult cuStream{Wait,Write}Value32(CUstream, CUdeviceptr, cuuint32_t, unsigned)

// stores `value` in a device-visible word, in stream order
static int stream_write32(pgrt_context* ctx, cudaStream_t st, void* addr, uint32_t value) {
    if (ctx->fn_write32 && ((pgrt_cu_memop32)ctx->fn_write32)(st, (unsigned long long)(uintptr_t)addr, value, 0u) == 0) return PGRT_OK;
    CUDA_TRY(cudaMemcpyAsync(addr, &value, sizeof value, cudaMemcpyHostToDevice, st));   // pageable source: staged before the call returns
    return PGRT_OK;
}

// Enqueues one frame on the slot's stream: every launch, the read-back of the counters and (host destinations) of the
// frame itself.  Nothing here waits for the GPU once the slot's buffers exist.
static int enqueue_frame(pgrt_context* ctx, FrameSlot& S) {
    int rc = prepare_frame(ctx, S);
    if (rc) return rc;
    cudaStream_t st = S.stream;
    if (S.sig_flag && !S.sig_add) { rc = stream_write32(ctx, st, &S.d_counters.p->sig_value, S.sig_value); if (rc) return rc; }   // outside the graph: it changes every frame
    CUDA_TRY(cudaEventRecord(S.ev_frame0, st));
    bool submitted = false;
    const bool graphable = ctx->use_graphs && S.profile == 0 && S.dest_mode != 2 && (!S.host_dst || host_ptr_is_pinned(S.host_dst));
    const uint64_t key = graphable ? frame_key(ctx, S) : 0;
    if (graphable && S.gexec && S.gkey == key) {
        S.rs.launches = S.g_launches; S.rs.trace_launches = S.g_trace_launches; S.rs.batches = S.g_batches;
        CUDA_TRY(cudaGraphLaunch(S.gexec, st));
        submitted = true;
    } else if (graphable && S.gseen == key) {
        // second frame with this key (the first one ran directly: lazily loaded kernels must not meet a capture): record it
        if (S.gexec) { cudaGraphExecDestroy(S.gexec); S.gexec = nullptr; }
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            rc = record_frame(ctx, S);
            const cudaError_t e = cudaStreamEndCapture(st, &graph);
            if (rc == PGRT_OK && e == cudaSuccess && graph && cudaGraphInstantiate(&S.gexec, graph, 0) == cudaSuccess) {
                S.gkey = key; S.g_launches = S.rs.launches; S.g_trace_launches = S.rs.trace_launches; S.g_batches = S.rs.batches;
                if (cudaGraphLaunch(S.gexec, st) == cudaSuccess) submitted = true;
            }
            if (graph) cudaGraphDestroy(graph);
        }
        if (!submitted) {   // capture is an optimisation: fall back to direct launches for good
            cudaGetLastError();
            if (S.gexec) { cudaGraphExecDestroy(S.gexec); S.gexec = nullptr; }
            ctx->use_graphs = false;
        }
    }
    if (!submitted) {
        rc = record_frame(ctx, S);
        if (rc) return rc;
    }
    S.gseen = key;
    CUDA_TRY(cudaEventRecord(S.ev_frame1, st));
    CUDA_TRY(cudaEventRecord(S.ev_done, st));
    ctx->launches += S.rs.launches;
    return PGRT_OK;
}

static int frame_begin(pgrt_context* ctx, int slot, const pgrt_render_params* p, void* dest, int dest_mode, void* host_dst, int profile, int fmt = 0) {
    if (slot < 0 || slot >= PGRT_MAX_INFLIGHT) return ctx->fail(PGRT_ERR_INVALID, "render: slot out of range");
    FrameSlot& S = ctx->slots[slot];
    if (S.busy) return ctx->fail(PGRT_ERR_INVALID, "render: the slot still holds a frame in flight (call pgrt_render_end first)");
    int rc = validate_frame(ctx, p);
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    S.params = *p; S.dest = dest; S.dest_mode = dest_mode; S.host_dst = host_dst; S.profile = profile; S.fmt = fmt;
    const int SPP = p->sampling_width * p->sampling_width;
    S.n_levels = (dest_mode == 2) ? 1 : p->max_depth + 1;
    {
        // scheduler 3 = automatic: the hybrid pays three more launches and the level-0 queues' round trip through L2 for lean,
        // refilled level-0 kernels -- faster from about 400 k samples per batch (C2: whole frame 0.40 against 0.54 ms, a quarter
        // of it 0.161 against 0.175, an eighth 0.117 against 0.106: profiles/r2_sched_by_shard.txt)
        int sched = p->scheduler;
        if (sched == 3) {
            const uint64_t bs = std::min<uint64_t>(shard_slots(ctx), std::max<uint64_t>(PGRT_TILE_PIXELS, ctx->max_batch_samples / SPP / PGRT_TILE_PIXELS * PGRT_TILE_PIXELS));
            sched = bs * (uint64_t)SPP >= ctx->auto_hybrid_min ? 2 : 0;
        }
        S.fused = sched == 0 && dest_mode != 2;
        S.hybrid = sched == 2 && dest_mode != 2;
    }
    if ((S.fused || S.hybrid) && ctx->frame_per_sm_max == 0) {
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->frame_per_sm_max, k_frame<false, false>, PGRT_FRAME_THREADS, 0));
        ctx->frame_per_sm_max = std::max(1, ctx->frame_per_sm_max);
        if (const char* e = getenv("PGRT_FRAME_CTAS_PER_SM")) ctx->frame_per_sm_env = std::max(1, atoi(e));
    }
    const uint64_t total_slots = shard_slots(ctx);
    uint64_t batch_slots = std::max<uint64_t>(PGRT_TILE_PIXELS, ctx->max_batch_samples / SPP / PGRT_TILE_PIXELS * PGRT_TILE_PIXELS);
    S.batch_slots = std::min(batch_slots, total_slots);
    if (ctx->batch_limit) S.batch_slots = std::min(S.batch_slots, ctx->batch_limit);   // do not overflow the same way every frame
    if (S.fused || S.hybrid) {
        // A persistent grid: as many CTAs as fit (register-bound) for a large batch; a small batch (a shard of a frame, a
        // 640x480 frame) takes one CTA per four 32-ray chunks per warp, so that the frames in flight behind it find room
        const uint64_t samples = S.batch_slots * (uint64_t)SPP;
        const int per_sm = ctx->frame_per_sm_env ? std::min(ctx->frame_per_sm_env, ctx->frame_per_sm_max) : ctx->frame_per_sm_max;
        int grid = (int)std::min<uint64_t>((uint64_t)ctx->sm_count * per_sm, std::max<uint64_t>(1, (samples + 4 * PGRT_FRAME_THREADS - 1) / (4 * PGRT_FRAME_THREADS)));
        if (const char* e = getenv("PGRT_FRAME_CTAS")) grid = std::max(1, atoi(e));
        S.frame_grid = grid;
        // A few CTAs stay for the dependent chains of the secondary rays; the rest leave when they run dry, so the next frame's
        // kernel finds room.  The end of a frame is a latency chain (max_depth traversals in a row), not a throughput problem: 4 or
        // 148 keepers give the same single-frame time (1.1 ms on C2), but every keeper holds an SM slot the following frames
        // cannot use: 8 keepers 0.486 ms per pipelined frame, 148 keepers 0.587 (profiles/r2_sweep_kframe_keepers.txt).
        S.keep_ctas = std::min(grid, ctx->keep_ctas * (128 / PGRT_FRAME_THREADS));      // (PGRT_KEEP_CTAS counts CTAs of 128 threads)
    }
    S.rs = pgrt_render_stats{};
    rc = enqueue_frame(ctx, S);
    if (rc) return rc;
    S.busy = true; S.seq_expected++;
    return PGRT_OK;
}

static int frame_end(pgrt_context* ctx, int slot, pgrt_render_stats* stats) {
    if (slot < 0 || slot >= PGRT_MAX_INFLIGHT) return ctx->fail(PGRT_ERR_INVALID, "render: slot out of range");
    FrameSlot& S = ctx->slots[slot];
    if (!S.busy) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render_end: no frame in flight in this slot");
    cudaSetDevice(ctx->device);
    S.busy = false;
    for (;;) {
        CUDA_TRY(cudaStreamSynchronize(S.stream));
        if (S.h_counters->watchdog) {
            S.seq_expected = S.h_counters->done_seq;
            const uint32_t* w = S.h_counters->wd_info;
            char buf[384];
            snprintf(buf, sizeof buf, "render: internal error in the frame kernel (a bounded wait expired; the frame is incomplete): ticket base %u mask %08x, q_tail %u q_head %u "
                     "outstanding %u, block %u more_primary %u, primary chunks claimed up to %u of %llu, pool cap %u, grid %d keepers %d", w[0], w[1], w[2], w[3], w[4], w[5] & 0x7fffffffu,
                     w[5] >> 31, w[6], (unsigned long long)S.batch_slots * S.params.sampling_width * S.params.sampling_width, w[7], S.frame_grid, S.keep_ctas);
            return ctx->fail(PGRT_ERR_CUDA, buf);
        }
        if (!S.h_counters->overflow) break;
        // A secondary-ray queue overflowed (rare; sizes are generous): the attempt did not signal completion (k_batch_end), so
        // no consumer ordered behind this slot has been released.  Render the frame again with four times the capacity -
        // remembered until the next commit - or, once that would take more than a few GB per slot, in smaller batches.
        const uint32_t retries = S.rs.overflow_retries + 1;
        const size_t pool_bytes = (S.fused || S.hybrid) ? S.pool_f4[0].n * 100 : (size_t)S.n_levels * S.lv_f4[1][0].n * 100;
        if (pool_bytes <= ((size_t)2 << 30)) ctx->pool_scale *= 4.0;
        else {
            if (S.batch_slots <= PGRT_TILE_PIXELS) { S.seq_expected = S.h_counters->done_seq; return ctx->fail(PGRT_ERR_OVERFLOW, "render: secondary-ray queues overflow at the minimum batch; raise PGRT_MIN_LEVEL_CAP"); }
            S.batch_slots = std::max<uint64_t>(PGRT_TILE_PIXELS, S.batch_slots / 2 / PGRT_TILE_PIXELS * PGRT_TILE_PIXELS);
            ctx->batch_limit = S.batch_slots;
        }
        S.rs = pgrt_render_stats{}; S.rs.overflow_retries = retries;
        int rc = enqueue_frame(ctx, S);
        if (rc) { S.seq_expected = S.h_counters->done_seq; return rc; }
    }
    pgrt_render_stats& rs = S.rs;
    const Counters& hc = *S.h_counters;
    rs.rays_primary = hc.tot_primary; rs.rays_shadow = hc.tot_shadow; rs.rays_reflection = hc.tot_reflection; rs.rays_refraction = hc.tot_refraction;
    cudaEventElapsedTime(&rs.frame_ms, S.ev_frame0, S.ev_frame1);
    ctx->level_stats_n = S.n_levels;
    for (int l = 0; l <= PGRT_MAX_LEVELS; ++l) {
        pgrt_level_stats& ls = ctx->level_stats[l];
        ls = pgrt_level_stats{};
        ls.rays = hc.lv_rays[l]; ls.shadow_rays = hc.lv_shadow[l]; ls.nodes = hc.lv_nodes[l]; ls.tris = hc.lv_tris[l];
        ls.shadow_nodes = hc.lv_sh_nodes[l]; ls.shadow_tris = hc.lv_sh_tris[l]; ls.max_nodes = hc.lv_max_nodes[l]; ls.shadow_max_nodes = hc.lv_sh_max_nodes[l];
        if (S.fused && l >= 1 && hc.lv_t_last[l] > hc.t_first && hc.lv_t_first[l] != ~0ull) {
            // fused scheduler: when the level's first ray was taken up / its last ray finished, in ms after the kernel's first warp
            ls.shade_ms = (float)((double)(hc.lv_t_first[l] - hc.t_first) * 1e-6); ls.trace_ms = (float)((double)(hc.lv_t_last[l] - hc.t_first) * 1e-6);
        }
        rs.nodes_visited += ls.nodes + ls.shadow_nodes; rs.tris_tested += ls.tris + ls.shadow_tris;
        rs.max_nodes_per_ray = std::max(rs.max_nodes_per_ray, std::max(ls.max_nodes, ls.shadow_max_nodes));
    }
    for (size_t k = 0; k < S.ev_used; k += 2) {
        float ms = 0.f; cudaEventElapsedTime(&ms, S.ev_pool[k], S.ev_pool[k + 1]);
        pgrt_level_stats& ls = ctx->level_stats[S.ev_level[k / 2]];
        if (S.ev_class[k / 2] == KC_TRACE) { rs.trace_ms += ms; ls.trace_ms += ms; } else { rs.shade_ms += ms; ls.shade_ms += ms; }
    }
    rs.reserved[0] = hc.q_peak;   // pool records of the largest batch (introspection; sizes PGRT_LEVEL_CAP_FACTOR)
    if (S.fused && hc.t_last > hc.t_first && hc.t_first != ~0ull) {   // phases of the (last batch's) frame kernel, microseconds
        rs.reserved[1] = (uint32_t)((hc.t_last - hc.t_first) / 1000ull);
        rs.reserved[2] = hc.t_primary_done > hc.t_first ? (uint32_t)((hc.t_primary_done - hc.t_first) / 1000ull) : 0u;
    }
    rs.reserved[3] = hc.pool_iters;
    if (stats) *stats = rs;
    return PGRT_OK;
}

static int render_frame(pgrt_context* ctx, const pgrt_render_params* p, void* dest, int dest_mode, void* host_dst, pgrt_render_stats* stats, int profile, int fmt = 0) {
    int rc = frame_begin(ctx, 0, p, dest, dest_mode, host_dst, profile, fmt);
    if (rc) return rc;
    return frame_end(ctx, 0, stats);
}

extern "C" int pgrt_render_device(pgrt_context* ctx, const pgrt_render_params* p, void* rgba_device, pgrt_render_stats* stats, int32_t profile) {
    CHECK_CTX(ctx);
    if (!rgba_device) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render_device: null destination");
    if (ctx->shard.n_ranks != 1) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render_device: context is sharded; use pgrt_render_shard_device");
    return render_frame(ctx, p, rgba_device, 0, nullptr, stats, profile);
}

extern "C" int pgrt_render(pgrt_context* ctx, const pgrt_render_params* p, float* rgba_host, pgrt_render_stats* stats, int32_t profile) {
    CHECK_CTX(ctx);
    if (!rgba_host) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render: null destination");
    if (ctx->shard.n_ranks != 1) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render: context is sharded; use pgrt_render_shard_device");
    return render_frame(ctx, p, nullptr, 0, rgba_host, stats, profile);
}

// 8-bit frames: the resolve writes R8G8B8A8_UNORM, the format the reference presents (simpleguidx11.cpp:229,290), so a frame
// that must end in host memory crosses PCIe as W*H*4 bytes instead of W*H*16
extern "C" int pgrt_render_rgba8(pgrt_context* ctx, const pgrt_render_params* p, uint8_t* rgba8_host, pgrt_render_stats* stats, int32_t profile) {
    CHECK_CTX(ctx);
    if (!rgba8_host) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render_rgba8: null destination");
    if (ctx->shard.n_ranks != 1) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render_rgba8: context is sharded; use pgrt_render_shard_to_frame_rgba8_begin");
    return render_frame(ctx, p, nullptr, 0, rgba8_host, stats, profile, 1);
}

// ---- cross-frame accumulation (the reference's Producer re-renders from scratch every iteration,
// simpleguidx11.cpp:95-118; here n finished frames with seeds seed, seed+1, ... are averaged on the device)
__global__ void __launch_bounds__(256) k_accumulate(float4* __restrict__ acc, const float4* __restrict__ frame, size_t n, int first) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 f = frame[i];
    if (first) { acc[i] = f; return; }
    const float4 a = acc[i];
    acc[i] = make_float4(a.x + f.x, a.y + f.y, a.z + f.z, a.w + f.w);   // in frame order: the sum is reproducible
}
__global__ void __launch_bounds__(256) k_accumulate_finish(float4* __restrict__ acc, size_t n, float count) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = acc[i];
    acc[i] = make_float4(a.x / count, a.y / count, a.z / count, a.w / count);
}
extern "C" int pgrt_render_accumulate(pgrt_context* ctx, const pgrt_render_params* p, int32_t n_frames, float* rgba_host, pgrt_render_stats* stats) {
    CHECK_CTX(ctx);
    if (!p || !rgba_host || n_frames < 1) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render_accumulate: bad arguments");
    if (ctx->shard.n_ranks != 1) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render_accumulate: context is sharded");
    int rc = validate_frame(ctx, p);
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    const size_t n = (size_t)ctx->cam.width * ctx->cam.height;
    const int D = std::min<int>(4, n_frames);
    sync_all_slots(ctx);
    CUDA_TRY(ctx->acc_sum.ensure(n));
    for (int s = 0; s < D; ++s) CUDA_TRY(ctx->acc_frames[s].ensure(n));
    if (!ctx->acc_event) CUDA_TRY(cudaEventCreateWithFlags(&ctx->acc_event, cudaEventDisableTiming));
    const unsigned blocks = (unsigned)div_up(n, 256);
    // whatever way this function is left, no slot it started stays marked busy (later renders use slot 0 .. 3)
    struct SlotGuard {
        pgrt_context* c; int d; bool armed = true;
        ~SlotGuard() { if (!armed) return; sync_all_slots(c); for (int s = 0; s < d; ++s) c->slots[s].busy = false; }
    } guard{ctx, D};
    for (int attempt = 0;; ++attempt) {
        pgrt_render_stats total = {}, rs;
        bool overflowed = false;
        auto collect = [&](int slot) -> int {
            int r = frame_end(ctx, slot, &rs);
            if (r) return r;
            overflowed |= rs.overflow_retries != 0;
            total.rays_primary += rs.rays_primary; total.rays_shadow += rs.rays_shadow; total.rays_reflection += rs.rays_reflection;
            total.rays_refraction += rs.rays_refraction; total.launches += rs.launches + 1; total.batches += rs.batches; total.frame_ms += rs.frame_ms;
            total.trace_launches += rs.trace_launches;
            return PGRT_OK;
        };
        for (int i = 0; i < n_frames; ++i) {
            const int slot = i % D;
            if (i >= D && (rc = collect(slot))) return rc;
            pgrt_render_params pi = *p;
            pi.seed = p->seed + (uint32_t)i;
            if ((rc = frame_begin(ctx, slot, &pi, ctx->acc_frames[slot].p, 0, nullptr, 0))) return rc;
            cudaStream_t st = ctx->slots[slot].stream;
            if (i > 0) CUDA_TRY(cudaStreamWaitEvent(st, ctx->acc_event, 0));
            k_accumulate<<<blocks, 256, 0, st>>>(ctx->acc_sum.p, ctx->acc_frames[slot].p, n, i == 0);
            LAUNCH_OK();
            CUDA_TRY(cudaEventRecord(ctx->acc_event, st));
            ctx->launches++;
        }
        for (int i = std::max(0, n_frames - D); i < n_frames; ++i)
            if ((rc = collect(i % D))) return rc;
        // a queue overflow re-renders a frame AFTER its first version was summed: start again (the smaller batch is remembered)
        if (overflowed && attempt == 0) continue;
        if (overflowed) return ctx->fail(PGRT_ERR_OVERFLOW, "pgrt_render_accumulate: secondary-ray queues keep overflowing; raise PGRT_MIN_LEVEL_CAP");
        cudaStream_t st = ctx->slots[0].stream;
        CUDA_TRY(cudaStreamWaitEvent(st, ctx->acc_event, 0));
        k_accumulate_finish<<<blocks, 256, 0, st>>>(ctx->acc_sum.p, n, (float)n_frames);
        LAUNCH_OK();
        ctx->launches++;
        CUDA_TRY(cudaMemcpyAsync(rgba_host, ctx->acc_sum.p, n * sizeof(float4), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        ctx->px_valid = false;
        if (stats) *stats = total;
        return PGRT_OK;
    }
}

// ---- pipelined frames
extern "C" int pgrt_render_begin(pgrt_context* ctx, const pgrt_render_params* p, float* rgba_host, int32_t slot, int32_t profile) {
    CHECK_CTX(ctx);
    if (!rgba_host) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render_begin: null destination");
    if (ctx->shard.n_ranks != 1) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render_begin: context is sharded; use pgrt_render_shard_device_begin");
    return frame_begin(ctx, slot, p, nullptr, 0, rgba_host, profile);
}
extern "C" int pgrt_render_device_begin(pgrt_context* ctx, const pgrt_render_params* p, void* rgba_device, int32_t slot, int32_t profile) {
    CHECK_CTX(ctx);
    if (!rgba_device) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render_device_begin: null destination");
    if (ctx->shard.n_ranks != 1) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render_device_begin: context is sharded; use pgrt_render_shard_device_begin");
    return frame_begin(ctx, slot, p, rgba_device, 0, nullptr, profile);
}
extern "C" int pgrt_render_shard_device_begin(pgrt_context* ctx, const pgrt_render_params* p, void* shard_rgba_device, int32_t slot, int32_t profile) {
    CHECK_CTX(ctx);
    if (!shard_rgba_device) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render_shard_device_begin: null destination");
    return frame_begin(ctx, slot, p, shard_rgba_device, 1, nullptr, profile);
}
extern "C" int pgrt_render_shard_to_frame_begin(pgrt_context* ctx, const pgrt_render_params* p, void* frame_device, int32_t slot, int32_t profile) {
    CHECK_CTX(ctx);
    if (!frame_device) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render_shard_to_frame_begin: null destination");
    return frame_begin(ctx, slot, p, frame_device, 0, nullptr, profile);   // full-frame addressing, this rank's tiles only
}
extern "C" int pgrt_render_rgba8_begin(pgrt_context* ctx, const pgrt_render_params* p, uint8_t* rgba8_host, int32_t slot, int32_t profile) {
    CHECK_CTX(ctx);
    if (!rgba8_host) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render_rgba8_begin: null destination");
    if (ctx->shard.n_ranks != 1) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render_rgba8_begin: context is sharded; use pgrt_render_shard_to_frame_rgba8_begin");
    return frame_begin(ctx, slot, p, nullptr, 0, rgba8_host, profile, 1);
}
extern "C" int pgrt_render_shard_to_frame_rgba8_begin(pgrt_context* ctx, const pgrt_render_params* p, void* frame_rgba8_device, int32_t slot, int32_t profile) {
    CHECK_CTX(ctx);
    if (!frame_rgba8_device) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render_shard_to_frame_rgba8_begin: null destination");
    return frame_begin(ctx, slot, p, frame_rgba8_device, 0, nullptr, profile, 1);
}
// ---- completion flags (no collective): the next frames begun on `slot` store `value` in *flag once they have finished
// without a queue overflow (k_batch_end).  `flag` is any word the device can write: its own memory, a peer-mapped frame
// header (CUDA IPC), registered host memory.  NULL switches the signal off.
extern "C" int pgrt_slot_signal(pgrt_context* ctx, int32_t slot, void* flag_device, uint32_t value) {
    CHECK_CTX(ctx);
    if (slot < 0 || slot >= PGRT_MAX_INFLIGHT) return ctx->fail(PGRT_ERR_INVALID, "pgrt_slot_signal: slot out of range");
    if (ctx->slots[slot].busy) return ctx->fail(PGRT_ERR_INVALID, "pgrt_slot_signal: the slot holds a frame in flight");
    ctx->slots[slot].sig_flag = (uint32_t*)flag_device; ctx->slots[slot].sig_value = value; ctx->slots[slot].sig_add = false;
    return PGRT_OK;
}
// the same with a COUNTER shared by all ranks: a finished frame adds 1 to *counter_device (system-scope atomic; the counter may live
// on another GPU behind NVLink), so that one pgrt_stream_wait_value32 on n_ranks * frames covers every rank
extern "C" int pgrt_slot_signal_add(pgrt_context* ctx, int32_t slot, void* counter_device) {
    CHECK_CTX(ctx);
    if (slot < 0 || slot >= PGRT_MAX_INFLIGHT) return ctx->fail(PGRT_ERR_INVALID, "pgrt_slot_signal_add: slot out of range");
    if (ctx->slots[slot].busy) return ctx->fail(PGRT_ERR_INVALID, "pgrt_slot_signal_add: the slot holds a frame in flight");
    ctx->slots[slot].sig_flag = (uint32_t*)counter_device; ctx->slots[slot].sig_value = 0; ctx->slots[slot].sig_add = counter_device != nullptr;
    return PGRT_OK;
}
// make `cuda_stream` wait until *flag >= value (cuStreamWaitValue32, GEQ) / store value in *flag in stream order
extern "C" int pgrt_stream_wait_value32(pgrt_context* ctx, void* cuda_stream, void* flag_device, uint32_t value) {
    CHECK_CTX(ctx);
    if (!flag_device) return ctx->fail(PGRT_ERR_INVALID, "pgrt_stream_wait_value32: null flag");
    if (!ctx->fn_wait32) return ctx->fail(PGRT_ERR_CUDA, "pgrt_stream_wait_value32: the driver does not export cuStreamWaitValue32");
    cudaSetDevice(ctx->device);
    const int rc = ((pgrt_cu_memop32)ctx->fn_wait32)((cudaStream_t)cuda_stream, (unsigned long long)(uintptr_t)flag_device, value, 0x0u /* CU_STREAM_WAIT_VALUE_GEQ */);
    if (rc != 0) return ctx->fail(PGRT_ERR_CUDA, "pgrt_stream_wait_value32: cuStreamWaitValue32 failed with CUresult " + std::to_string(rc));
    return PGRT_OK;
}
extern "C" int pgrt_stream_write_value32(pgrt_context* ctx, void* cuda_stream, void* flag_device, uint32_t value) {
    CHECK_CTX(ctx);
    if (!flag_device) return ctx->fail(PGRT_ERR_INVALID, "pgrt_stream_write_value32: null flag");
    cudaSetDevice(ctx->device);
    return stream_write32(ctx, (cudaStream_t)cuda_stream, flag_device, value);
}
// ---- frames shared between the processes of one box (one process per GPU): CUDA IPC over NVLink
extern "C" int pgrt_frame_alloc(pgrt_context* ctx, uint64_t bytes, void** device_ptr) {
    CHECK_CTX(ctx);
    if (!device_ptr || !bytes) return ctx->fail(PGRT_ERR_INVALID, "pgrt_frame_alloc: bad arguments");
    cudaSetDevice(ctx->device);
    CUDA_TRY(cudaMalloc(device_ptr, bytes));
    CUDA_TRY(cudaMemset(*device_ptr, 0, bytes));
    return PGRT_OK;
}
extern "C" int pgrt_frame_free(pgrt_context* ctx, void* device_ptr) {
    CHECK_CTX(ctx);
    cudaSetDevice(ctx->device);
    sync_all_slots(ctx);
    CUDA_TRY(cudaFree(device_ptr));
    return PGRT_OK;
}
extern "C" int pgrt_host_frame_register(pgrt_context* ctx, void* host, uint64_t bytes, void** device_ptr) {
    CHECK_CTX(ctx);
    if (!host || !bytes || !device_ptr || ((uintptr_t)host & 4095u)) return ctx->fail(PGRT_ERR_INVALID, "pgrt_host_frame_register: need a page-aligned host buffer");
    cudaSetDevice(ctx->device);
    CUDA_TRY(cudaHostRegister(host, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    cudaError_t e = cudaHostGetDevicePointer(device_ptr, host, 0);
    if (e != cudaSuccess) { cudaHostUnregister(host); CUDA_TRY(e); }
    return PGRT_OK;
}
extern "C" int pgrt_host_frame_unregister(pgrt_context* ctx, void* host) {
    CHECK_CTX(ctx);
    cudaSetDevice(ctx->device);
    sync_all_slots(ctx);
    CUDA_TRY(cudaHostUnregister(host));
    return PGRT_OK;
}
extern "C" int pgrt_frame_export(pgrt_context* ctx, void* device_ptr, uint8_t handle[64]) {
    CHECK_CTX(ctx);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaSetDevice(ctx->device);
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, device_ptr));
    memcpy(handle, &h, 64);
    return PGRT_OK;
}
extern "C" int pgrt_frame_import(pgrt_context* ctx, const uint8_t handle[64], void** device_ptr) {
    CHECK_CTX(ctx);
    if (!device_ptr) return ctx->fail(PGRT_ERR_INVALID, "pgrt_frame_import: null pointer");
    cudaSetDevice(ctx->device);   // opened from THIS device: peer access to the owner is enabled lazily by the driver
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PGRT_OK;
}
extern "C" int pgrt_frame_unmap(pgrt_context* ctx, void* device_ptr) {
    CHECK_CTX(ctx);
    cudaSetDevice(ctx->device);
    sync_all_slots(ctx);
    CUDA_TRY(cudaIpcCloseMemHandle(device_ptr));
    return PGRT_OK;
}

// measurement helper: evict L2 by streaming `bytes` of writes through a scratch buffer, on the slot's stream
__global__ void __launch_bounds__(256) k_l2_flush(uint4* __restrict__ buf, size_t n, uint32_t v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = make_uint4(v, v, v, v);
}
extern "C" int pgrt_debug_flush_l2(pgrt_context* ctx, int32_t slot, uint64_t bytes, uint32_t value) {
    CHECK_CTX(ctx);
    if (slot < 0 || slot >= PGRT_MAX_INFLIGHT || bytes < 16) return ctx->fail(PGRT_ERR_INVALID, "pgrt_debug_flush_l2: bad arguments");
    cudaSetDevice(ctx->device);
    if (ctx->flush_buf.n < bytes / 16) { sync_all_slots(ctx); CUDA_TRY(ctx->flush_buf.ensure(bytes / 16)); }
    // a grid that leaves room on every SM: the fill runs beside the other frames' kernels instead of displacing them
    static const int flush_ctas = getenv("PGRT_FLUSH_CTAS_PER_SM") ? std::max(1, atoi(getenv("PGRT_FLUSH_CTAS_PER_SM"))) : 2;
    k_l2_flush<<<ctx->sm_count * flush_ctas, 256, 0, ctx->slots[slot].stream>>>(ctx->flush_buf.p, bytes / 16, value);
    ctx->launches++;
    LAUNCH_OK();
    return PGRT_OK;
}

extern "C" int pgrt_debug_frame_cycles(pgrt_context* ctx, int32_t slot, uint64_t out[4]) {
    CHECK_CTX(ctx);
    if (slot < 0 || slot >= PGRT_MAX_INFLIGHT || !out) return ctx->fail(PGRT_ERR_INVALID, "pgrt_debug_frame_cycles: bad arguments");
    if (ctx->slots[slot].busy) return ctx->fail(PGRT_ERR_INVALID, "pgrt_debug_frame_cycles: the slot holds a frame in flight");
    for (int i = 0; i < 4; ++i) out[i] = ctx->slots[slot].h_counters->warp_cycles[i];
    return PGRT_OK;
}

// measurement helper: read bandwidth of an L2-resident buffer from all SMs (the memory roof of scenes whose nodes and
// triangles fit L2).  ld.global.cg: the loads bypass L1, so every byte crosses the L2 -> SM fabric.
__global__ void __launch_bounds__(256) k_l2_read(const uint4* __restrict__ buf, size_t n, int iters, uint32_t* __restrict__ sink) {
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (int it = 0; it < iters; ++it)
        for (size_t i = i0; i + 3 * stride < n; i += 4 * stride) {     // four independent 16-byte loads in flight per thread
            const uint4 a = __ldcg(buf + i), b = __ldcg(buf + i + stride), c = __ldcg(buf + i + 2 * stride), d = __ldcg(buf + i + 3 * stride);
            acc += a.x ^ b.y ^ c.z ^ d.w;
        }
    if (acc == 0x9E3779B9u) *sink = acc;      // keeps the loads alive; practically never true
}
extern "C" int pgrt_debug_l2_bandwidth(pgrt_context* ctx, uint64_t bytes, int32_t iters, float* gb_per_s) {
    CHECK_CTX(ctx);
    if (!gb_per_s || bytes < (1u << 20) || iters < 1) return ctx->fail(PGRT_ERR_INVALID, "pgrt_debug_l2_bandwidth: bad arguments");
    cudaSetDevice(ctx->device);
    sync_all_slots(ctx);
    const size_t n = bytes / 16;
    CUDA_TRY(ctx->flush_buf.ensure(n + 1));
    cudaStream_t st = ctx->stream;
    CUDA_TRY(cudaMemsetAsync(ctx->flush_buf.p, 1, n * 16, st));
    const int grid = ctx->sm_count * 8;
    const size_t per_iter = (n / ((size_t)grid * 256 * 4)) * ((size_t)grid * 256 * 4);     // elements one pass really reads
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    k_l2_read<<<grid, 256, 0, st>>>(ctx->flush_buf.p, n, 2, (uint32_t*)(ctx->flush_buf.p + n));      // warm: the buffer becomes L2-resident
    cudaEventRecord(e0, st);
    k_l2_read<<<grid, 256, 0, st>>>(ctx->flush_buf.p, n, iters, (uint32_t*)(ctx->flush_buf.p + n));
    cudaEventRecord(e1, st);
    ctx->launches += 2;
    cudaError_t e = cudaStreamSynchronize(st);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CUDA_TRY(e);
    *gb_per_s = ms > 0.f ? (float)((double)per_iter * 16.0 * iters / (ms * 1e-3) / 1e9) : 0.f;
    return PGRT_OK;
}

extern "C" int pgrt_enable_peer_access(pgrt_context* ctx, int32_t peer_device) {
    CHECK_CTX(ctx);
    cudaSetDevice(ctx->device);
    if (peer_device == ctx->device) return PGRT_OK;
    int can = 0;
    CUDA_TRY(cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
    if (!can) return ctx->fail(PGRT_ERR_INVALID, "pgrt_enable_peer_access: the two devices have no peer path");
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return PGRT_OK; }
    CUDA_TRY(e);
    return PGRT_OK;
}
extern "C" int pgrt_render_end(pgrt_context* ctx, int32_t slot, pgrt_render_stats* stats) {
    CHECK_CTX(ctx);
    return frame_end(ctx, slot, stats);
}
extern "C" int pgrt_stream_wait_slot(pgrt_context* ctx, int32_t slot, void* cuda_stream) {
    CHECK_CTX(ctx);
    if (slot < 0 || slot >= PGRT_MAX_INFLIGHT) return ctx->fail(PGRT_ERR_INVALID, "pgrt_stream_wait_slot: slot out of range");
    cudaSetDevice(ctx->device);
    FrameSlot& S = ctx->slots[slot];
    // The slot's completion count only moves for a frame that finished WITHOUT a queue overflow: a consumer ordered here is
    // not released by an attempt whose retry (pgrt_render_end) is still to come.  Without stream memory operations the
    // event of the last enqueue is all there is; then a frame is only known to be good once pgrt_render_end has returned.
    if (ctx->fn_wait32 && ((pgrt_cu_memop32)ctx->fn_wait32)((cudaStream_t)cuda_stream, (unsigned long long)(uintptr_t)&S.d_counters.p->done_seq, S.seq_expected, 0x0u /* GEQ */) == 0) return PGRT_OK;
    CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)cuda_stream, S.ev_done, 0));
    return PGRT_OK;
}

extern "C" int pgrt_get_pixel(pgrt_context* ctx, const pgrt_render_params* p, int32_t x, int32_t y, float rgba[4]) {
    CHECK_CTX(ctx);
    if (!p || !rgba || !ctx->cam_set || x < 0 || y < 0 || x >= ctx->cam.width || y >= ctx->cam.height) return ctx->fail(PGRT_ERR_INVALID, "pgrt_get_pixel: bad arguments");
    if (!ctx->px_valid || memcmp(&ctx->px_params, p, sizeof *p) != 0) {
        ctx->px_cache.resize((size_t)ctx->cam.width * ctx->cam.height * 4);
        int rc = pgrt_render(ctx, p, ctx->px_cache.data(), nullptr, 0);
        if (rc) return rc;
        ctx->px_params = *p; ctx->px_valid = true;
    }
    memcpy(rgba, &ctx->px_cache[((size_t)y * ctx->cam.width + x) * 4], 16);
    return PGRT_OK;
}

extern "C" int pgrt_primary_ids(pgrt_context* ctx, const pgrt_render_params* p, uint32_t* geom_host, uint32_t* prim_host) {
    CHECK_CTX(ctx);
    if (!geom_host || !prim_host) return ctx->fail(PGRT_ERR_INVALID, "pgrt_primary_ids: null destination");
    if (!ctx->cam_set) return ctx->fail(PGRT_ERR_INVALID, "render: pgrt_set_camera has not been called");
    cudaSetDevice(ctx->device);
    const size_t npx = (size_t)ctx->cam.width * ctx->cam.height;
    CUDA_TRY(ctx->d_ids.ensure(2 * npx));
    CUDA_TRY(cudaMemsetAsync(ctx->d_ids.p, 0xFF, 2 * npx * sizeof(uint32_t), ctx->stream));
    int rc = render_frame(ctx, p, nullptr, 2, nullptr, nullptr, 0);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(geom_host, ctx->d_ids.p, npx * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(prim_host, ctx->d_ids.p + npx, npx * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PGRT_OK;
}

extern "C" int pgrt_render_shard_device(pgrt_context* ctx, const pgrt_render_params* p, void* shard_rgba_device, pgrt_render_stats* stats, int32_t profile) {
    CHECK_CTX(ctx);
    if (!shard_rgba_device) return ctx->fail(PGRT_ERR_INVALID, "pgrt_render_shard_device: null destination");
    return render_frame(ctx, p, shard_rgba_device, 1, nullptr, stats, profile);
}

static int untile_on(pgrt_context* ctx, const void* gathered_device, int32_t n_ranks, void* rgba_device, cudaStream_t st);

extern "C" int pgrt_untile(pgrt_context* ctx, const void* gathered_device, int32_t n_ranks, void* rgba_device) {
    CHECK_CTX(ctx);
    return untile_on(ctx, gathered_device, n_ranks, rgba_device, ctx->stream);
}
extern "C" int pgrt_untile_on_stream(pgrt_context* ctx, const void* gathered_device, int32_t n_ranks, void* rgba_device, void* cuda_stream) {
    CHECK_CTX(ctx);
    return untile_on(ctx, gathered_device, n_ranks, rgba_device, (cudaStream_t)cuda_stream);
}

static int untile_on(pgrt_context* ctx, const void* gathered_device, int32_t n_ranks, void* rgba_device, cudaStream_t st) {
    if (!gathered_device || !rgba_device || n_ranks < 1 || !ctx->cam_set) return ctx->fail(PGRT_ERR_INVALID, "pgrt_untile: bad arguments");
    cudaSetDevice(ctx->device);
    const uint64_t tiles = (uint64_t)ctx->shard.tiles_x * ctx->shard.tiles_y;
    const uint32_t spr = (uint32_t)((tiles + n_ranks - 1) / n_ranks * PGRT_TILE_PIXELS);
    k_untile<<<div_up((size_t)spr * n_ranks, 256), 256, 0, st>>>(ctx->cam, n_ranks, ctx->shard.tiles_x, ctx->shard.tiles_y, spr, (const float4*)gathered_device, (float4*)rgba_device);
    ctx->launches++;
    LAUNCH_OK();
    return PGRT_OK;
}

// --------------------------------------------------------------------------------------------------- batch queries
extern "C" int pgrt_intersect(pgrt_context* ctx, pgrt_rayhit* rh, uint64_t n) {
    CHECK_CTX(ctx);
    if (!ctx->committed) return ctx->fail(PGRT_ERR_INVALID, "pgrt_intersect: scene not committed");
    if (n == 0) return PGRT_OK;
    if (!rh) return ctx->fail(PGRT_ERR_INVALID, "pgrt_intersect: null rays");
    int rc = upload_tables(ctx); if (rc) return rc;
    cudaSetDevice(ctx->device);
    DevBuf<pgrt_rayhit> d;
    CUDA_TRY(d.ensure(n));
    cudaError_t e = cudaMemcpyAsync(d.p, rh, n * sizeof(pgrt_rayhit), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        k_intersect<<<std::min<unsigned>(div_up(n, 128), ctx->sm_count * 16), 128, 0, ctx->stream>>>(ctx->dev_scene(), ctx->d_pos.p, d.p, n);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(rh, d.p, n * sizeof(pgrt_rayhit), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    d.release();
    if (e != cudaSuccess) return ctx->fail_cuda(e, "pgrt_intersect", __FILE__, __LINE__);
    return PGRT_OK;
}

namespace {
// small RAII helper for the eval entry points: host -> device -> kernel -> host
struct Scratch {
    std::vector<void*> ptrs;
    ~Scratch() { for (void* p : ptrs) cudaFree(p); }
    void* up(const void* h, size_t bytes, cudaStream_t st) {
        void* d = nullptr;
        if (cudaMalloc(&d, std::max<size_t>(bytes, 16)) != cudaSuccess) return nullptr;
        ptrs.push_back(d);
        if (h && cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return nullptr;
        return d;
    }
};
}  // namespace

#define EVAL_FINISH(dptr, hptr, bytes)                                                                  \
    do {                                                                                                \
        ctx->launches++;                                                                                \
        CUDA_TRY(cudaGetLastError());                                                                   \
        CUDA_TRY(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));              \
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));                                                   \
        return PGRT_OK;                                                                                 \
    } while (0)

extern "C" int pgrt_interpolate(pgrt_context* ctx, const uint32_t* geom, const uint32_t* prim, const float* u, const float* v, uint64_t n, int32_t slot, float* out) {
    CHECK_CTX(ctx);
    if (!ctx->committed || !geom || !prim || !u || !v || !out || (slot != 0 && slot != 1)) return ctx->fail(PGRT_ERR_INVALID, "pgrt_interpolate: bad arguments");
    for (uint64_t i = 0; i < n; ++i) {
        if (geom[i] >= ctx->h_geom_first.size()) return ctx->fail(PGRT_ERR_INVALID, "pgrt_interpolate: geomID out of range");
        const uint32_t end = geom[i] + 1 < ctx->h_geom_first.size() ? ctx->h_geom_first[geom[i] + 1] : ctx->n_tris;
        if (ctx->h_geom_first[geom[i]] + prim[i] >= end) return ctx->fail(PGRT_ERR_INVALID, "pgrt_interpolate: primID out of range");
    }
    cudaSetDevice(ctx->device);
    Scratch s; const size_t w = slot == 0 ? 3 : 2;
    void* dg = s.up(geom, n * 4, ctx->stream); void* dp = s.up(prim, n * 4, ctx->stream); void* du = s.up(u, n * 4, ctx->stream); void* dv = s.up(v, n * 4, ctx->stream);
    void* dout = s.up(nullptr, n * w * 4, ctx->stream);
    if (!dg || !dp || !du || !dv || !dout) return ctx->fail(PGRT_ERR_CUDA, "pgrt_interpolate: allocation failed");
    k_interpolate<<<div_up(n, 256), 256, 0, ctx->stream>>>(ctx->dev_scene(), (uint32_t*)dg, (uint32_t*)dp, (float*)du, (float*)dv, n, slot, (float*)dout);
    EVAL_FINISH(dout, out, n * w * 4);
}

// Raytracer::trace / Raytracer::is_illuminated over batches (the two public query methods of the class, raytracer.h:31,34)
extern "C" int pgrt_trace(pgrt_context* ctx, const pgrt_render_params* p, const pgrt_ray* rays, uint64_t n, int32_t level, float* rgba) {
    CHECK_CTX(ctx);
    if (n == 0) return PGRT_OK;
    if (!p || !rays || !rgba || level < 0) return ctx->fail(PGRT_ERR_INVALID, "pgrt_trace: bad arguments");
    if (!ctx->committed) return ctx->fail(PGRT_ERR_INVALID, "pgrt_trace: scene not committed");
    if (p->max_depth < 0 || p->max_depth >= PGRT_MAX_LEVELS) return ctx->fail(PGRT_ERR_INVALID, "pgrt_trace: max_depth out of range [0,32]");
    if (p->shadow_mode != 0 && p->shadow_mode != 1) return ctx->fail(PGRT_ERR_INVALID, "pgrt_trace: shadow_mode must be 0 or 1");
    int rc = upload_tables(ctx); if (rc) return rc;
    cudaSetDevice(ctx->device);
    static_assert(sizeof(pgrt_ray) == 48, "RTCRay layout");
    const uint64_t chunk = std::min<uint64_t>(n, 1u << 17);   // one explicit recursion stack per ray in flight
    Scratch s;
    void* d_rays = s.up(nullptr, chunk * sizeof(pgrt_ray), ctx->stream); void* d_out = s.up(nullptr, chunk * 16, ctx->stream);
    void* d_stk = s.up(nullptr, chunk * (PGRT_MAX_LEVELS + 1) * sizeof(TraceFrame), ctx->stream);
    if (!d_rays || !d_out || !d_stk) return ctx->fail(PGRT_ERR_CUDA, "pgrt_trace: allocation failed");
    const DevScene sc = ctx->dev_scene();
    for (uint64_t i0 = 0; i0 < n; i0 += chunk) {
        const uint64_t m = std::min(chunk, n - i0);
        CUDA_TRY(cudaMemcpyAsync(d_rays, rays + i0, m * sizeof(pgrt_ray), cudaMemcpyHostToDevice, ctx->stream));
        if (p->shader_mode == 3) k_trace_rays<true><<<div_up(m, 128), 128, 0, ctx->stream>>>(sc, *p, (const pgrt_ray*)d_rays, m, level, (float4*)d_out, (TraceFrame*)d_stk, ctx->slots[0].d_counters.p);
        else k_trace_rays<false><<<div_up(m, 128), 128, 0, ctx->stream>>>(sc, *p, (const pgrt_ray*)d_rays, m, level, (float4*)d_out, (TraceFrame*)d_stk, ctx->slots[0].d_counters.p);
        ctx->launches++;
        LAUNCH_OK();
        CUDA_TRY(cudaMemcpyAsync(rgba + 4 * i0, d_out, m * 16, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return PGRT_OK;
}

extern "C" int pgrt_is_illuminated(pgrt_context* ctx, const pgrt_render_params* p, const float* light, const float* hit, const float* nrm, uint64_t n, int32_t* lit) {
    CHECK_CTX(ctx);
    if (n == 0) return PGRT_OK;
    if (!p || !light || !hit || !nrm || !lit) return ctx->fail(PGRT_ERR_INVALID, "pgrt_is_illuminated: bad arguments");
    if (!ctx->committed) return ctx->fail(PGRT_ERR_INVALID, "pgrt_is_illuminated: scene not committed");
    if (p->shadow_mode != 0 && p->shadow_mode != 1) return ctx->fail(PGRT_ERR_INVALID, "pgrt_is_illuminated: shadow_mode must be 0 or 1");
    int rc = upload_tables(ctx); if (rc) return rc;
    cudaSetDevice(ctx->device);
    Scratch s;
    void* dl = s.up(light, n * 12, ctx->stream); void* dh = s.up(hit, n * 12, ctx->stream); void* dn = s.up(nrm, n * 12, ctx->stream); void* o = s.up(nullptr, n * 4, ctx->stream);
    if (!dl || !dh || !dn || !o) return ctx->fail(PGRT_ERR_CUDA, "pgrt_is_illuminated: allocation failed");
    k_is_illuminated<<<div_up(n, 128), 128, 0, ctx->stream>>>(ctx->dev_scene(), *p, (const float*)dl, (const float*)dh, (const float*)dn, n, (int32_t*)o);
    EVAL_FINISH(o, lit, n * 4);
}

extern "C" int pgrt_eval_mix_srgb(pgrt_context* ctx, const float* c0, const float* c1, const float* alpha, uint64_t n, float* out) {
    CHECK_CTX(ctx);
    if (!c0 || !c1 || !alpha || !out) return ctx->fail(PGRT_ERR_INVALID, "pgrt_eval_mix_srgb: null buffer");
    cudaSetDevice(ctx->device);
    Scratch s;
    void* a = s.up(c0, n * 16, ctx->stream); void* b = s.up(c1, n * 16, ctx->stream); void* al = s.up(alpha, n * 4, ctx->stream); void* o = s.up(nullptr, n * 16, ctx->stream);
    if (!a || !b || !al || !o) return ctx->fail(PGRT_ERR_CUDA, "pgrt_eval_mix_srgb: allocation failed");
    k_eval_mix<<<div_up(n, 256), 256, 0, ctx->stream>>>((float4*)a, (float4*)b, (float*)al, n, (float4*)o);
    EVAL_FINISH(o, out, n * 16);
}

extern "C" int pgrt_eval_texture(pgrt_context* ctx, int32_t tex_id, const float* uv, uint64_t n, float* out3) {
    CHECK_CTX(ctx);
    if (!uv || !out3) return ctx->fail(PGRT_ERR_INVALID, "pgrt_eval_texture: null buffer");
    const HostTex* t = tex_id < 0 ? &ctx->env : ((size_t)tex_id < ctx->textures.size() ? &ctx->textures[tex_id] : nullptr);
    if (!t || !t->set) return ctx->fail(PGRT_ERR_INVALID, "pgrt_eval_texture: no such texture");
    cudaSetDevice(ctx->device);
    DevTexture dt; dt.data = t->bytes.p; dt.width = t->w; dt.height = t->h; dt.pitch = t->pitch; dt.bpp = t->bpp;
    Scratch s;
    void* d = s.up(uv, n * 8, ctx->stream); void* o = s.up(nullptr, n * 12, ctx->stream);
    if (!d || !o) return ctx->fail(PGRT_ERR_CUDA, "pgrt_eval_texture: allocation failed");
    k_eval_texture<<<div_up(n, 256), 256, 0, ctx->stream>>>(dt, (float2*)d, n, (float*)o);
    EVAL_FINISH(o, out3, n * 12);
}

extern "C" int pgrt_eval_envmap(pgrt_context* ctx, const float* dirs, uint64_t n, float* out4) {
    CHECK_CTX(ctx);
    if (!dirs || !out4) return ctx->fail(PGRT_ERR_INVALID, "pgrt_eval_envmap: null buffer");
    cudaSetDevice(ctx->device);
    DevTexture dt; dt.data = ctx->env.set ? ctx->env.bytes.p : nullptr; dt.width = ctx->env.w; dt.height = ctx->env.h; dt.pitch = ctx->env.pitch; dt.bpp = ctx->env.bpp;
    Scratch s;
    void* d = s.up(dirs, n * 12, ctx->stream); void* o = s.up(nullptr, n * 16, ctx->stream);
    if (!d || !o) return ctx->fail(PGRT_ERR_CUDA, "pgrt_eval_envmap: allocation failed");
    k_eval_env<<<div_up(n, 256), 256, 0, ctx->stream>>>(dt, (float*)d, n, (float4*)o);
    EVAL_FINISH(o, out4, n * 16);
}

extern "C" int pgrt_eval_gamma(pgrt_context* ctx, const float* in4, float gamma_level, uint64_t n, float* out4) {
    CHECK_CTX(ctx);
    if (!in4 || !out4) return ctx->fail(PGRT_ERR_INVALID, "pgrt_eval_gamma: null buffer");
    cudaSetDevice(ctx->device);
    Scratch s;
    void* d = s.up(in4, n * 16, ctx->stream); void* o = s.up(nullptr, n * 16, ctx->stream);
    if (!d || !o) return ctx->fail(PGRT_ERR_CUDA, "pgrt_eval_gamma: allocation failed");
    k_eval_gamma<<<div_up(n, 256), 256, 0, ctx->stream>>>((float4*)d, gamma_level, n, (float4*)o);
    EVAL_FINISH(o, out4, n * 16);
}

extern "C" int pgrt_eval_primary_rays(pgrt_context* ctx, const pgrt_render_params* p, float* out9) {
    CHECK_CTX(ctx);
    if (!p || !out9 || !ctx->cam_set || p->sampling_width < 1) return ctx->fail(PGRT_ERR_INVALID, "pgrt_eval_primary_rays: bad arguments");
    cudaSetDevice(ctx->device);
    const uint64_t n = (uint64_t)ctx->cam.width * ctx->cam.height * p->sampling_width * p->sampling_width;
    Scratch s;
    void* o = s.up(nullptr, n * 36, ctx->stream);
    if (!o) return ctx->fail(PGRT_ERR_CUDA, "pgrt_eval_primary_rays: allocation failed");
    k_eval_primary<<<div_up(n, 256), 256, 0, ctx->stream>>>(ctx->cam, *p, (float*)o);
    EVAL_FINISH(o, out9, n * 36);
}

extern "C" int pgrt_eval_secondary_rays(pgrt_context* ctx, const float* in11, uint64_t n, int32_t refraction, float* out9) {
    CHECK_CTX(ctx);
    if (!in11 || !out9) return ctx->fail(PGRT_ERR_INVALID, "pgrt_eval_secondary_rays: null buffer");
    cudaSetDevice(ctx->device);
    Scratch s;
    void* d = s.up(in11, n * 44, ctx->stream); void* o = s.up(nullptr, n * 36, ctx->stream);
    if (!d || !o) return ctx->fail(PGRT_ERR_CUDA, "pgrt_eval_secondary_rays: allocation failed");
    k_eval_secondary<<<div_up(n, 256), 256, 0, ctx->stream>>>((float*)d, n, refraction, (float*)o);
    EVAL_FINISH(o, out9, n * 36);
}

// --------------------------------------------------------------------------------------------------- introspection
extern "C" int pgrt_last_level_stats(const pgrt_context* ctx, int32_t level, pgrt_level_stats* out) {
    if (!ctx || !out || level < 0 || level >= ctx->level_stats_n) return PGRT_ERR_INVALID;
    *out = ctx->level_stats[level];
    return PGRT_OK;
}
extern "C" uint32_t pgrt_num_triangles(const pgrt_context* ctx) { return ctx ? ctx->n_added : 0; }
extern "C" uint32_t pgrt_num_geometries(const pgrt_context* ctx) { return ctx ? (uint32_t)ctx->h_geom_first.size() : 0; }
extern "C" uint64_t pgrt_kernel_launches(const pgrt_context* ctx) { return ctx ? ctx->launches : 0; }
