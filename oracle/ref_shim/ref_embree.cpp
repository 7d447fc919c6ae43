// ref_embree.cpp -- TEST INFRASTRUCTURE ONLY.  The Embree 3 entry points the reference calls (pg1/raytracer.cpp:26-47,
// :71-127, :130-148, :168-169, :249-253, :344; pg1/tutorials.cpp:32-164), defined over the CPU oracle's intersector
// (oracle/pg_oracle.cpp, orc_* C API).  Embree's binary is not vendored by the reference (pg1_embree.vcxproj:79,87,157), so
// when the reference's own sources are compiled into oracle/_ref/libpg_ref.so THIS is the one part that is not the reference's
// object code: the closest-hit query and rtcInterpolate follow the oracle's restatement of Embree's published forms.
// Everything above that boundary -- trace(), is_illuminated(), the ray makers, mix_srgb, the texture and environment look-ups,
// the camera, the OBJ/MTL loader -- is the reference's own code.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <sys/types.h>
#include <embree3/rtcore.h>

extern "C" {
// oracle/pg_oracle.cpp
struct orc_rayhit {
    float org_x, org_y, org_z, tnear, dir_x, dir_y, dir_z, time, tfar; uint32_t mask, id, flags;
    float Ng_x, Ng_y, Ng_z, u, v; uint32_t primID, geomID, instID;
};
void* orc_create();
void orc_destroy(void* h);
int orc_add_mesh(void* h, const float* pos, const float* nrm, const float* uv, uint32_t T, int32_t material_id, uint32_t* geom_id);
int orc_commit(void* h);
int orc_intersect(void* h, orc_rayhit* rh, uint64_t n, int brute, int threads);
int orc_interpolate(void* h, const uint32_t* geom, const uint32_t* prim, const float* u, const float* v, uint64_t n, int slot, float* out);
}
static_assert(sizeof(orc_rayhit) == sizeof(RTCRayHit), "RTCRayHit layout");

namespace {
struct Buffer { std::vector<uint8_t> bytes; size_t stride = 0, count = 0; RTCFormat format = RTC_FORMAT_UNDEFINED; };
struct Device { int refs = 1; RTCErrorFunction on_error = nullptr; void* user = nullptr; };
struct Scene;
struct Geometry {
    int refs = 1; Device* dev = nullptr; void* user_data = nullptr;
    Buffer vertex, index; std::vector<Buffer> attr;
    Scene* scene = nullptr; unsigned id = RTC_INVALID_GEOMETRY_ID;
};
struct Scene { int refs = 1; Device* dev = nullptr; std::vector<Geometry*> geoms; void* orc = nullptr; int brute = 0; };
void release(Geometry* g) { if (g && --g->refs == 0) delete g; }
}

RTC_API RTCDevice rtcNewDevice(const char*) { return (RTCDevice) new Device(); }
RTC_API void rtcReleaseDevice(RTCDevice d) { Device* x = (Device*)d; if (x && --x->refs == 0) delete x; }
RTC_API ssize_t rtcGetDeviceProperty(RTCDevice, enum RTCDeviceProperty prop) { return prop == RTC_DEVICE_PROPERTY_TRIANGLE_GEOMETRY_SUPPORTED ? 1 : 0; }
RTC_API enum RTCError rtcGetDeviceError(RTCDevice) { return RTC_ERROR_NONE; }
RTC_API void rtcSetDeviceErrorFunction(RTCDevice d, RTCErrorFunction f, void* user) { Device* x = (Device*)d; x->on_error = f; x->user = user; }

RTC_API RTCScene rtcNewScene(RTCDevice d) {
    Scene* s = new Scene(); s->dev = (Device*)d; s->orc = orc_create();
    if (const char* e = getenv("PG_REF_BRUTE")) s->brute = atoi(e);
    return (RTCScene)s;
}
RTC_API void rtcReleaseScene(RTCScene sc) {
    Scene* s = (Scene*)sc;
    if (!s || --s->refs) return;
    for (Geometry* g : s->geoms) release(g);
    orc_destroy(s->orc);
    delete s;
}
RTC_API RTCGeometry rtcNewGeometry(RTCDevice d, enum RTCGeometryType) { Geometry* g = new Geometry(); g->dev = (Device*)d; return (RTCGeometry)g; }
RTC_API void rtcReleaseGeometry(RTCGeometry g) { release((Geometry*)g); }
RTC_API void rtcCommitGeometry(RTCGeometry) {}
RTC_API void rtcSetGeometryVertexAttributeCount(RTCGeometry g, unsigned n) { ((Geometry*)g)->attr.resize(n); }
RTC_API void* rtcSetNewGeometryBuffer(RTCGeometry gg, enum RTCBufferType type, unsigned slot, enum RTCFormat format, size_t stride, size_t count) {
    Geometry* g = (Geometry*)gg;
    Buffer* b = nullptr;
    if (type == RTC_BUFFER_TYPE_VERTEX) b = &g->vertex;
    else if (type == RTC_BUFFER_TYPE_INDEX) b = &g->index;
    else if (type == RTC_BUFFER_TYPE_VERTEX_ATTRIBUTE) { if (slot >= g->attr.size()) g->attr.resize(slot + 1); b = &g->attr[slot]; }
    if (!b) return nullptr;
    b->bytes.assign(stride * count + 16, 0); b->stride = stride; b->count = count; b->format = format;   // Embree pads vertex buffers by 16 bytes
    return b->bytes.data();
}
RTC_API void rtcSetGeometryUserData(RTCGeometry g, void* p) { ((Geometry*)g)->user_data = p; }
RTC_API void* rtcGetGeometryUserData(RTCGeometry g) { return ((Geometry*)g)->user_data; }
RTC_API unsigned int rtcAttachGeometry(RTCScene sc, RTCGeometry gg) {
    Scene* s = (Scene*)sc; Geometry* g = (Geometry*)gg;
    g->refs++; g->scene = s; g->id = (unsigned)s->geoms.size();
    s->geoms.push_back(g);
    return g->id;
}
RTC_API RTCGeometry rtcGetGeometry(RTCScene sc, unsigned int id) { Scene* s = (Scene*)sc; return id < s->geoms.size() ? (RTCGeometry)s->geoms[id] : nullptr; }

// the oracle keeps un-indexed corners: flatten every triangle through its index buffer
RTC_API void rtcCommitScene(RTCScene sc) {
    Scene* s = (Scene*)sc;
    orc_destroy(s->orc); s->orc = orc_create();
    for (Geometry* g : s->geoms) {
        const size_t T = g->index.count;
        std::vector<float> pos(9 * T), nrm(9 * T, 0.0f), uv(6 * T, 0.0f);
        for (size_t t = 0; t < T; ++t) {
            const unsigned* idx = (const unsigned*)(g->index.bytes.data() + t * g->index.stride);
            for (int c = 0; c < 3; ++c) {
                const float* p = (const float*)(g->vertex.bytes.data() + (size_t)idx[c] * g->vertex.stride);
                memcpy(&pos[9 * t + 3 * c], p, 12);
                if (g->attr.size() > 0 && g->attr[0].count) memcpy(&nrm[9 * t + 3 * c], g->attr[0].bytes.data() + (size_t)idx[c] * g->attr[0].stride, 12);
                if (g->attr.size() > 1 && g->attr[1].count) memcpy(&uv[6 * t + 2 * c], g->attr[1].bytes.data() + (size_t)idx[c] * g->attr[1].stride, 8);
            }
        }
        uint32_t id = 0;
        orc_add_mesh(s->orc, pos.data(), nrm.data(), uv.data(), (uint32_t)T, (int32_t)g->id, &id);
    }
    orc_commit(s->orc);
}

RTC_API void rtcIntersect1(RTCScene sc, struct RTCIntersectContext*, struct RTCRayHit* rh) {
    Scene* s = (Scene*)sc;
    orc_intersect(s->orc, (orc_rayhit*)rh, 1, s->brute, 0);
}

RTC_API void rtcInterpolate(const struct RTCInterpolateArguments* a) {
    Geometry* g = (Geometry*)a->geometry;
    if (!g || !g->scene || !a->P || a->bufferType != RTC_BUFFER_TYPE_VERTEX_ATTRIBUTE) return;
    const uint32_t geom = g->id, prim = a->primID;
    float out[3] = {0, 0, 0};
    orc_interpolate(g->scene->orc, &geom, &prim, &a->u, &a->v, 1, a->bufferSlot == 0 ? 0 : 1, out);
    for (unsigned k = 0; k < a->valueCount && k < 3; ++k) a->P[k] = out[k];
}
