// objloader.cpp -- see objloader.h.  Behaviour kept from the reference (pg1/objloader.cpp), line by line where it
// decides what the scene IS:
//   * any line whose first character is 'm' names a material library (":273-279"); libraries load before geometry;
//   * only "v ", "vn", "vt" records; normals are normalised on load (:323); "vt" keeps u, v;
//   * faces need all of v/vt/vn, 1-based, digits only (:458-466); three corners = triangle, four = quad split
//     (0,1,2) + (0,2,3) (:455-471); the corner count is the number of spaces in the trimmed line (:431-439);
//   * a surface is what lies between two 'g' lines, flushed at the next 'g' (or the end) with the LAST "usemtl" seen,
//     matched by exact name, first match (:382-400, :479-495); a 'g' with no faces since the last flush only renames;
//   * MTL keys match by PREFIX on the trimmed line, every key tested independently (:131-186); sscanf-style partial
//     parses keep the defaults of Material() for the fields not read ("Ks 1.0. 1.0 1.0" -> (1.0, 0.8, 0.8));
//   * a material is pushed when the next "newmtl" arrives unless one of that name exists; the last one always (:112-120,
//     :193-198).
// Defined where the reference is undefined: faces with another corner count are skipped; indices out of range are
// skipped; faces before the first 'g' go to a surface named "" ; a property line before any "newmtl" is ignored.
#include "objloader.h"
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <functional>
#include <chrono>
#include <thread>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

bool read_file(const char* file_name, std::string& out) {
    FILE* f = fopen(file_name, "rb");
    if (!f) { printf("File %s not found.\n", file_name); return false; }
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const size_t got = out.empty() ? 0 : fread(&out[0], 1, out.size(), f);
    fclose(f);
    out.resize(got);
    return true;
}

// The OBJ text: malloc'ed (not value-initialised: the reading threads touch its pages first, in parallel), NUL-terminated
struct TextBuf {
    char* p = nullptr; size_t n = 0;
    TextBuf() = default;
    TextBuf(const TextBuf&) = delete;
    TextBuf& operator=(const TextBuf&) = delete;
    ~TextBuf() { free(p); }
    const char* data() const { return p; }
    size_t size() const { return n; }
    char operator[](size_t i) const { return p[i]; }
};

// Every thread preads its own part of the file: page faults of the fresh buffer and the copy out of the page cache scale.
bool read_file_parallel(const char* file_name, TextBuf& out, unsigned n_threads) {
    const int fd = open(file_name, O_RDONLY);
    if (fd < 0) { printf("File %s not found.\n", file_name); return false; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 0) { close(fd); return false; }
    const size_t n = (size_t)st.st_size;
    out.p = (char*)malloc(n + 1);
    if (!out.p) { close(fd); return false; }
    out.n = n; out.p[n] = 0;
    if (n < (size_t)(8u << 20)) n_threads = 1;
    std::vector<std::thread> pool;
    std::vector<size_t> got_to(n_threads, 0);
    const size_t chunk = (n + n_threads - 1) / n_threads;
    auto work = [&](unsigned t) {
        size_t off = std::min(n, chunk * t);
        const size_t end = std::min(n, off + chunk);
        while (off < end) {
            const ssize_t got = pread(fd, out.p + off, end - off, (off_t)off);
            if (got <= 0) break;
            off += (size_t)got;
        }
        got_to[t] = off;
    };
    for (unsigned t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    close(fd);
    for (unsigned t = 0; t < n_threads; ++t)
        if (got_to[t] != std::min(n, chunk * t + chunk)) {          // the file shrank under us: keep what is contiguous
            out.n = got_to[t]; out.p[out.n] = 0;
            break;
        }
    return true;
}

// one line at a time, "\n"-separated, empty lines skipped (strtok semantics), '\r' left for the trims
struct Lines {
    const char* p; const char* end;
    explicit Lines(const std::string& s) : p(s.data()), end(s.data() + s.size()) {}
    bool next(const char*& b, const char*& e) {
        while (p < end && *p == '\n') ++p;
        if (p >= end) return false;
        b = p;
        while (p < end && *p != '\n') ++p;
        e = p;
        return true;
    }
};

inline void trim(const char*& b, const char*& e) {
    while (b < e && isspace((unsigned char)*b)) ++b;
    while (e > b && isspace((unsigned char)e[-1])) --e;
}

// "%*s %s": skip the first token, return the second
std::string second_token(const char* b, const char* e) {
    while (b < e && isspace((unsigned char)*b)) ++b;
    while (b < e && !isspace((unsigned char)*b)) ++b;
    while (b < e && isspace((unsigned char)*b)) ++b;
    const char* t = b;
    while (b < e && !isspace((unsigned char)*b)) ++b;
    return std::string(t, b);
}

// "%*s %f %f %f" with scanf's partial-assignment behaviour: stops at the first field that does not parse.
// Parses in place (no per-line allocation: the workers of LoadOBJ would serialise on the allocator): the buffer is
// NUL-terminated and a number never contains the '\n' that ends the line, so strtof cannot run past it.
// strtof for the common case, about 8x faster than glibc's: up to 19 significant digits with |decimal exponent| <= 22 give
// m * 10^E (or m / 10^-E) as ONE correctly rounded double operation (m <= 2^53 and 10^|E| are exact doubles); the cast to
// float then rounds a second time, which can only go wrong when that double sits exactly on the midpoint of two floats
// (low 29 mantissa bits == 0x10000000): those, sub-normal / overflowing values, hex floats, inf / nan and longer digit
// strings return false and go through strtof.  Result: bit-identical to strtof (tests/test_host_cpp.py fuzzes it).
inline bool fast_strtof(const char* s, const char* e, const char*& after, float& out) {
    static const double P10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const char* p = s;
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
    if (p >= e || !((*p >= '0' && *p <= '9') || *p == '.')) return false;
    if (*p == '0' && p + 1 < e && (p[1] == 'x' || p[1] == 'X')) return false;
    uint64_t m = 0; int nd = 0, dec = 0; bool any = false;
    for (; p < e && *p >= '0' && *p <= '9'; ++p) {
        any = true;
        if (nd >= 19) return false;
        m = m * 10 + (uint64_t)(*p - '0');
        if (m) ++nd;
    }
    if (p < e && *p == '.') {
        ++p;
        for (; p < e && *p >= '0' && *p <= '9'; ++p) {
            any = true;
            if (nd >= 19) return false;
            m = m * 10 + (uint64_t)(*p - '0');
            if (m) ++nd;
            --dec;
        }
    }
    if (!any) return false;
    if (p < e && (*p == 'e' || *p == 'E')) {
        const char* q = p + 1;
        bool eneg = false;
        if (q < e && (*q == '-' || *q == '+')) { eneg = *q == '-'; ++q; }
        if (q < e && *q >= '0' && *q <= '9') {
            int ex = 0;
            for (; q < e && *q >= '0' && *q <= '9'; ++q) if (ex < 100000) ex = ex * 10 + (*q - '0');
            dec += eneg ? -ex : ex;
            p = q;
        }                                   // "1e" / "1e+": strtof stops before the 'e'
    }
    if (m == 0) { out = neg ? -0.0f : 0.0f; after = p; return true; }
    if (m <= (1ull << 53) && dec >= -22 && dec <= 22) {
        double d = (double)m;
        d = dec < 0 ? d / P10[-dec] : d * P10[dec];
        if (!(d > 1.2e-37 && d < 3.0e38)) return false;
        uint64_t bits; memcpy(&bits, &d, sizeof bits);
        if ((bits & 0x1FFFFFFFull) == 0x10000000ull) return false;
        out = (float)(neg ? -d : d);
        after = p;
        return true;
    }
#if defined(__x86_64__) && defined(__GNUC__)
    // 16-19 digits (a float printed as a double's repr): the same argument one precision up.  The x87 long double has a 64-bit
    // significand, so any 19-digit m and 10^|E| up to 10^27 are exact and m * 10^E is one correctly rounded operation; the
    // cast to float is ambiguous only on a float midpoint (low 40 significand bits == 0x8000000000).
    if (dec >= -27 && dec <= 27) {
        static const long double P10L[28] = {1e0L, 1e1L, 1e2L, 1e3L, 1e4L, 1e5L, 1e6L, 1e7L, 1e8L, 1e9L, 1e10L, 1e11L, 1e12L, 1e13L, 1e14L, 1e15L, 1e16L,
                                             1e17L, 1e18L, 1e19L, 1e20L, 1e21L, 1e22L, 1e23L, 1e24L, 1e25L, 1e26L, 1e27L};
        long double d = (long double)m;
        d = dec < 0 ? d / P10L[-dec] : d * P10L[dec];
        if (!(d > 1.2e-37L && d < 3.0e38L)) return false;
        uint64_t sig; memcpy(&sig, &d, sizeof sig);          // little endian: the 64-bit significand comes first
        if ((sig & 0xFFFFFFFFFFull) == 0x8000000000ull) return false;
        out = (float)(neg ? -d : d);
        after = p;
        return true;
    }
#endif
    return false;
}

int scan_floats(const char* b, const char* e, float* dst[], int n) {
    const char* s = b;
    while (s < e && isspace((unsigned char)*s)) ++s;
    while (s < e && !isspace((unsigned char)*s)) ++s;
    int got = 0;
    for (; got < n; ++got) {
        while (s < e && isspace((unsigned char)*s)) ++s;
        if (s >= e) break;
        const char* fast_after = nullptr; float v;
        if (fast_strtof(s, e, fast_after, v)) { *dst[got] = v; s = fast_after; continue; }
        char* after = nullptr;
        v = strtof(s, &after);
        if (after == s || after > e) break;
        *dst[got] = v;
        s = after;
    }
    return got;
}

inline bool starts_with(const char* b, const char* e, const char* key) {
    const size_t n = strlen(key);
    return (size_t)(e - b) >= n && memcmp(b, key, n) == 0;
}

bool material_exists(const std::vector<Material*>& materials, const std::string& name) {
    for (const Material* m : materials) if (m->get_name() == name) return true;
    return false;
}

Texture* texture_proxy(const std::string& full_name, TextureCache& cache) {   // pg1/objloader.cpp:28-44
    auto it = cache.find(full_name);
    if (it != cache.end()) return it->second;
    Texture* t = new Texture(full_name.c_str());
    cache[full_name] = t;
    return t;
}

TextureCache g_default_cache;

// "%[0-9]/%[0-9]/%[0-9]" then atoi - 1
bool parse_corner(const char* b, const char* e, int idx[3]) {
    for (int k = 0; k < 3; ++k) {
        const char* d = b;
        long v = 0;
        while (d < e && *d >= '0' && *d <= '9') { v = v * 10 + (*d - '0'); if (v > 0x7fffffff) return false; ++d; }
        if (d == b) return false;
        idx[k] = (int)v - 1;
        b = d;
        if (k < 2) { if (b >= e || *b != '/') return false; ++b; }
    }
    return true;
}

}  // namespace

void ReleaseTextureCache(TextureCache& cache) {
    for (auto& kv : cache) delete kv.second;
    cache.clear();
}

int LoadMTL(const char* file_name, const char* path, std::vector<Material*>& materials, TextureCache* cache) {
    std::string text;
    if (!read_file(file_name, text)) return -1;
    TextureCache& tc = cache ? *cache : g_default_cache;
    Material* material = nullptr;
    std::string material_name;
    Lines lines(text);
    const char *b, *e;
    while (lines.next(b, e)) {
        if (*b == '#') continue;
        if (starts_with(b, e, "newmtl")) {
            if (material) {
                material->set_name(material_name.c_str());
                if (!material_exists(materials, material_name)) materials.push_back(material); else delete material;
            }
            material_name = second_token(b, e);
            material = new Material();
            continue;
        }
        trim(b, e);
        if (!material || b >= e) continue;
        float* v3a[3] = {&material->ambient.x, &material->ambient.y, &material->ambient.z};
        float* v3t[3] = {&material->refractivity.x, &material->refractivity.y, &material->refractivity.z};
        float* v3d[3] = {&material->diffuse.x, &material->diffuse.y, &material->diffuse.z};
        float* v3s[3] = {&material->specular.x, &material->specular.y, &material->specular.z};
        float* v3e[3] = {&material->emission.x, &material->emission.y, &material->emission.z};
        if (starts_with(b, e, "Ka")) scan_floats(b, e, v3a, 3);
        if (starts_with(b, e, "shader")) {                                  // "%*s %d": sign and digits only ("4.7" and "4e1" read as 4)
            const char* s = b;
            while (s < e && !isspace((unsigned char)*s)) ++s;
            while (s < e && isspace((unsigned char)*s)) ++s;
            bool neg = false;
            if (s < e && (*s == '-' || *s == '+')) { neg = *s == '-'; ++s; }
            if (s < e && *s >= '0' && *s <= '9') {
                long v = 0;
                for (; s < e && *s >= '0' && *s <= '9'; ++s) if (v < 0x7fffffff) v = v * 10 + (*s - '0');
                material->type = (int)(neg ? -v : v);
            }
        }
        if (starts_with(b, e, "Ni")) { float* p[1] = {&material->ior}; scan_floats(b, e, p, 1); }
        if (starts_with(b, e, "Tf")) scan_floats(b, e, v3t, 3);
        if (starts_with(b, e, "Kd")) scan_floats(b, e, v3d, 3);
        if (starts_with(b, e, "Ks")) scan_floats(b, e, v3s, 3);
        if (starts_with(b, e, "Ke")) scan_floats(b, e, v3e, 3);
        if (starts_with(b, e, "Ns")) { float* p[1] = {&material->shininess}; scan_floats(b, e, p, 1); }
        if (starts_with(b, e, "map_Kd")) material->set_texture(Material::kDiffuseMapSlot, texture_proxy(std::string(path) + second_token(b, e), tc));
        if (starts_with(b, e, "map_Ks")) material->set_texture(Material::kSpecularMapSlot, texture_proxy(std::string(path) + second_token(b, e), tc));
        if (starts_with(b, e, "map_D")) material->set_texture(Material::kOpacityMapSlot, texture_proxy(std::string(path) + second_token(b, e), tc));
        if (starts_with(b, e, "map_bump")) {   // "%*s %*s %f %s": map_bump -bm <f> <file>
            std::string rest(b, e);
            char name[256] = {0}; float bm = 0;
            if (sscanf(rest.c_str(), "%*s %*s %f %255s", &bm, name) == 2)
                material->set_texture(Material::kNormalMapSlot, texture_proxy(std::string(path) + name, tc));
        }
    }
    if (material) { material->set_name(material_name.c_str()); materials.push_back(material); }
    return 0;
}

namespace {

// What one worker extracts from its slice of the file (whole lines only).  Floats and face indices -- where the time
// goes -- are parsed here, in parallel; everything that depends on file order is replayed sequentially afterwards.
struct FaceRec { int idx[4][3]; int corners; };             // corners: 3, 4, or 0 = a line the loader skips
struct Event { char type; uint32_t a, b; };                 // 'g' / 'u' / 'm': text[a, b) is the line; 'f': a = index into faces
struct Slice {
    std::vector<Vector3> vertices, normals; std::vector<Coord2f> tex_coords;
    std::vector<FaceRec> faces; std::vector<Event> events;
};

void parse_slice(const TextBuf& text, size_t lo, size_t hi, bool flip_yz, Slice& out) {
    const char* p = text.data() + lo; const char* end = text.data() + hi;
    while (p < end) {
        while (p < end && *p == '\n') ++p;
        if (p >= end) break;
        const char* b = p;
        while (p < end && *p != '\n') ++p;
        const char* e = p;
        switch (*b) {
        case 'v':
            if (e - b >= 2) {
                Vector3 v; float* q[3] = {&v.x, &v.y, &v.z};
                if (b[1] == ' ' || b[1] == 'n') {
                    if (flip_yz) { q[1] = &v.z; q[2] = &v.y; }
                    scan_floats(b, e, q, 3);
                    if (flip_yz) v.y *= -1;
                    if (b[1] == 'n') { v.Normalize(); out.normals.push_back(v); } else out.vertices.push_back(v);
                } else if (b[1] == 't') {
                    Coord2f t{0.0f, 0.0f}; float w = 0; float* r[3] = {&t.u, &t.v, &w};
                    scan_floats(b, e, r, 3);
                    out.tex_coords.push_back(t);
                }
            }
            break;
        case 'g': case 'u': case 'm':
            out.events.push_back(Event{*b, (uint32_t)(b - text.data()), (uint32_t)(e - text.data())});
            break;
        case 'f': {
            FaceRec f; f.corners = 0;
            const char *tb = b, *te = e;
            trim(tb, te);
            int spaces = 0;
            for (const char* c = tb; c < te; ++c) spaces += (*c == ' ');
            if (spaces == 3 || spaces == 4) {
                bool ok = true;
                const char* c = tb;
                while (c < te && !isspace((unsigned char)*c)) ++c;      // the "f" token
                for (int k = 0; k < spaces && ok; ++k) {
                    while (c < te && isspace((unsigned char)*c)) ++c;
                    const char* t = c;
                    while (c < te && !isspace((unsigned char)*c)) ++c;
                    ok = parse_corner(t, c, f.idx[k]);
                }
                if (ok) f.corners = spaces;
            }
            out.events.push_back(Event{'f', (uint32_t)out.faces.size(), 0});
            out.faces.push_back(f);
        } break;
        default: break;
        }
    }
}

}  // namespace

int LoadOBJ(const char* file_name, std::vector<Surface*>& surfaces, std::vector<Material*>& materials, const bool flip_yz,
            const Vector3 /*default_color: vertex colours are never read by the path*/, TextureCache* cache) {
    const bool timing = getenv("PG1_LOADER_TIMING") != nullptr;
    auto now = []() { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t_start = now();
    unsigned n_threads = std::thread::hardware_concurrency();
    if (const char* e = getenv("PG1_LOADER_THREADS")) n_threads = (unsigned)atoi(e);
    n_threads = std::max(1u, std::min(n_threads, 64u));
    TextBuf text;
    if (!read_file_parallel(file_name, text, n_threads)) return -1;
    if (text.size() >= 0xFFFFFFFFull) { printf("File %s is larger than 4 GiB.\n", file_name); return -1; }
    const auto t_read = now();
    std::string path;
    if (const char* slash = strrchr(file_name, '/')) path.assign(file_name, slash - file_name + 1);

    // ---- parallel part: the file is cut at line ends into one slice per hardware thread
    if (text.size() < (1u << 20)) n_threads = 1;
    std::vector<size_t> cut(n_threads + 1, text.size());
    cut[0] = 0;
    for (unsigned t = 1; t < n_threads; ++t) {
        size_t c = std::max(cut[t - 1], text.size() / n_threads * t);
        while (c < text.size() && text[c] != '\n') ++c;
        cut[t] = c;
    }
    std::vector<Slice> slices(n_threads);
    {
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < n_threads; ++t) pool.emplace_back(parse_slice, std::cref(text), cut[t], cut[t + 1], flip_yz, std::ref(slices[t]));
        parse_slice(text, cut[0], cut[1], flip_yz, slices[0]);
        for (auto& th : pool) th.join();
    }

    const auto t_parse = now();
    // ---- sequential part, in file order.  Material libraries first (the reference's first pass)
    for (const Slice& sl : slices)
        for (const Event& ev : sl.events)
            if (ev.type == 'm') LoadMTL((path + second_token(text.data() + ev.a, text.data() + ev.b)).c_str(), path.c_str(), materials, cache);
    // all coordinates of the file (OBJ indices are absolute: a face may name any record, as in the reference's second pass)
    std::vector<Vector3> vertices, normals; std::vector<Coord2f> tex_coords;
    {
        size_t nv = 0, nn = 0, nt = 0;
        for (const Slice& sl : slices) { nv += sl.vertices.size(); nn += sl.normals.size(); nt += sl.tex_coords.size(); }
        vertices.reserve(nv); normals.reserve(nn); tex_coords.reserve(nt);
        for (const Slice& sl : slices) {
            vertices.insert(vertices.end(), sl.vertices.begin(), sl.vertices.end());
            normals.insert(normals.end(), sl.normals.begin(), sl.normals.end());
            tex_coords.insert(tex_coords.end(), sl.tex_coords.begin(), sl.tex_coords.end());
        }
    }
    printf("%zu vertices, %zu normals and %zu texture coords.\n", vertices.size(), normals.size(), tex_coords.size());
    const auto t_concat = now();

    // Surfaces.  The walk over the events is sequential (group / material names and the "last usemtl before the next g wins"
    // rule depend on file order) but only counts: every valid face gets its surface and its triangle offset there.  The
    // corners -- where the bytes are -- are then copied by all threads.
    struct Job { const FaceRec* face; uint32_t surface, first_triangle; };
    std::vector<Job> jobs;
    {
        size_t n_faces = 0;
        for (const Slice& sl : slices) n_faces += sl.faces.size();
        jobs.reserve(n_faces);
    }
    int no_surfaces = 0;
    const size_t first_new = surfaces.size();
    std::string group_name, material_name;
    uint32_t pending = 0;                        // triangles of the group being read
    auto flush = [&]() {
        if (pending == 0) return;
        Surface* sf = new Surface(group_name, 0);
        sf->positions.resize(9 * (size_t)pending); sf->normals.resize(9 * (size_t)pending); sf->tex_coords.resize(6 * (size_t)pending);
        for (Material* m : materials) if (m->get_name() == material_name) { sf->set_material(m); break; }
        surfaces.push_back(sf);
        ++no_surfaces;
        pending = 0;
    };
    for (const Slice& sl : slices)
        for (const Event& ev : sl.events) {
            if (ev.type == 'g') { flush(); group_name = second_token(text.data() + ev.a, text.data() + ev.b); }
            else if (ev.type == 'u') material_name = second_token(text.data() + ev.a, text.data() + ev.b);
            else if (ev.type == 'f') {
                const FaceRec& f = sl.faces[ev.a];
                bool ok = f.corners != 0;
                for (int k = 0; k < f.corners && ok; ++k)
                    ok = f.idx[k][0] >= 0 && (size_t)f.idx[k][0] < vertices.size() && f.idx[k][1] >= 0 && (size_t)f.idx[k][1] < tex_coords.size() &&
                         f.idx[k][2] >= 0 && (size_t)f.idx[k][2] < normals.size();
                if (!ok) continue;
                jobs.push_back(Job{&f, (uint32_t)(first_new + (size_t)no_surfaces), pending});
                pending += f.corners == 4 ? 2u : 1u;
            }
        }
    flush();
    {
        auto fill = [&](size_t lo, size_t hi) {
            const int order[6] = {0, 1, 2, 0, 2, 3};
            for (size_t j = lo; j < hi; ++j) {
                const Job& job = jobs[j];
                Surface* sf = surfaces[job.surface];
                const int n_corners = job.face->corners == 4 ? 6 : 3;
                for (int k = 0; k < n_corners; ++k) {
                    const int* i = job.face->idx[order[k]];
                    const size_t c = 3 * (size_t)job.first_triangle + (size_t)k;
                    const Vector3& pv = vertices[i[0]]; const Vector3& nv = normals[i[2]]; const Coord2f& tv = tex_coords[i[1]];
                    sf->positions[3 * c] = pv.x; sf->positions[3 * c + 1] = pv.y; sf->positions[3 * c + 2] = pv.z;
                    sf->normals[3 * c] = nv.x; sf->normals[3 * c + 1] = nv.y; sf->normals[3 * c + 2] = nv.z;
                    sf->tex_coords[2 * c] = tv.u; sf->tex_coords[2 * c + 1] = tv.v;
                }
            }
        };
        const unsigned n_fill = jobs.size() < (1u << 16) ? 1u : n_threads;
        std::vector<std::thread> pool;
        const size_t per = (jobs.size() + n_fill - 1) / n_fill;
        for (unsigned t = 1; t < n_fill; ++t) pool.emplace_back(fill, std::min(jobs.size(), per * t), std::min(jobs.size(), per * t + per));
        fill(0, std::min(jobs.size(), per));
        for (auto& th : pool) th.join();
    }
    printf("%d group(s), %zu material(s)\n", no_surfaces, materials.size());
    if (timing) printf("LoadOBJ: read %.1f ms, parse (%u threads) %.1f ms, materials + concatenate %.1f ms, surfaces %.1f ms\n", ms(t_start, t_read), n_threads,
                       ms(t_read, t_parse), ms(t_parse, t_concat), ms(t_concat, now()));
    return no_surfaces;
}
