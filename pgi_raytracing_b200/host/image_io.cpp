// image_io.cpp -- see image_io.h.  The JPEG decoder follows the published baseline process (ITU T.81) with the
// arithmetic the IJG library made the de-facto standard, so that the bytes agree with what FreeImage (libjpeg inside)
// hands the reference: 13-bit fixed-point "slow integer" inverse DCT after Loeffler-Ligtenberg-Moschytz, triangle-filter
// ("fancy") 2x chroma up-sampling with the alternating 7/8 and 1/2 rounding biases, 16-bit fixed-point YCbCr -> RGB.
#include "image_io.h"
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

bool fail(std::string* e, const char* msg) { if (e) *e = msg; return false; }

bool read_all(const char* file_name, std::vector<uint8_t>& out) {
    FILE* f = fopen(file_name, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const size_t got = out.empty() ? 0 : fread(out.data(), 1, out.size(), f);
    fclose(f);
    out.resize(got);
    return true;
}

void alloc_bgr(RawImage& im, int w, int h, int bpp) {
    im.width = w; im.height = h; im.bpp = bpp; im.pitch = (w * bpp + 3) & ~3;
    im.bytes.assign((size_t)im.pitch * h, 0);
}

// ------------------------------------------------------------------------------------------------ JPEG
const int ZIGZAG[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                        35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
    uint8_t bits[17] = {0}; uint8_t vals[256] = {0};
    int mincode[17], maxcode[18], valptr[17];
    uint16_t look[512];      // 9-bit lookahead: (length << 8) | symbol, 0 = longer code
    bool present = false;
    bool build() {      // false: the code lengths do not describe a prefix code (more codes of a length than that length has room for)
        int code = 0, k = 0;
        for (int l = 1; l <= 16; ++l) {
            valptr[l] = k; mincode[l] = code;
            code += bits[l]; k += bits[l];
            if (code > (1 << l)) { present = false; return false; }
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        memset(look, 0, sizeof look);
        code = 0; k = 0;
        for (int l = 1; l <= 9; ++l) {
            for (int i = 0; i < bits[l]; ++i, ++k, ++code) {
                const int first = code << (9 - l);
                for (int f = 0; f < (1 << (9 - l)); ++f) look[first + f] = (uint16_t)((l << 8) | vals[k]);
            }
            code <<= 1;
        }
        present = true;
        return true;
    }
};

struct BitReader {
    const uint8_t* p; const uint8_t* end;
    uint32_t acc = 0; int n = 0; bool hit_marker = false;
    void fill() {
        while (n <= 24) {
            int b = 0;
            if (!hit_marker && p < end) {
                b = *p;
                if (b == 0xFF) {
                    if (p + 1 < end && p[1] == 0x00) p += 2;
                    else { hit_marker = true; b = 0; }
                } else ++p;
            }
            acc |= (uint32_t)b << (24 - n);
            n += 8;
        }
    }
    int peek(int k) { if (n < k) fill(); return (int)(acc >> (32 - k)); }
    void skip(int k) { acc <<= k; n -= k; }
    int get(int k) { if (k == 0) return 0; const int v = peek(k); skip(k); return v; }
    void reset() { acc = 0; n = 0; hit_marker = false; }
};

inline int decode_symbol(BitReader& br, const Huff& h) {
    const int look = br.peek(9);
    const uint16_t e = h.look[look];
    if (e) { br.skip(e >> 8); return e & 0xFF; }
    int code = br.peek(16);
    for (int l = 10; l <= 16; ++l) {
        const int c = code >> (16 - l);
        if (c <= h.maxcode[l] && h.maxcode[l] >= 0) { br.skip(l); return h.vals[h.valptr[l] + c - h.mincode[l]]; }
    }
    br.skip(16);
    return 0;
}

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

// slow-but-accurate integer inverse DCT (LL&M), CONST_BITS = 13, PASS1_BITS = 2; output range-limited to 0..255 after +128
inline int descale(long x, int n) { return (int)((x + (1L << (n - 1))) >> n); }
void idct_islow(const int* in /*dequantised, natural order*/, uint8_t* out, int stride) {
    const long F_0_298 = 2446, F_0_390 = 3196, F_0_541 = 4433, F_0_765 = 6270, F_0_899 = 7373, F_1_175 = 9633, F_1_501 = 12299, F_1_847 = 15137,
               F_1_961 = 16069, F_2_053 = 16819, F_2_562 = 20995, F_3_072 = 25172;
    long ws[64];
    for (int c = 0; c < 8; ++c) {
        const int* p = in + c;
        if (!p[8] && !p[16] && !p[24] && !p[32] && !p[40] && !p[48] && !p[56]) {
            const long dc = (long)p[0] << 2;
            for (int r = 0; r < 8; ++r) ws[r * 8 + c] = dc;
            continue;
        }
        long z2 = p[16], z3 = p[48];
        long z1 = (z2 + z3) * F_0_541;
        long tmp2 = z1 + z3 * (-F_1_847), tmp3 = z1 + z2 * F_0_765;
        z2 = p[0]; z3 = p[32];
        long tmp0 = (z2 + z3) << 13, tmp1 = (z2 - z3) << 13;
        const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = p[56]; tmp1 = p[40]; tmp2 = p[24]; tmp3 = p[8];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2; long z4 = tmp1 + tmp3;
        const long z5 = (z3 + z4) * F_1_175;
        tmp0 *= F_0_298; tmp1 *= F_2_053; tmp2 *= F_3_072; tmp3 *= F_1_501;
        z1 *= -F_0_899; z2 *= -F_2_562; z3 *= -F_1_961; z4 *= -F_0_390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        ws[0 * 8 + c] = descale(tmp10 + tmp3, 11); ws[7 * 8 + c] = descale(tmp10 - tmp3, 11);
        ws[1 * 8 + c] = descale(tmp11 + tmp2, 11); ws[6 * 8 + c] = descale(tmp11 - tmp2, 11);
        ws[2 * 8 + c] = descale(tmp12 + tmp1, 11); ws[5 * 8 + c] = descale(tmp12 - tmp1, 11);
        ws[3 * 8 + c] = descale(tmp13 + tmp0, 11); ws[4 * 8 + c] = descale(tmp13 - tmp0, 11);
    }
    auto clamp8 = [](long v) { v += 128; return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); };
    for (int r = 0; r < 8; ++r) {
        const long* p = ws + r * 8;
        uint8_t* o = out + (size_t)r * stride;
        long z2 = p[2], z3 = p[6];
        long z1 = (z2 + z3) * F_0_541;
        long tmp2 = z1 + z3 * (-F_1_847), tmp3 = z1 + z2 * F_0_765;
        long tmp0 = (p[0] + p[4]) << 13, tmp1 = (p[0] - p[4]) << 13;
        const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = p[7]; tmp1 = p[5]; tmp2 = p[3]; tmp3 = p[1];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2; long z4 = tmp1 + tmp3;
        const long z5 = (z3 + z4) * F_1_175;
        tmp0 *= F_0_298; tmp1 *= F_2_053; tmp2 *= F_3_072; tmp3 *= F_1_501;
        z1 *= -F_0_899; z2 *= -F_2_562; z3 *= -F_1_961; z4 *= -F_0_390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        o[0] = clamp8(descale(tmp10 + tmp3, 18)); o[7] = clamp8(descale(tmp10 - tmp3, 18));
        o[1] = clamp8(descale(tmp11 + tmp2, 18)); o[6] = clamp8(descale(tmp11 - tmp2, 18));
        o[2] = clamp8(descale(tmp12 + tmp1, 18)); o[5] = clamp8(descale(tmp12 - tmp1, 18));
        o[3] = clamp8(descale(tmp13 + tmp0, 18)); o[4] = clamp8(descale(tmp13 - tmp0, 18));
    }
}

struct Component { int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, pred = 0; int bw = 0, bh = 0; std::vector<uint8_t> plane; int stride = 0; };


// ------------------------------------------------------------------------------------------------ PNG
// Non-interlaced 8-bit PNG (grey, grey+alpha, RGB, RGBA, palette) -> top-down BGR(A): zlib inflate (RFC 1950/1951:
// stored, fixed and dynamic Huffman blocks) + the five scan-line filters of the PNG specification.  The reference
// reaches the same bytes through FreeImage (pg1/texture.cpp:15-50; tutorial_2 loads data/test4.png, 64x32 RGBA).
struct Inflater {
    const uint8_t* p; const uint8_t* end; uint32_t acc = 0; int n = 0; bool bad = false;
    int bit() { if (n == 0) { if (p >= end) { bad = true; return 0; } acc = *p++; n = 8; } const int b = acc & 1; acc >>= 1; --n; return b; }
    uint32_t bits(int k) { uint32_t v = 0; for (int i = 0; i < k; ++i) v |= (uint32_t)bit() << i; return v; }
};
struct HuffTable {
    uint16_t count[16] = {0}; uint16_t symbol[320] = {0};
    void build(const uint8_t* lengths, int n) {
        for (int i = 0; i < 16; ++i) count[i] = 0;
        for (int i = 0; i < n; ++i) count[lengths[i]]++;
        count[0] = 0;
        uint16_t offs[16]; offs[1] = 0;
        for (int i = 1; i < 15; ++i) offs[i + 1] = (uint16_t)(offs[i] + count[i]);
        for (int i = 0; i < n; ++i) if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
    }
    int decode(Inflater& in) const {
        int code = 0, first = 0, index = 0;
        for (int len = 1; len <= 15; ++len) {
            code |= in.bit();
            const int c = count[len];
            if (code - c < first) return symbol[index + (code - first)];
            index += c; first += c; first <<= 1; code <<= 1;
            if (in.bad) return -1;
        }
        return -1;
    }
};

bool inflate_zlib(const uint8_t* d, size_t size, std::vector<uint8_t>& out) {
    if (size < 6) return false;
    Inflater in; in.p = d + 2; in.end = d + size;      // 2-byte zlib header; the Adler-32 trailer is not checked
    static const uint16_t LBASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint16_t LEXT[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t DBASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint16_t DEXT[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    for (;;) {
        const int last = in.bit();
        const uint32_t type = in.bits(2);
        if (in.bad) return false;
        if (type == 0) {
            in.n = 0;
            if (in.p + 4 > in.end) return false;
            const uint32_t len = in.p[0] | (in.p[1] << 8); in.p += 4;
            if (in.p + len > in.end) return false;
            out.insert(out.end(), in.p, in.p + len); in.p += len;
        } else if (type == 1 || type == 2) {
            HuffTable lit, dist; uint8_t lengths[320];
            if (type == 1) {
                for (int i = 0; i < 144; ++i) lengths[i] = 8;
                for (int i = 144; i < 256; ++i) lengths[i] = 9;
                for (int i = 256; i < 280; ++i) lengths[i] = 7;
                for (int i = 280; i < 288; ++i) lengths[i] = 8;
                lit.build(lengths, 288);
                for (int i = 0; i < 30; ++i) lengths[i] = 5;
                dist.build(lengths, 30);
            } else {
                const int nlen = (int)in.bits(5) + 257, ndist = (int)in.bits(5) + 1, ncode = (int)in.bits(4) + 4;
                static const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                uint8_t cl[19] = {0};
                for (int i = 0; i < ncode; ++i) cl[ORDER[i]] = (uint8_t)in.bits(3);
                HuffTable clt; clt.build(cl, 19);
                int idx = 0;
                while (idx < nlen + ndist) {
                    const int sym = clt.decode(in);
                    if (sym < 0) return false;
                    if (sym < 16) lengths[idx++] = (uint8_t)sym;
                    else {
                        int rep, val = 0;
                        if (sym == 16) { if (idx == 0) return false; val = lengths[idx - 1]; rep = 3 + (int)in.bits(2); }
                        else if (sym == 17) rep = 3 + (int)in.bits(3);
                        else rep = 11 + (int)in.bits(7);
                        if (idx + rep > nlen + ndist) return false;
                        while (rep--) lengths[idx++] = (uint8_t)val;
                    }
                }
                lit.build(lengths, nlen); dist.build(lengths + nlen, ndist);
            }
            for (;;) {
                const int sym = lit.decode(in);
                if (sym < 0 || in.bad) return false;
                if (sym < 256) out.push_back((uint8_t)sym);
                else if (sym == 256) break;
                else {
                    if (sym > 285) return false;
                    const int len = LBASE[sym - 257] + (int)in.bits(LEXT[sym - 257]);
                    const int ds = dist.decode(in);
                    if (ds < 0 || ds > 29) return false;
                    const size_t dd = DBASE[ds] + in.bits(DEXT[ds]);
                    if (dd > out.size()) return false;
                    const size_t from = out.size() - dd;
                    for (int i = 0; i < len; ++i) out.push_back(out[from + i]);
                }
            }
        } else return false;
        if (last) break;
    }
    return !in.bad;
}

bool decode_png(const uint8_t* d, size_t size, RawImage& out, std::string* error) {
    static const uint8_t SIG[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (size < 33 || memcmp(d, SIG, 8) != 0) return fail(error, "not a PNG stream");
    auto be32 = [&](size_t o) { return ((uint32_t)d[o] << 24) | ((uint32_t)d[o + 1] << 16) | ((uint32_t)d[o + 2] << 8) | d[o + 3]; };
    int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    for (size_t i = 8; i + 12 <= size;) {
        const uint32_t len = be32(i);
        if (i + 12 + (size_t)len > size) return fail(error, "truncated PNG chunk");
        const uint8_t* c = d + i + 8;
        if (!memcmp(d + i + 4, "IHDR", 4)) { w = (int)be32(i + 8); h = (int)be32(i + 12); depth = c[8]; ctype = c[9]; interlace = c[12]; }
        else if (!memcmp(d + i + 4, "PLTE", 4)) plte.assign(c, c + len);
        else if (!memcmp(d + i + 4, "tRNS", 4)) trns.assign(c, c + len);
        else if (!memcmp(d + i + 4, "IDAT", 4)) idat.insert(idat.end(), c, c + len);
        else if (!memcmp(d + i + 4, "IEND", 4)) break;
        i += 12 + (size_t)len;
    }
    if (w <= 0 || h <= 0 || depth != 8 || interlace != 0) return fail(error, "unsupported PNG (8-bit non-interlaced only)");
    int ch;
    switch (ctype) { case 0: ch = 1; break; case 2: ch = 3; break; case 3: ch = 1; break; case 4: ch = 2; break; case 6: ch = 4; break; default: return fail(error, "unsupported PNG colour type"); }
    std::vector<uint8_t> raw;
    raw.reserve((size_t)(w * ch + 1) * h);
    if (!inflate_zlib(idat.data(), idat.size(), raw) || raw.size() < (size_t)(w * ch + 1) * h) return fail(error, "PNG inflate failed");
    const int stride = w * ch;
    std::vector<uint8_t> img((size_t)stride * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t f = raw[(size_t)y * (stride + 1)];
        const uint8_t* src = &raw[(size_t)y * (stride + 1) + 1];
        uint8_t* dst = &img[(size_t)y * stride];
        const uint8_t* up = y ? dst - stride : nullptr;
        for (int x = 0; x < stride; ++x) {
            const int a = x >= ch ? dst[x - ch] : 0, b = up ? up[x] : 0, c = (up && x >= ch) ? up[x - ch] : 0;
            int v = src[x];
            switch (f) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: { const int pa = abs(b - c), pb = abs(a - c), pc = abs(a + b - 2 * c); v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); } break;
                default: return fail(error, "bad PNG filter");
            }
            dst[x] = (uint8_t)v;
        }
    }
    const bool alpha = ctype == 4 || ctype == 6 || (ctype == 3 && !trns.empty());
    alloc_bgr(out, w, h, alpha ? 4 : 3);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const uint8_t* s = &img[(size_t)y * stride + (size_t)x * ch];
            uint8_t r, g, b, a = 255;
            if (ctype == 0) { r = g = b = s[0]; }
            else if (ctype == 4) { r = g = b = s[0]; a = s[1]; }
            else if (ctype == 3) {
                if ((size_t)s[0] * 3 + 2 >= plte.size()) return fail(error, "PNG palette index out of range");
                r = plte[s[0] * 3]; g = plte[s[0] * 3 + 1]; b = plte[s[0] * 3 + 2]; a = s[0] < trns.size() ? trns[s[0]] : 255;
            } else { r = s[0]; g = s[1]; b = s[2]; if (ctype == 6) a = s[3]; }
            uint8_t* o = &out.bytes[(size_t)y * out.pitch + (size_t)x * out.bpp];
            o[0] = b; o[1] = g; o[2] = r; if (alpha) o[3] = a;
        }
    return true;
}
}  // namespace

bool DecodeJpeg(const uint8_t* d, size_t size, RawImage& out, std::string* error) {
    if (size < 4 || d[0] != 0xFF || d[1] != 0xD8) return fail(error, "not a JPEG stream");
    uint16_t qt[4][64]; bool have_qt[4] = {false, false, false, false};
    Huff dc[4], ac[4];
    Component comp[3]; int ncomp = 0, width = 0, height = 0, restart = 0;
    size_t i = 2;
    bool sof = false;
    while (i + 4 <= size) {
        if (d[i] != 0xFF) { ++i; continue; }
        const int m = d[i + 1];
        if (m == 0xFF) { ++i; continue; }
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) { i += 2; continue; }
        if (m == 0xD9) break;
        const size_t len = ((size_t)d[i + 2] << 8) | d[i + 3];
        if (len < 2 || i + 2 + len > size) return fail(error, "truncated JPEG segment");
        const uint8_t* s = d + i + 4; const uint8_t* se = d + i + 2 + len;
        // every field below comes from the file: nothing is read past the end of its segment (se), no table or component
        // index is used before it has been range-checked
        if (m == 0xDB) {
            while (s < se) {
                const int pq = *s >> 4, tq = *s & 15; ++s;
                if (tq > 3 || pq > 1) return fail(error, "bad quantisation table id");
                if (s + (pq ? 128 : 64) > se) return fail(error, "truncated quantisation table");
                for (int k = 0; k < 64; ++k) { qt[tq][ZIGZAG[k]] = pq ? (uint16_t)((s[0] << 8) | s[1]) : s[0]; s += pq ? 2 : 1; }
                have_qt[tq] = true;
            }
        } else if (m == 0xC4) {
            while (s < se) {
                const int tc = *s >> 4, th = *s & 15; ++s;
                if (th > 3 || tc > 1) return fail(error, "bad Huffman table id");
                if (s + 16 > se) return fail(error, "truncated Huffman table");
                Huff& h = tc ? ac[th] : dc[th];
                int total = 0;
                for (int l = 1; l <= 16; ++l) { h.bits[l] = *s++; total += h.bits[l]; }
                if (total > 256 || s + total > se) return fail(error, "bad Huffman table");
                memcpy(h.vals, s, total); s += total;
                if (!h.build()) return fail(error, "bad Huffman table (over-subscribed code lengths)");
            }
        } else if (m == 0xC0 || m == 0xC1) {
            if (se - s < 6) return fail(error, "truncated frame header");
            if (s[0] != 8) return fail(error, "only 8-bit JPEG is supported");
            height = (s[1] << 8) | s[2]; width = (s[3] << 8) | s[4]; ncomp = s[5];
            if ((ncomp != 1 && ncomp != 3) || width <= 0 || height <= 0) return fail(error, "unsupported JPEG component count");
            if ((size_t)width * (size_t)height > ((size_t)1 << 28)) return fail(error, "JPEG frame too large");
            if (se - s < 6 + 3 * ncomp) return fail(error, "truncated frame header");
            for (int c = 0; c < ncomp; ++c) {
                comp[c].id = s[6 + 3 * c]; comp[c].h = s[7 + 3 * c] >> 4; comp[c].v = s[7 + 3 * c] & 15; comp[c].tq = s[8 + 3 * c];
                comp[c].td = comp[c].ta = -1;
                if (comp[c].h < 1 || comp[c].h > 4 || comp[c].v < 1 || comp[c].v > 4 || comp[c].tq > 3) return fail(error, "bad JPEG component parameters");
            }
            sof = true;
        } else if (m == 0xC2 || (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC)) {
            return fail(error, "only baseline (SOF0) JPEG is supported");
        } else if (m == 0xDD) {
            if (se - s < 2) return fail(error, "truncated restart interval");
            restart = (s[0] << 8) | s[1];
        } else if (m == 0xDA) {
            if (!sof) return fail(error, "SOS before SOF");
            if (se - s < 1) return fail(error, "truncated scan header");
            const int ns = s[0];
            if (ns != ncomp) return fail(error, "non-interleaved scans are not supported");
            if (se - s < 1 + 2 * ns + 3) return fail(error, "truncated scan header");
            for (int k = 0; k < ns; ++k) {
                const int id = s[1 + 2 * k];
                for (int c = 0; c < ncomp; ++c) if (comp[c].id == id) { comp[c].td = s[2 + 2 * k] >> 4; comp[c].ta = s[2 + 2 * k] & 15; }
            }
            for (int c = 0; c < ncomp; ++c)
                if (comp[c].td < 0 || comp[c].td > 3 || comp[c].ta < 0 || comp[c].ta > 3) return fail(error, "bad Huffman table selector in the scan header");
            // ---- entropy-coded data
            int hmax = 1, vmax = 1;
            for (int c = 0; c < ncomp; ++c) { hmax = comp[c].h > hmax ? comp[c].h : hmax; vmax = comp[c].v > vmax ? comp[c].v : vmax; }
            const int mcux = (width + 8 * hmax - 1) / (8 * hmax), mcuy = (height + 8 * vmax - 1) / (8 * vmax);
            for (int c = 0; c < ncomp; ++c) {
                if (!have_qt[comp[c].tq] || !dc[comp[c].td].present || !ac[comp[c].ta].present) return fail(error, "missing JPEG table");
                comp[c].bw = mcux * comp[c].h; comp[c].bh = mcuy * comp[c].v;
                comp[c].stride = comp[c].bw * 8;
                comp[c].plane.assign((size_t)comp[c].stride * comp[c].bh * 8, 0);
                comp[c].pred = 0;
            }
            BitReader br; br.p = se; br.end = d + size;
            int coef[64], to_restart = restart;
            for (int my = 0; my < mcuy; ++my)
                for (int mx = 0; mx < mcux; ++mx) {
                    if (restart && to_restart == 0) {
                        // byte-align, expect RSTn
                        br.reset();
                        while (br.p + 1 < br.end && !(br.p[0] == 0xFF && br.p[1] >= 0xD0 && br.p[1] <= 0xD7)) ++br.p;
                        if (br.p + 1 < br.end) br.p += 2;
                        for (int c = 0; c < ncomp; ++c) comp[c].pred = 0;
                        to_restart = restart;
                    }
                    for (int c = 0; c < ncomp; ++c) {
                        Component& C = comp[c];
                        for (int by = 0; by < C.v; ++by)
                            for (int bx = 0; bx < C.h; ++bx) {
                                memset(coef, 0, sizeof coef);
                                const int t = decode_symbol(br, dc[C.td]);
                                if (t > 11) return fail(error, "bad DC difference category");      // 8-bit baseline: 0..11; a crafted table could name up to 255
                                const int diff = t ? extend(br.get(t), t) : 0;
                                C.pred += diff;
                                coef[0] = C.pred * qt[C.tq][0];
                                for (int k = 1; k < 64;) {
                                    const int rs = decode_symbol(br, ac[C.ta]);
                                    const int r = rs >> 4, sz = rs & 15;
                                    if (sz == 0) { if (r == 15) { k += 16; continue; } break; }
                                    k += r;
                                    if (k > 63) break;
                                    const int nat = ZIGZAG[k];
                                    coef[nat] = extend(br.get(sz), sz) * qt[C.tq][nat];
                                    ++k;
                                }
                                const int X = (mx * C.h + bx) * 8, Y = (my * C.v + by) * 8;
                                idct_islow(coef, &C.plane[(size_t)Y * C.stride + X], C.stride);
                            }
                    }
                    if (restart) --to_restart;
                }
            // ---- up-sampling + colour conversion
            alloc_bgr(out, width, height, 3);
            if (ncomp == 1) {
                for (int y = 0; y < height; ++y)
                    for (int x = 0; x < width; ++x) {
                        const uint8_t v = comp[0].plane[(size_t)y * comp[0].stride + x];
                        uint8_t* o = &out.bytes[(size_t)y * out.pitch + 3 * x]; o[0] = o[1] = o[2] = v;
                    }
                return true;
            }
            const bool h2v2 = comp[0].h == 2 && comp[0].v == 2 && comp[1].h == 1 && comp[1].v == 1 && comp[2].h == 1 && comp[2].v == 1;
            const bool h1v1 = comp[0].h == 1 && comp[0].v == 1 && comp[1].h == 1 && comp[1].v == 1 && comp[2].h == 1 && comp[2].v == 1;
            const bool h2v1 = comp[0].h == 2 && comp[0].v == 1 && comp[1].h == 1 && comp[1].v == 1 && comp[2].h == 1 && comp[2].v == 1;
            if (!h2v2 && !h1v1 && !h2v1) return fail(error, "unsupported chroma sub-sampling (only 4:4:4, 4:2:2 and 4:2:0)");
            std::vector<uint8_t> up[2];
            const int cw = (width + 1) / 2, chh = (height + 1) / 2;     // down-sampled size actually carrying image data
            // libjpeg picks the triangle ("fancy") filters only when the down-sampled row has more than two samples
            // (jdsample.c, jinit_upsampler); narrower images get plain replication
            const bool fancy = cw > 2;
            if (h2v2 || h2v1) {
                for (int c = 0; c < 2; ++c) {
                    const Component& C = comp[1 + c];
                    up[c].assign((size_t)width * height + 2 * width + 4, 0);
                    std::vector<int> colsum(cw);
                    for (int y = 0; y < height && (!fancy || h2v1); ++y) {
                        const uint8_t* a = &C.plane[(size_t)(h2v1 ? y : y >> 1) * C.stride];
                        uint8_t* o = &up[c][(size_t)y * width];
                        for (int x = 0; x < cw; ++x) {
                            int e0 = a[x], e1 = a[x];                                       // replication
                            if (fancy) {                                                      // h2v1_fancy_upsample
                                e0 = x == 0 ? a[x] : (3 * a[x] + a[x - 1] + 1) >> 2;
                                e1 = x + 1 == cw ? a[x] : (3 * a[x] + a[x + 1] + 2) >> 2;
                            }
                            if (2 * x < width) o[2 * x] = (uint8_t)e0;
                            if (2 * x + 1 < width) o[2 * x + 1] = (uint8_t)e1;
                        }
                    }
                    for (int y = 0; y < height && fancy && h2v2; ++y) {
                        const int r = y >> 1;
                        int rn = (y & 1) ? r + 1 : r - 1;                 // nearer neighbour row; edges replicate
                        rn = rn < 0 ? 0 : (rn >= chh ? chh - 1 : rn);
                        const uint8_t* a = &C.plane[(size_t)r * C.stride]; const uint8_t* b = &C.plane[(size_t)rn * C.stride];
                        for (int x = 0; x < cw; ++x) colsum[x] = 3 * a[x] + b[x];
                        uint8_t* o = &up[c][(size_t)y * width];
                        for (int x = 0; x < cw; ++x) {
                            const int cur = colsum[x], last = x > 0 ? colsum[x - 1] : 0, next = x + 1 < cw ? colsum[x + 1] : 0;
                            const int e0 = x == 0 ? (cur * 4 + 8) >> 4 : (cur * 3 + last + 8) >> 4;
                            const int e1 = x + 1 == cw ? (cur * 4 + 7) >> 4 : (cur * 3 + next + 7) >> 4;
                            if (2 * x < width) o[2 * x] = (uint8_t)e0;
                            if (2 * x + 1 < width) o[2 * x + 1] = (uint8_t)e1;
                        }
                    }
                }
            }
            for (int y = 0; y < height; ++y) {
                const uint8_t* Y = &comp[0].plane[(size_t)y * comp[0].stride];
                const uint8_t* Cb = !h1v1 ? &up[0][(size_t)y * width] : &comp[1].plane[(size_t)y * comp[1].stride];
                const uint8_t* Cr = !h1v1 ? &up[1][(size_t)y * width] : &comp[2].plane[(size_t)y * comp[2].stride];
                uint8_t* o = &out.bytes[(size_t)y * out.pitch];
                for (int x = 0; x < width; ++x) {
                    const int yy = Y[x], cb = Cb[x] - 128, cr = Cr[x] - 128;
                    int r = yy + (int)((91881L * cr + 32768) >> 16);
                    int g = yy + (int)((-22554L * cb - 46802L * cr + 32768) >> 16);
                    int b = yy + (int)((116130L * cb + 32768) >> 16);
                    r = r < 0 ? 0 : (r > 255 ? 255 : r); g = g < 0 ? 0 : (g > 255 ? 255 : g); b = b < 0 ? 0 : (b > 255 ? 255 : b);
                    o[3 * x] = (uint8_t)b; o[3 * x + 1] = (uint8_t)g; o[3 * x + 2] = (uint8_t)r;
                }
            }
            return true;
        }
        i += 2 + len;
    }
    return fail(error, "no scan found in JPEG stream");
}

bool LoadImageFile(const char* file_name, RawImage& out, std::string* error) {
    std::vector<uint8_t> d;
    if (!read_all(file_name, d)) return fail(error, "cannot open image file");
    if (d.size() >= 2 && d[0] == 0xFF && d[1] == 0xD8) return DecodeJpeg(d.data(), d.size(), out, error);
    if (d.size() >= 8 && d[0] == 0x89 && d[1] == 'P' && d[2] == 'N' && d[3] == 'G') return decode_png(d.data(), d.size(), out, error);
    if (d.size() >= 2 && d[0] == 'P' && d[1] == '6') {          // binary PPM, maxval 255
        size_t p = 2; int vals[3], got = 0;
        while (got < 3 && p < d.size()) {
            while (p < d.size() && (isspace(d[p]) || d[p] == '#')) { if (d[p] == '#') while (p < d.size() && d[p] != '\n') ++p; else ++p; }
            int v = 0; bool any = false;
            while (p < d.size() && d[p] >= '0' && d[p] <= '9') { v = v * 10 + (d[p] - '0'); ++p; any = true; }
            if (!any) break;
            vals[got++] = v;
        }
        if (got != 3 || vals[2] != 255 || p >= d.size()) return fail(error, "unsupported PPM");
        ++p;
        if (d.size() - p < (size_t)vals[0] * vals[1] * 3) return fail(error, "truncated PPM");
        alloc_bgr(out, vals[0], vals[1], 3);
        for (int y = 0; y < vals[1]; ++y)
            for (int x = 0; x < vals[0]; ++x) {
                const uint8_t* s = &d[p + ((size_t)y * vals[0] + x) * 3]; uint8_t* o = &out.bytes[(size_t)y * out.pitch + 3 * x];
                o[0] = s[2]; o[1] = s[1]; o[2] = s[0];
            }
        return true;
    }
    if (d.size() >= 54 && d[0] == 'B' && d[1] == 'M') {          // uncompressed 24 / 32-bit BMP
        auto u32 = [&](size_t o) { return (uint32_t)d[o] | ((uint32_t)d[o + 1] << 8) | ((uint32_t)d[o + 2] << 16) | ((uint32_t)d[o + 3] << 24); };
        const uint32_t off = u32(10); const int w = (int)u32(18); int h = (int)u32(22); const int bits = d[28] | (d[29] << 8); const uint32_t compr = u32(30);
        const bool top_down = h < 0; if (top_down) h = -h;
        if ((bits != 24 && bits != 32) || compr != 0 || w <= 0 || h <= 0) return fail(error, "unsupported BMP");
        const int bpp = bits / 8, src_pitch = (w * bpp + 3) & ~3;
        if (d.size() < off + (size_t)src_pitch * h) return fail(error, "truncated BMP");
        alloc_bgr(out, w, h, bpp);
        for (int y = 0; y < h; ++y) memcpy(&out.bytes[(size_t)y * out.pitch], &d[off + (size_t)(top_down ? y : h - 1 - y) * src_pitch], (size_t)w * bpp);
        return true;
    }
    return fail(error, "unknown image format (JPEG, PNG, PPM P6 and BMP are supported)");
}

static inline uint8_t to8(float c) { if (!(c == c)) return 0; c = c < 0.f ? 0.f : (c > 1.f ? 1.f : c); return (uint8_t)std::floor(c * 255.0f + 0.5f); }

bool WritePPM(const char* file_name, const float* rgba, int width, int height) {
    FILE* f = fopen(file_name, "wb");
    if (!f) return false;
    fprintf(f, "P6\n%d %d\n255\n", width, height);
    std::vector<uint8_t> row((size_t)width * 3);
    for (int y = 0; y < height; ++y) {
        for (int x = 0; x < width; ++x) for (int c = 0; c < 3; ++c) row[3 * x + c] = to8(rgba[((size_t)y * width + x) * 4 + c]);
        fwrite(row.data(), 1, row.size(), f);
    }
    fclose(f);
    return true;
}

bool WritePFM(const char* file_name, const float* rgba, int width, int height) {
    FILE* f = fopen(file_name, "wb");
    if (!f) return false;
    fprintf(f, "PF\n%d %d\n-1.0\n", width, height);
    std::vector<float> row((size_t)width * 3);
    for (int y = height - 1; y >= 0; --y) {
        for (int x = 0; x < width; ++x) for (int c = 0; c < 3; ++c) row[3 * x + c] = rgba[((size_t)y * width + x) * 4 + c];
        fwrite(row.data(), sizeof(float), row.size(), f);
    }
    fclose(f);
    return true;
}

// ---- PNG out (headless presentation): 8-bit RGB, filter 0, zlib "stored" blocks (valid, uncompressed)
static uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
    static uint32_t table[256]; static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
        ready = true;
    }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFFu] ^ (crc >> 8);
    return crc;
}
static void put_be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(uint8_t(x >> 24)); v.push_back(uint8_t(x >> 16)); v.push_back(uint8_t(x >> 8)); v.push_back(uint8_t(x)); }
static bool write_chunk(FILE* f, const char type[4], const std::vector<uint8_t>& data) {
    std::vector<uint8_t> head; put_be32(head, (uint32_t)data.size());
    uint32_t crc = crc32_update(0xFFFFFFFFu, (const uint8_t*)type, 4);
    crc = crc32_update(crc, data.data(), data.size()) ^ 0xFFFFFFFFu;
    std::vector<uint8_t> tail; put_be32(tail, crc);
    return fwrite(head.data(), 1, 4, f) == 4 && fwrite(type, 1, 4, f) == 4 && fwrite(data.data(), 1, data.size(), f) == data.size() && fwrite(tail.data(), 1, 4, f) == 4;
}

bool WritePNG(const char* file_name, const float* rgba, int width, int height) {
    if (width <= 0 || height <= 0) return false;
    FILE* f = fopen(file_name, "wb");
    if (!f) return false;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    bool ok = fwrite(sig, 1, 8, f) == 8;
    std::vector<uint8_t> ihdr; put_be32(ihdr, (uint32_t)width); put_be32(ihdr, (uint32_t)height);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 8 bit, RGB, deflate, adaptive filters, no interlace
    ok = ok && write_chunk(f, "IHDR", ihdr);
    const size_t row_bytes = 1 + (size_t)width * 3;
    std::vector<uint8_t> raw(row_bytes * (size_t)height);
    for (int y = 0; y < height; ++y) {
        uint8_t* row = &raw[row_bytes * (size_t)y];
        row[0] = 0;                                                                                // filter type None
        for (int x = 0; x < width; ++x) for (int c = 0; c < 3; ++c) row[1 + 3 * x + c] = to8(rgba[((size_t)y * width + x) * 4 + c]);
    }
    uint32_t a = 1, b = 0;                                                                         // Adler-32 of the raw stream
    for (size_t i = 0; i < raw.size();) {
        const size_t n = std::min<size_t>(5552, raw.size() - i);
        for (size_t k = 0; k < n; ++k) { a += raw[i + k]; b += a; }
        a %= 65521u; b %= 65521u; i += n;
    }
    std::vector<uint8_t> z; z.reserve(raw.size() + raw.size() / 65535 * 5 + 16);
    z.push_back(0x78); z.push_back(0x01);
    for (size_t i = 0; i < raw.size();) {
        const size_t n = std::min<size_t>(65535, raw.size() - i);
        z.push_back(i + n == raw.size() ? 1 : 0);                                                  // BFINAL, BTYPE = 00 (stored)
        z.push_back(uint8_t(n)); z.push_back(uint8_t(n >> 8)); z.push_back(uint8_t(~n)); z.push_back(uint8_t((~n) >> 8));
        z.insert(z.end(), raw.begin() + (ptrdiff_t)i, raw.begin() + (ptrdiff_t)(i + n));
        i += n;
    }
    put_be32(z, (b << 16) | a);
    ok = ok && write_chunk(f, "IDAT", z) && write_chunk(f, "IEND", {});
    fclose(f);
    return ok;
}

bool WriteImageFile(const char* file_name, const float* rgba, int width, int height) {
    const std::string n(file_name);
    auto ends = [&](const char* e) { const size_t l = strlen(e); return n.size() >= l && strcasecmp(n.c_str() + n.size() - l, e) == 0; };
    if (ends(".png")) return WritePNG(file_name, rgba, width, height);
    if (ends(".pfm")) return WritePFM(file_name, rgba, width, height);
    return WritePPM(file_name, rgba, width, height);
}
