// ref_freeimage.cpp -- TEST INFRASTRUCTURE ONLY.  The FreeImage calls of Texture::Texture (pg1/texture.cpp:5-54), over raw
// side-car files: "<image>.bgr" = 16-byte header {int32 width, height, pitch, bytes per pixel} + pitch*height bytes,
// TOP-DOWN B,G,R(,A) rows -- what FreeImage_ConvertToRawBits(..., topdown = TRUE) leaves in Texture::data_ (texture.cpp:46-47).
// The tests write the side-car from a PIL decode (libjpeg-turbo / libpng), FreeImage's pitch (rows padded to 4 bytes).
// FreeImage's binary is not vendored by the reference (pg1_embree.vcxproj:79,87,157).
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <FreeImage.h>

namespace { struct Image { int32_t w = 0, h = 0, pitch = 0, bpp = 0; std::vector<uint8_t> bytes; }; }

static bool sidecar_exists(const char* f) { FILE* fp = fopen((std::string(f) + ".bgr").c_str(), "rb"); if (fp) fclose(fp); return fp != nullptr; }
DLL_API FREE_IMAGE_FORMAT DLL_CALLCONV FreeImage_GetFileType(const char* f, int) { return sidecar_exists(f) ? FIF_RAW : FIF_UNKNOWN; }
DLL_API FREE_IMAGE_FORMAT DLL_CALLCONV FreeImage_GetFIFFromFilename(const char* f) { return sidecar_exists(f) ? FIF_RAW : FIF_UNKNOWN; }
DLL_API BOOL DLL_CALLCONV FreeImage_FIFSupportsReading(FREE_IMAGE_FORMAT fif) { return fif == FIF_RAW; }
DLL_API FIBITMAP* DLL_CALLCONV FreeImage_Load(FREE_IMAGE_FORMAT, const char* f, int) {
    FILE* fp = fopen((std::string(f) + ".bgr").c_str(), "rb");
    if (!fp) return nullptr;
    Image* im = new Image();
    int32_t hdr[4];
    bool ok = fread(hdr, 4, 4, fp) == 4;
    if (ok) { im->w = hdr[0]; im->h = hdr[1]; im->pitch = hdr[2]; im->bpp = hdr[3]; ok = im->w > 0 && im->h > 0 && im->pitch >= im->w * im->bpp; }
    if (ok) { im->bytes.resize((size_t)im->pitch * im->h); ok = fread(im->bytes.data(), 1, im->bytes.size(), fp) == im->bytes.size(); }
    fclose(fp);
    if (!ok) { delete im; return nullptr; }
    FIBITMAP* dib = new FIBITMAP(); dib->data = im;
    return dib;
}
DLL_API void DLL_CALLCONV FreeImage_Unload(FIBITMAP* dib) { if (dib) { delete (Image*)dib->data; delete dib; } }
DLL_API BYTE* DLL_CALLCONV FreeImage_GetBits(FIBITMAP* dib) { return ((Image*)dib->data)->bytes.data(); }
DLL_API unsigned DLL_CALLCONV FreeImage_GetBPP(FIBITMAP* dib) { return 8u * (unsigned)((Image*)dib->data)->bpp; }
DLL_API unsigned DLL_CALLCONV FreeImage_GetWidth(FIBITMAP* dib) { return (unsigned)((Image*)dib->data)->w; }
DLL_API unsigned DLL_CALLCONV FreeImage_GetHeight(FIBITMAP* dib) { return (unsigned)((Image*)dib->data)->h; }
DLL_API unsigned DLL_CALLCONV FreeImage_GetPitch(FIBITMAP* dib) { return (unsigned)((Image*)dib->data)->pitch; }
DLL_API void DLL_CALLCONV FreeImage_ConvertToRawBits(BYTE* bits, FIBITMAP* dib, int pitch, unsigned, unsigned, unsigned, unsigned, BOOL) {
    const Image* im = (Image*)dib->data;       // the side-car already holds the top-down rows the call would produce
    for (int y = 0; y < im->h; ++y) memcpy(bits + (size_t)y * pitch, im->bytes.data() + (size_t)y * im->pitch, (size_t)(pitch < im->pitch ? pitch : im->pitch));
}
