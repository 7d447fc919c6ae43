// image_io.h -- image decode / encode for the host surface.  Replaces what the reference gets from FreeImage 3.18.0
// (pg1/texture.cpp:15-50; the library is not vendored): baseline JPEG (every .jpg under the reference's data/ is
// SOF0, 4:2:0 or 4:4:4, one with restart intervals; 4:2:2 is decoded as well), 8-bit non-interlaced PNG (tutorial_2's data/test4.png), binary PPM and
// 24/32-bit BMP in; PPM, PFM and PNG out (the headless counterpart of the D3D11 presentation).
// Decoded images are returned the way Texture keeps them: top-down rows, B,G,R(,A) byte order, rows padded to 4 bytes
// (FreeImage_GetPitch).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct RawImage {
    int width = 0, height = 0, pitch = 0, bpp = 0;     // bpp in BYTES per pixel: 3 or 4
    std::vector<uint8_t> bytes;                        // top-down, BGR(A)
};

bool LoadImageFile(const char* file_name, RawImage& out, std::string* error = nullptr);
bool DecodeJpeg(const uint8_t* data, size_t size, RawImage& out, std::string* error = nullptr);

// rgba: width*height*4 floats, row 0 = top (the layout of tex_data_, pg1/simpleguidx11.cpp:108-114)
bool WritePPM(const char* file_name, const float* rgba, int width, int height);   // 8-bit: round(clamp(c,0,1)*255), NaN -> 0
bool WritePFM(const char* file_name, const float* rgba, int width, int height);   // float RGB, bottom-up as PFM requires
bool WritePNG(const char* file_name, const float* rgba, int width, int height);   // 8-bit RGB, same quantisation as WritePPM
bool WriteImageFile(const char* file_name, const float* rgba, int width, int height);   // by extension: .png, .pfm, else PPM
