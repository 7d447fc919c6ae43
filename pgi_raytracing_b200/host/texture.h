// texture.h -- Texture (pg1/texture.h): raw top-down B,G,R(,A) bytes exactly as the reference's constructor leaves them
// (pg1/texture.cpp:36-47: FreeImage_ConvertToRawBits, topdown = TRUE, pitch = 4-byte aligned rows).
// The reference decodes with FreeImage (absent); this one decodes baseline JPEG, binary PPM and BMP itself
// (image_io.h).  Texel FILTERING is not done here: Texture::get_texel (pg1/texture.cpp:77-130) runs on the GPU
// (csrc/shading.cuh tex_get_texel); the host object only owns the bytes that pgrt_set_texture uploads.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

class Texture {
public:
    Texture() = default;
    explicit Texture(const char* file_name);                       // pg1/texture.cpp:5-54
    Texture(const uint8_t* bgr, int width, int height, int scan_width, int pixel_size);
    bool valid() const { return !data_.empty(); }
    int width() const { return width_; }
    int height() const { return height_; }
    int scan_width() const { return scan_width_; }
    int pixel_size() const { return pixel_size_; }
    const uint8_t* data() const { return data_.data(); }

private:
    int width_ = 0, height_ = 0, scan_width_ = 0, pixel_size_ = 0;
    std::vector<uint8_t> data_;
};
