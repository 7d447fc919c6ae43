#include "texture.h"
#include <cstdio>
#include <cstring>
#include "image_io.h"

Texture::Texture(const char* file_name) {
    RawImage im; std::string err;
    if (!LoadImageFile(file_name, im, &err)) { printf("Texture '%s' not loaded: %s\n", file_name, err.c_str()); return; }
    width_ = im.width; height_ = im.height; scan_width_ = im.pitch; pixel_size_ = im.bpp;
    data_.swap(im.bytes);
}

Texture::Texture(const uint8_t* bgr, int width, int height, int scan_width, int pixel_size)
    : width_(width), height_(height), scan_width_(scan_width), pixel_size_(pixel_size), data_(bgr, bgr + (size_t)scan_width * height) {}
