// camera.h -- PinHoleCamera (pg1/PinHoleCamera.h:15-41) and SphericalMap (pg1/SphericalMap.h).  The device builds the
// same basis from the same five constructor arguments (pgrt_set_camera); this host copy serves callers that want the
// reference's accessors and the pinhole generate_ray (the "camera obscura" of the README).
#pragma once
#include <memory>
#include <string>
#include "pg1_types.h"
#include "texture.h"

class PinHoleCamera {
public:
    PinHoleCamera() {}
    PinHoleCamera(const int width, const int height, const float fov_y, const Vector3 view_from, const Vector3 view_at);   // PinHoleCamera.cpp:5-29
    RTCRay generate_ray(const float xi, const float yi) const;                                                              // :31-63
    // thin lens (:65-105).  The reference draws the lens sample from a clock-seeded mt19937 (:77); here it is an argument
    // (rand1, rand2 in [-aperture/2, aperture/2)), so the call is reproducible.  The render path draws them on the device.
    RTCRay generate_ray(const float x_i, const float y_i, const float focal_lenght, const float rand1, const float rand2) const;
    Vector3 get_origin() const { return view_from_; }
    Vector3 get_direction() const { return view_at_ - view_from_; }
    int width() const { return width_; }
    int height() const { return height_; }
    float fov_y() const { return fov_y_; }
    Vector3 view_from() const { return view_from_; }
    Vector3 view_at() const { return view_at_; }

private:
    int width_{640}, height_{480};
    float fov_y_{0.785f};
    Vector3 view_from_, view_at_;
    Vector3 up_{Vector3(0.0f, 0.0f, 1.0f)};
    float f_y_{1.0f};
    Matrix3x3 M_c_w_;
};

class SphericalMap {   // owns the background texture; get_texel (SphericalMap.cpp:17-29) runs on the device (env_get_texel)
public:
    SphericalMap() {}
    explicit SphericalMap(const std::string& file_name) : texture_(std::make_shared<Texture>(file_name.c_str())) {}
    const Texture* texture() const { return texture_.get(); }
private:
    std::shared_ptr<Texture> texture_;
};
