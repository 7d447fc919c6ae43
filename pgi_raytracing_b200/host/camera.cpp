#include "camera.h"
#include <cfloat>

PinHoleCamera::PinHoleCamera(const int width, const int height, const float fov_y, const Vector3 view_from, const Vector3 view_at)
    : width_(width), height_(height), fov_y_(fov_y), view_from_(view_from), view_at_(view_at) {
    f_y_ = height_ / (2 * std::tan(fov_y_ / 2));
    Vector3 z_c = view_from_ - view_at_;
    Vector3 x_c = up_.CrossProduct(z_c);
    Vector3 y_c = z_c.CrossProduct(x_c);
    z_c.Normalize(); x_c.Normalize(); y_c.Normalize();
    M_c_w_ = Matrix3x3(x_c, y_c, z_c);
}

static RTCRay make_ray(const Vector3& o, const Vector3& d, float tnear) {
    RTCRay r;
    r.org_x = o.x; r.org_y = o.y; r.org_z = o.z; r.tnear = tnear;
    r.dir_x = d.x; r.dir_y = d.y; r.dir_z = d.z; r.time = 0.0f;
    r.tfar = FLT_MAX; r.mask = 0; r.id = 0; r.flags = 0;
    return r;
}

RTCRay PinHoleCamera::generate_ray(const float x_i, const float y_i) const {
    Vector3 d_c(x_i - (float)(width_ / 2), (float)(height_ / 2) - y_i, -f_y_);   // integer halves, no half-pixel offset (App. A-13)
    d_c.Normalize();
    return make_ray(view_from_, M_c_w_ * d_c, 0.001f);
}

RTCRay PinHoleCamera::generate_ray(const float x_i, const float y_i, const float focal_lenght, const float rand1, const float rand2) const {
    Vector3 d_c(x_i - (float)(width_ / 2), (float)(height_ / 2) - y_i, -f_y_);
    d_c.Normalize();
    Vector3 d_ws = M_c_w_ * d_c;
    d_ws.Normalize();
    const Vector3 focal_point(view_from_.x + d_ws.x * focal_lenght, view_from_.y + d_ws.y * focal_lenght, view_from_.z + d_ws.z * focal_lenght);
    const Vector3 shift = M_c_w_ * Vector3(rand1, rand2, 0.0f);
    const Vector3 org(view_from_.x + shift.x, view_from_.y + shift.y, view_from_.z + shift.z);
    Vector3 dir(focal_point.x - org.x, focal_point.y - org.y, focal_point.z - org.z);
    dir.Normalize();
    return make_ray(org, dir, 0.01f);
}
