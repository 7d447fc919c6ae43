// surface.h -- Surface / Triangle / Vertex (pg1/surface.h, pg1/triangle.h, pg1/vertex.h).
// The reference stores an array of Triangle = 3 x Vertex = 3 x 14 floats (168 B / triangle) and copies it corner by
// corner into Embree buffers (pg1/raytracer.cpp:98-119).  Here a Surface owns the three un-indexed SoA arrays that
// copy would produce (9 + 9 + 6 floats / triangle), so LoadScene hands them to pgrt_add_mesh without a transform;
// Triangle / Vertex are views for callers that want the reference's accessors.
#pragma once
#include <string>
#include <vector>
#include "material.h"

struct Vertex {            // pg1/vertex.h:24-28 (colour and tangent are never read by the path)
    Vector3 position, normal;
    Coord2f texture_coords[1];
};

class Surface;
class Triangle {
public:
    Triangle(const Surface* s, int i) : surface_(s), index_(i) {}
    Vertex vertex(int i) const;
private:
    const Surface* surface_; int index_;
};

class Surface {
public:
    Surface() = default;
    Surface(const std::string& name, int n) : name_(name) { reserve(n); }
    void reserve(int n) { positions.reserve(9 * (size_t)n); normals.reserve(9 * (size_t)n); tex_coords.reserve(6 * (size_t)n); }
    void push_corner(const Vector3& p, const Vector3& n, const Coord2f& t) {
        positions.push_back(p.x); positions.push_back(p.y); positions.push_back(p.z);
        normals.push_back(n.x); normals.push_back(n.y); normals.push_back(n.z);
        tex_coords.push_back(t.u); tex_coords.push_back(t.v);
    }
    Triangle get_triangle(int i) const { return Triangle(this, i); }
    std::string get_name() const { return name_; }
    int no_triangles() const { return (int)(positions.size() / 9); }
    int no_vertices() const { return 3 * no_triangles(); }
    void set_material(Material* m) { material_ = m; }
    Material* get_material() const { return material_; }

    std::vector<float> positions, normals, tex_coords;   // 9, 9, 6 floats per triangle

private:
    std::string name_ = "unknown";
    Material* material_ = nullptr;
};

inline Vertex Triangle::vertex(int i) const {
    Vertex v;
    const size_t c = 3 * (size_t)index_ + (size_t)i;
    v.position = Vector3(surface_->positions[3 * c], surface_->positions[3 * c + 1], surface_->positions[3 * c + 2]);
    v.normal = Vector3(surface_->normals[3 * c], surface_->normals[3 * c + 1], surface_->normals[3 * c + 2]);
    v.texture_coords[0] = Coord2f{surface_->tex_coords[2 * c], surface_->tex_coords[2 * c + 1]};
    return v;
}
