// material.h -- Material (pg1/material.h:34-125, defaults pg1/material.cpp:9-26) and LightSource (pg1/LightSource.h:3-16).
#pragma once
#include <string>
#include "pg1_types.h"
#include "texture.h"

#define NO_TEXTURES 4
#define IOR_AIR 1.000293f     // pg1/material.h:15
#define IOR_WATER 1.33f
#define IOR_GLASS 1.5f

class Material {
public:
    Material() = default;
    void set_name(const char* name) { name_ = name; }
    std::string get_name() const { return name_; }
    void set_texture(int slot, Texture* texture) { textures_[slot] = texture; }      // borrowed: the loader's cache owns textures
    Texture* get_texture(int slot) const { return textures_[slot]; }

    Vector3 ambient{0.1f, 0.1f, 0.1f};
    Vector3 diffuse{0.4f, 0.4f, 0.4f};
    Vector3 specular{0.8f, 0.8f, 0.8f};
    Vector3 emission{0.0f, 0.0f, 0.0f};
    Vector3 refractivity{0.0f, 0.0f, 0.0f};   // Tf, parsed and unused by the path
    float reflectivity = 0.99f;
    float shininess = 1.0f;
    float ior = 1.5f;
    // `type` (MTL "shader N") is left uninitialised by the reference's default constructor (pg1/material.cpp:9-26);
    // 3 = Phong is what every non-dielectric material of the shipped scene says, so that is the defined default here.
    int type = 3;

    static const char kDiffuseMapSlot = 0, kSpecularMapSlot = 1, kNormalMapSlot = 2, kOpacityMapSlot = 3;

private:
    Texture* textures_[NO_TEXTURES] = {nullptr, nullptr, nullptr, nullptr};
    std::string name_ = "default";
};

class LightSource {
public:
    LightSource(Vector3 position, Vector3 ambient, Vector3 diffuse, Vector3 spectular)
        : position_(position), ambient_(ambient), diffuse_(diffuse), spectular_(spectular) {}
    Vector3 position_, ambient_, diffuse_, spectular_;
};
