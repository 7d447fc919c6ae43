// objloader.h -- LoadOBJ / LoadMTL (pg1/objloader.h:19-20, pg1/objloader.cpp:53-507), same signatures and the same
// observable behaviour.  The reference makes three strtok passes over two copies of the file on one thread; here the
// file is cut at line ends into one slice per hardware thread, floats and face indices are parsed in parallel, and only
// what depends on file order (groups, usemtl, the corner copy into SoA surfaces, surface.h) is replayed sequentially.
// PG1_LOADER_THREADS overrides the thread count.
#pragma once
#include <map>
#include <string>
#include <vector>
#include "surface.h"

// Textures created while loading are owned by this cache (the reference leaks them through Material::~Material's
// mismatched delete[], pg1/material.cpp:59); free with ReleaseTextureCache after the scene is uploaded or dropped.
typedef std::map<std::string, Texture*> TextureCache;
void ReleaseTextureCache(TextureCache& cache);

int LoadMTL(const char* file_name, const char* path, std::vector<Material*>& materials, TextureCache* cache = nullptr);
int LoadOBJ(const char* file_name, std::vector<Surface*>& surfaces, std::vector<Material*>& materials, const bool flip_yz = false,
            const Vector3 default_color = Vector3(0.5f, 0.5f, 0.5f), TextureCache* cache = nullptr);
