"""ctypes front-end of ``oracle/_ref/libpg_ref.so``: the reference's OWN ``pg1/*.cpp`` sources, compiled unmodified by
``oracle/ref.mk`` (stubs only for what the reference does not vendor: the Embree and FreeImage binaries and the window).

TEST INFRASTRUCTURE ONLY, like everything under ``oracle/``.  The library is built in the development container (where
``/root/reference`` exists) and travels to the GPU box as a prebuilt file; ``available()`` says whether it is there.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libpg_ref.so")
_REF_SRC = "/root/reference/src/pg/pg1_embree"
_lib_cache = None


def build() -> str | None:
    """(Re)build when the reference sources are present; otherwise keep whatever prebuilt library exists."""
    if os.path.isdir(_REF_SRC):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libpg_oracle.so"])
        subprocess.check_call(["make", "-C", _HERE, "-s", "-f", "ref.mk"], stdout=subprocess.DEVNULL)
    return _LIB_PATH if os.path.exists(_LIB_PATH) else None


def available() -> bool:
    return build() is not None


def _lib():
    global _lib_cache
    if _lib_cache is None:
        if build() is None:
            raise RuntimeError("oracle/_ref/libpg_ref.so is missing and /root/reference is not here to build it from")
        lib = C.CDLL(_LIB_PATH)
        for name in ("ref_create", "ref_texture_load", "ref_envmap_load", "ref_obj_load"):
            getattr(lib, name).restype = C.c_void_p
        lib.ref_deg2rad.restype = C.c_float; lib.ref_deg2rad.argtypes = [C.c_float]
        _lib_cache = lib
    return _lib_cache


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def write_sidecar(image, path: str) -> None:
    """``<path>.bgr``: what the FreeImage stub loads in place of decoding ``path`` (oracle/ref_shim/ref_freeimage.cpp)."""
    with open(path + ".bgr", "wb") as f:
        f.write(np.array([image.width, image.height, image.pitch, image.bpp], np.int32).tobytes())
        f.write(np.ascontiguousarray(image.data).tobytes())


def write_scene(scene, directory: str):
    """The scene as the files ``Raytracer::LoadScene`` reads: OBJ + MTL (``scenes.write_obj``) with ``map_Kd`` lines for the
    textured materials, one raw side-car per image.  Returns (obj path, background path)."""
    import copy
    from pgi_raytracing_b200 import scenes
    os.makedirs(directory, exist_ok=True)
    sc = copy.copy(scene)
    sc.materials = [copy.copy(m) for m in scene.materials]
    for m in sc.materials:
        if m.diffuse_tex >= 0:
            m.map_kd = f"tex{m.diffuse_tex}.png"
            write_sidecar(scene.textures[m.diffuse_tex], os.path.join(directory, m.map_kd))
    obj = os.path.join(directory, "scene.obj")
    scenes.write_obj(sc, obj)
    bg = os.path.join(directory, "background.jpg")
    if scene.env is not None:
        write_sidecar(scene.env, bg)
    return obj, bg


class Ref:
    """The reference's ``Raytracer`` (pg1/raytracer.h:15-52) on a scene written to ``directory``."""

    def __init__(self, scene, directory: str):
        self.lib = _lib()
        c = scene.camera
        f = (C.c_float * 3)(*c.view_from); a = (C.c_float * 3)(*c.view_at)
        self.h = C.c_void_p(self.lib.ref_create(c.width, c.height, C.c_float(c.fov_y), f, a))
        self.width, self.height = c.width, c.height
        obj, bg = write_scene(scene, directory)
        if self.lib.ref_load_scene(self.h, obj.encode(), bg.encode()):
            raise RuntimeError("Raytracer::LoadScene failed")

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h); self.h = None

    def set_gamma(self, g: float):
        self.lib.ref_set_gamma(self.h, C.c_float(g))

    def trace(self, rays9: np.ndarray, level: int = 0, threads: int = 0) -> np.ndarray:
        rays9 = np.ascontiguousarray(rays9, np.float32); out = np.zeros((rays9.shape[0], 4), np.float32)
        self.lib.ref_trace(self.h, _p(rays9), C.c_uint64(rays9.shape[0]), int(level), _p(out), int(threads))
        return out

    def generate_rays(self, xy: np.ndarray, focal=200.0, aperture=0.0, pinhole=False) -> np.ndarray:
        xy = np.ascontiguousarray(xy, np.float32); out = np.zeros((xy.shape[0], 9), np.float32)
        self.lib.ref_generate_rays(self.h, _p(xy), C.c_uint64(xy.shape[0]), C.c_float(focal), C.c_float(aperture), int(pinhole), _p(out))
        return out

    def render_unjittered(self, focal=200.0, threads: int = 0) -> np.ndarray:
        out = np.zeros((self.height, self.width, 4), np.float32)
        self.lib.ref_render_unjittered(self.h, C.c_float(focal), _p(out), int(threads))
        return out

    def get_pixels(self, x0, y0, x1, y1, threads: int = 0) -> np.ndarray:
        """``Raytracer::get_pixel`` as shipped (clock-seeded 3x3 jitter and lens shift) over a window."""
        out = np.zeros((y1 - y0, x1 - x0, 4), np.float32)
        self.lib.ref_get_pixels(self.h, x0, y0, x1, y1, _p(out), int(threads))
        return out

    def gamma(self, c: np.ndarray) -> np.ndarray:
        c = np.ascontiguousarray(c, np.float32); out = np.zeros_like(c)
        self.lib.ref_gamma(self.h, _p(c), C.c_uint64(c.shape[0]), _p(out))
        return out

    def is_illuminated(self, light, hit, normal) -> np.ndarray:
        hit = np.ascontiguousarray(np.asarray(hit, np.float32).reshape(-1, 3)); n = hit.shape[0]
        light = np.ascontiguousarray(np.broadcast_to(np.asarray(light, np.float32).reshape(-1, 3), (n, 3)))
        normal = np.ascontiguousarray(np.broadcast_to(np.asarray(normal, np.float32).reshape(-1, 3), (n, 3)))
        out = np.zeros(n, np.int32)
        self.lib.ref_is_illuminated(self.h, _p(light), _p(hit), _p(normal), C.c_uint64(n), _p(out))
        return out.astype(bool)

    def secondary_rays(self, items: np.ndarray, refraction: bool) -> np.ndarray:
        items = np.ascontiguousarray(items, np.float32); out = np.zeros((items.shape[0], 9), np.float32)
        self.lib.ref_secondary_rays(self.h, _p(items), C.c_uint64(items.shape[0]), int(refraction), _p(out))
        return out


def mix_srgb(c0, c1, alpha) -> np.ndarray:
    c0 = np.ascontiguousarray(c0, np.float32); c1 = np.ascontiguousarray(c1, np.float32); alpha = np.ascontiguousarray(alpha, np.float32)
    out = np.zeros_like(c0)
    _lib().ref_mix_srgb(_p(c0), _p(c1), _p(alpha), C.c_uint64(c0.shape[0]), _p(out))
    return out


class RefTexture:
    """The reference's ``Texture`` (pg1/texture.cpp) on an image given as raw bytes (side-car next to ``path``)."""

    def __init__(self, image, path: str):
        write_sidecar(image, path)
        self.lib = _lib()
        self.h = C.c_void_p(self.lib.ref_texture_load(path.encode()))
        if not self.h:
            raise RuntimeError("Texture::Texture failed")

    def size(self):
        wh = (C.c_int * 2)(); self.lib.ref_texture_size(self.h, wh); return wh[0], wh[1]

    def get_texel(self, uv) -> np.ndarray:
        uv = np.ascontiguousarray(uv, np.float32); out = np.zeros((uv.shape[0], 3), np.float32)
        self.lib.ref_texture_get_texel(self.h, _p(uv), C.c_uint64(uv.shape[0]), _p(out))
        return out


class RefEnv:
    """The reference's ``SphericalMap`` (pg1/SphericalMap.cpp)."""

    def __init__(self, image, path: str):
        write_sidecar(image, path)
        self.lib = _lib()
        self.h = C.c_void_p(self.lib.ref_envmap_load(path.encode()))

    def get_texel(self, dirs) -> np.ndarray:
        dirs = np.ascontiguousarray(dirs, np.float32); out = np.zeros((dirs.shape[0], 4), np.float32)
        self.lib.ref_env_get_texel(self.h, _p(dirs), C.c_uint64(dirs.shape[0]), _p(out))
        return out


def load_obj(path: str):
    """``LoadOBJ`` (pg1/objloader.cpp:210-507) -> (surfaces [(name, material index, pos[T,3,3], nrm[T,3,3], uv[T,3,2])], materials [dict])."""
    lib = _lib()
    ns, nm = C.c_int(), C.c_int()
    h = C.c_void_p(lib.ref_obj_load(path.encode(), C.byref(ns), C.byref(nm)))
    surfaces, materials = [], []
    for s in range(ns.value):
        T = lib.ref_obj_surface_triangles(h, s)
        pos = np.zeros((T, 3, 3), np.float32); nrm = np.zeros((T, 3, 3), np.float32); uv = np.zeros((T, 3, 2), np.float32)
        name = C.create_string_buffer(64); mi = C.c_int()
        lib.ref_obj_surface(h, s, _p(pos), _p(nrm), _p(uv), name, C.byref(mi))
        surfaces.append((name.value.decode(errors="replace"), mi.value, pos, nrm, uv))
    for m in range(nm.value):
        out = np.zeros(10, np.float32); name = C.create_string_buffer(64)
        lib.ref_obj_material(h, m, _p(out), name)
        materials.append(dict(name=name.value.decode(errors="replace"), diffuse=tuple(out[0:3]), specular=tuple(out[3:6]), shininess=float(out[6]),
                              ior=float(out[7]), type=int(out[8]), has_diffuse_texture=bool(out[9])))
    lib.ref_obj_free(h)
    return surfaces, materials


def run_captured(fn, *args) -> str:
    """Run a library call that prints with printf and return what it wrote to stdout."""
    import tempfile
    import sys
    sys.stdout.flush()
    with tempfile.TemporaryFile() as tmp:
        saved = os.dup(1)
        os.dup2(tmp.fileno(), 1)
        try:
            fn(*args)
        finally:
            os.dup2(saved, 1); os.close(saved)
        tmp.seek(0)
        return tmp.read().decode(errors="replace")


def tutorial_1() -> str:
    lib = _lib()
    return run_captured(lib.ref_tutorial_1, b"threads=0,verbose=3")


def tutorial_2(test4_image, workdir: str) -> str:
    """``tutorial_2`` opens ``../../../data/test4.png`` relative to the working directory (pg1/tutorials.cpp:173)."""
    lib = _lib()
    deep = os.path.join(workdir, "a", "b", "c"); os.makedirs(deep, exist_ok=True); os.makedirs(os.path.join(workdir, "data"), exist_ok=True)
    write_sidecar(test4_image, os.path.join(workdir, "data", "test4.png"))
    cwd = os.getcwd()
    os.chdir(deep)
    try:
        return run_captured(lib.ref_tutorial_2)
    finally:
        os.chdir(cwd)
