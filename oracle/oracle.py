"""ctypes front-end of the CPU oracle (``oracle/pg_oracle.cpp``).

TEST INFRASTRUCTURE ONLY.  May be imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``; never by ``pgi_raytracing_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpg_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pg_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class OrcMaterial(C.Structure):
    _fields_ = [("diffuse", C.c_float * 3), ("specular", C.c_float * 3), ("shininess", C.c_float), ("ior", C.c_float),
                ("type", C.c_int32), ("diffuse_tex", C.c_int32)]


class OrcLight(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("ambient", C.c_float * 3), ("diffuse", C.c_float * 3), ("specular", C.c_float * 3)]


class OrcParams(C.Structure):
    _fields_ = [("sampling_width", C.c_int32), ("jitter", C.c_int32), ("focal_distance", C.c_float), ("aperture", C.c_float),
                ("max_depth", C.c_int32), ("gamma_level", C.c_float), ("seed", C.c_uint32), ("camera_mode", C.c_int32),
                ("shader_mode", C.c_int32), ("shadow_mode", C.c_int32), ("reserved", C.c_int32 * 6)]


class OrcStats(C.Structure):
    _fields_ = [("rays_primary", C.c_uint64), ("rays_shadow", C.c_uint64), ("rays_reflection", C.c_uint64), ("rays_refraction", C.c_uint64)]


RAYHIT_DTYPE = np.dtype([("org_x", "f4"), ("org_y", "f4"), ("org_z", "f4"), ("tnear", "f4"), ("dir_x", "f4"), ("dir_y", "f4"),
                         ("dir_z", "f4"), ("time", "f4"), ("tfar", "f4"), ("mask", "u4"), ("id", "u4"), ("flags", "u4"),
                         ("Ng_x", "f4"), ("Ng_y", "f4"), ("Ng_z", "f4"), ("u", "f4"), ("v", "f4"), ("primID", "u4"),
                         ("geomID", "u4"), ("instID", "u4")])
assert RAYHIT_DTYPE.itemsize == 80


def _lib():
    build()
    lib = C.CDLL(_LIB_PATH)
    lib.orc_create.restype = C.c_void_p
    lib.orc_path_bounce.restype = None
    lib.orc_path_bounce.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_void_p]
    lib.orc_rng_u01.restype = C.c_float
    lib.orc_rng_u01.argtypes = [C.c_uint32] * 4
    lib.orc_num_triangles.restype = C.c_uint32
    lib.orc_num_nodes.restype = C.c_uint32
    return lib


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t) if a is not None else None


def make_params(sampling_width=3, jitter=1, focal_distance=200.0, aperture=5.0, max_depth=7, gamma_level=0.5, seed=1,
                camera_mode=0, shader_mode=0, shadow_mode=0) -> dict:
    """Defaults = the values hard-coded in the reference (pg1/raytracer.cpp:398-400,282,450)."""
    return dict(sampling_width=sampling_width, jitter=jitter, focal_distance=focal_distance, aperture=aperture, max_depth=max_depth,
                gamma_level=gamma_level, seed=seed, camera_mode=camera_mode, shader_mode=shader_mode, shadow_mode=shadow_mode)


class Oracle:
    """One scene + camera held by the CPU oracle."""

    def __init__(self, scene=None):
        self.lib = _lib()
        self.h = C.c_void_p(self.lib.orc_create())
        self.width = self.height = 0
        if scene is not None:
            self.load(scene)

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- scene
    def load(self, scene, commit=True):
        for m in scene.meshes:
            g = C.c_uint32()
            self.lib.orc_add_mesh(self.h, _p(m.pos), _p(m.nrm), _p(m.uv), C.c_uint32(m.ntris), C.c_int32(m.material), C.byref(g))
        mats = (OrcMaterial * len(scene.materials))()
        for i, mt in enumerate(scene.materials):
            mats[i].diffuse[:] = mt.diffuse; mats[i].specular[:] = mt.specular
            mats[i].shininess = mt.shininess; mats[i].ior = mt.ior; mats[i].type = mt.type; mats[i].diffuse_tex = mt.diffuse_tex
        self.lib.orc_set_materials(self.h, mats, len(scene.materials))
        for i, t in enumerate(scene.textures):
            self.lib.orc_set_texture(self.h, i, _p(t.data), t.width, t.height, t.pitch, t.bpp)
        if scene.env is not None:
            e = scene.env
            self.lib.orc_set_envmap(self.h, _p(e.data), e.width, e.height, e.pitch, e.bpp)
        ls = (OrcLight * len(scene.lights))()
        for i, l in enumerate(scene.lights):
            ls[i].position[:] = l.position; ls[i].ambient[:] = l.ambient; ls[i].diffuse[:] = l.diffuse; ls[i].specular[:] = l.specular
        self.lib.orc_set_lights(self.h, ls, len(scene.lights))
        c = scene.camera
        self.set_camera(c.width, c.height, c.fov_y, c.view_from, c.view_at)
        if commit:
            self.lib.orc_commit(self.h)

    def set_camera(self, w, h, fov_y, view_from, view_at):
        f = (C.c_float * 3)(*view_from); a = (C.c_float * 3)(*view_at)
        self.lib.orc_set_camera(self.h, w, h, C.c_float(fov_y), f, a)
        self.width, self.height = w, h

    def camera_constants(self):
        out = np.zeros(10, np.float32)
        self.lib.orc_get_camera(self.h, _p(out))
        return out

    # ---- path
    @staticmethod
    def _params(p: dict) -> OrcParams:
        q = OrcParams()
        for k, v in p.items():
            setattr(q, k, v)
        return q

    def render(self, params: dict, want_ids=True, brute=False, threads=0, region=None):
        W, H = self.width, self.height
        rgba = np.zeros((H, W, 4), np.float32)
        geom = np.full((H, W), 0xFFFFFFFF, np.uint32) if want_ids else None
        prim = np.full((H, W), 0xFFFFFFFF, np.uint32) if want_ids else None
        st = OrcStats()
        q = self._params(params)
        x0, y0, x1, y1 = region if region is not None else (0, 0, W, H)
        rc = self.lib.orc_render_region(self.h, C.byref(q), x0, y0, x1, y1, _p(rgba), _p(geom), _p(prim), C.byref(st), int(brute), int(threads))
        if rc:
            raise RuntimeError(f"orc_render_region failed: {rc}")
        stats = dict(primary=st.rays_primary, shadow=st.rays_shadow, reflection=st.rays_reflection, refraction=st.rays_refraction)
        stats["total"] = sum(stats.values())
        return rgba, geom, prim, stats

    def intersect(self, rayhits: np.ndarray, brute=False, threads=0) -> np.ndarray:
        rh = np.ascontiguousarray(rayhits.copy())
        assert rh.dtype == RAYHIT_DTYPE
        rc = self.lib.orc_intersect(self.h, _p(rh), C.c_uint64(rh.shape[0]), int(brute), int(threads))
        if rc:
            raise RuntimeError(f"orc_intersect failed: {rc}")
        return rh

    def trace(self, params: dict, rays9: np.ndarray, level: int = 0, brute=False, threads=0) -> np.ndarray:
        """``Raytracer::trace(ray, level)`` on caller-supplied rays [n, 9] = (org, tnear, dir, time, tfar) -> Color4f [n, 4]."""
        rays9 = np.ascontiguousarray(rays9, np.float32); out = np.zeros((rays9.shape[0], 4), np.float32)
        q = self._params(params)
        rc = self.lib.orc_trace(self.h, C.byref(q), _p(rays9), C.c_uint64(rays9.shape[0]), int(level), _p(out), int(brute), int(threads))
        if rc:
            raise RuntimeError(f"orc_trace failed: {rc}")
        return out

    def is_illuminated(self, params: dict, light, hit, normal, brute=False) -> np.ndarray:
        """``Raytracer::is_illuminated(light, hit_position, normal)`` per query -> bool [n]."""
        hit = np.ascontiguousarray(np.asarray(hit, np.float32).reshape(-1, 3)); n = hit.shape[0]
        light = np.ascontiguousarray(np.broadcast_to(np.asarray(light, np.float32).reshape(-1, 3), (n, 3)))
        normal = np.ascontiguousarray(np.broadcast_to(np.asarray(normal, np.float32).reshape(-1, 3), (n, 3)))
        out = np.zeros(n, np.int32); q = self._params(params)
        rc = self.lib.orc_is_illuminated(self.h, C.byref(q), _p(light), _p(hit), _p(normal), C.c_uint64(n), _p(out), int(brute))
        if rc:
            raise RuntimeError(f"orc_is_illuminated failed: {rc}")
        return out.astype(bool)

    def primary_rays(self, params: dict) -> np.ndarray:
        w = params["sampling_width"]
        out = np.zeros((self.height * self.width * w * w, 9), np.float32)
        q = self._params(params)
        self.lib.orc_primary_rays(self.h, C.byref(q), _p(out))
        return out

    def texture_get_texel(self, tex_id: int, uv: np.ndarray) -> np.ndarray:
        uv = np.ascontiguousarray(uv, np.float32); out = np.zeros((uv.shape[0], 3), np.float32)
        rc = self.lib.orc_texture_get_texel(self.h, tex_id, _p(uv), C.c_uint64(uv.shape[0]), _p(out))
        if rc:
            raise RuntimeError("no such texture")
        return out

    def env_get_texel(self, dirs: np.ndarray) -> np.ndarray:
        dirs = np.ascontiguousarray(dirs, np.float32); out = np.zeros((dirs.shape[0], 4), np.float32)
        self.lib.orc_env_get_texel(self.h, _p(dirs), C.c_uint64(dirs.shape[0]), _p(out))
        return out

    # ---- stateless helpers
    def mix_srgb(self, c0, c1, alpha):
        c0 = np.ascontiguousarray(c0, np.float32); c1 = np.ascontiguousarray(c1, np.float32); alpha = np.ascontiguousarray(alpha, np.float32)
        out = np.zeros_like(c0)
        self.lib.orc_mix_srgb(_p(c0), _p(c1), _p(alpha), C.c_uint64(c0.shape[0]), _p(out))
        return out

    def gamma(self, c, gamma_level):
        c = np.ascontiguousarray(c, np.float32); out = np.zeros_like(c)
        self.lib.orc_gamma(_p(c), C.c_float(gamma_level), C.c_uint64(c.shape[0]), _p(out))
        return out

    def secondary_rays(self, items: np.ndarray, refraction: bool) -> np.ndarray:
        items = np.ascontiguousarray(items, np.float32); out = np.zeros((items.shape[0], 9), np.float32)
        self.lib.orc_secondary_rays(_p(items), C.c_uint64(items.shape[0]), int(refraction), _p(out))
        return out

    def interpolate(self, geom, prim, u, v, slot: int) -> np.ndarray:
        geom = np.ascontiguousarray(geom, np.uint32); prim = np.ascontiguousarray(prim, np.uint32)
        u = np.ascontiguousarray(u, np.float32); v = np.ascontiguousarray(v, np.float32)
        out = np.zeros((geom.shape[0], 3 if slot == 0 else 2), np.float32)
        rc = self.lib.orc_interpolate(self.h, _p(geom), _p(prim), _p(u), _p(v), C.c_uint64(geom.shape[0]), int(slot), _p(out))
        if rc:
            raise RuntimeError("orc_interpolate: id out of range")
        return out

    def path_bounce(self, normal, hits: np.ndarray, seed: int, level: int) -> np.ndarray:
        """Directions of the diffuse bounce of shader_mode 3 at the given hit points (one normal for all)."""
        n = np.ascontiguousarray(np.asarray(normal, np.float32)); h = np.ascontiguousarray(np.asarray(hits, np.float32).reshape(-1, 3))
        out = np.zeros_like(h)
        self.lib.orc_path_bounce(_p(n), _p(h), C.c_uint64(h.shape[0]), int(seed), int(level), _p(out))
        return out

    def rng_u01(self, seed, pixel, sample, dim) -> float:
        return float(self.lib.orc_rng_u01(seed, pixel, sample, dim))

    def max_threads(self) -> int:
        return int(self.lib.orc_max_threads())

    def num_triangles(self) -> int:
        return int(self.lib.orc_num_triangles(self.h))


def make_rayhits(org, direction, tnear=0.0, tfar=np.inf, time=0.0) -> np.ndarray:
    org = np.asarray(org, np.float32).reshape(-1, 3); direction = np.asarray(direction, np.float32).reshape(-1, 3)
    n = org.shape[0]
    rh = np.zeros(n, RAYHIT_DTYPE)
    rh["org_x"], rh["org_y"], rh["org_z"] = org[:, 0], org[:, 1], org[:, 2]
    rh["dir_x"], rh["dir_y"], rh["dir_z"] = direction[:, 0], direction[:, 1], direction[:, 2]
    rh["tnear"] = tnear; rh["tfar"] = np.float32(np.finfo(np.float32).max) if np.isinf(tfar) else tfar; rh["time"] = time
    rh["geomID"] = 0xFFFFFFFF; rh["primID"] = 0xFFFFFFFF; rh["instID"] = 0xFFFFFFFF
    return rh
