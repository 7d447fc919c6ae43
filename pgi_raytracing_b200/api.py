"""Host-side mirror of the reference's ``Raytracer`` surface (pg1/raytracer.h:15-52) over the C ABI.

Same construction arguments, same method roles and the same hard-coded defaults as the reference; every
method that computes goes through ``libpgrt_b200.so`` (CUDA).  Nothing here touches ``oracle/``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .scenes import Scene

RAYHIT_DTYPE = np.dtype([("org_x", "f4"), ("org_y", "f4"), ("org_z", "f4"), ("tnear", "f4"), ("dir_x", "f4"), ("dir_y", "f4"),
                         ("dir_z", "f4"), ("time", "f4"), ("tfar", "f4"), ("mask", "u4"), ("id", "u4"), ("flags", "u4"),
                         ("Ng_x", "f4"), ("Ng_y", "f4"), ("Ng_z", "f4"), ("u", "f4"), ("v", "f4"), ("primID", "u4"),
                         ("geomID", "u4"), ("instID", "u4")])   # RTCRayHit, embree3/rtcore_ray.h:11-49
assert RAYHIT_DTYPE.itemsize == 80


class PgrtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pgrt error {code}: {msg}")
        self.code = code


def default_params(**over) -> L.RenderParams:
    """``pgrt_default_params``: the values get_pixel / trace hard-code (pg1/raytracer.cpp:398-400, :282, :450)."""
    p = L.RenderParams()
    L.load().pgrt_default_params(C.byref(p))
    for k, v in over.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Raytracer:
    """``Raytracer(width, height, fov_y, view_from, view_at)`` (pg1/raytracer.cpp:12-19), GPU-backed."""

    def __init__(self, width: int, height: int, fov_y: float, view_from, view_at, device: int = 0):
        self.lib = L.load()
        h = C.c_void_p()
        rc = self.lib.pgrt_create(C.byref(h), device)
        if rc != L.PGRT_OK:
            raise PgrtError(rc, "pgrt_create failed (no CUDA device? the render loop has no CPU fallback)")
        self.h = h
        self.width, self.height = width, height
        self.gamma_level = 0.5          # public member of the reference class (raytracer.h:23), UI default raytracer.cpp:450
        self.build_stats = None
        self.set_camera(width, height, fov_y, view_from, view_at)

    # ---- plumbing
    def _check(self, rc):
        if rc != L.PGRT_OK:
            raise PgrtError(rc, self.lib.pgrt_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.pgrt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr: int | None):
        self._check(self.lib.pgrt_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)))

    # ---- camera (PinHoleCamera ctor)
    def set_camera(self, width, height, fov_y, view_from, view_at):
        f = (C.c_float * 3)(*view_from); a = (C.c_float * 3)(*view_at)
        self._check(self.lib.pgrt_set_camera(self.h, width, height, C.c_float(fov_y), f, a))
        self.width, self.height = width, height

    # ---- LoadScene (pg1/raytracer.cpp:48-128): surfaces -> meshes, materials, textures, env map, light, commit
    def LoadScene(self, scene: Scene):
        self._check(self.lib.pgrt_clear_scene(self.h))
        for m in scene.meshes:
            g = C.c_uint32()
            self._check(self.lib.pgrt_add_mesh(self.h, _ptr(m.pos), _ptr(m.nrm), _ptr(m.uv), m.ntris, m.material, C.byref(g)))
        mats = (L.Material * max(len(scene.materials), 1))()
        for i, mt in enumerate(scene.materials):
            mats[i].diffuse[:] = mt.diffuse; mats[i].specular[:] = mt.specular
            mats[i].shininess = mt.shininess; mats[i].ior = mt.ior; mats[i].type = mt.type; mats[i].diffuse_tex = mt.diffuse_tex
        self._check(self.lib.pgrt_set_materials(self.h, mats, len(scene.materials)))
        for i, t in enumerate(scene.textures):
            self._check(self.lib.pgrt_set_texture(self.h, i, _ptr(t.data), t.width, t.height, t.pitch, t.bpp))
        if scene.env is not None:
            e = scene.env
            self._check(self.lib.pgrt_set_envmap(self.h, _ptr(e.data), e.width, e.height, e.pitch, e.bpp))
        ls = (L.Light * max(len(scene.lights), 1))()
        for i, l in enumerate(scene.lights):
            ls[i].position[:] = l.position; ls[i].ambient[:] = l.ambient; ls[i].diffuse[:] = l.diffuse; ls[i].specular[:] = l.specular
        self._check(self.lib.pgrt_set_lights(self.h, ls, len(scene.lights)))
        return self.commit()

    def commit(self) -> dict:
        bs = L.BuildStats()
        self._check(self.lib.pgrt_commit(self.h, C.byref(bs)))
        self.build_stats = {k: getattr(bs, k) for k, _ in L.BuildStats._fields_ if k != "reserved"}
        return self.build_stats

    # ---- the path
    def _params(self, params) -> L.RenderParams:
        if params is None:
            params = default_params(gamma_level=self.gamma_level)
        elif isinstance(params, dict):
            params = default_params(**params)
        return params

    @staticmethod
    def _stats(rs: L.RenderStats) -> dict:
        d = dict(primary=rs.rays_primary, shadow=rs.rays_shadow, reflection=rs.rays_reflection, refraction=rs.rays_refraction,
                 frame_ms=rs.frame_ms, trace_ms=rs.trace_ms, shade_ms=rs.shade_ms, trace_launches=rs.trace_launches,
                 launches=rs.launches, batches=rs.batches, overflow_retries=rs.overflow_retries,
                 nodes_visited=rs.nodes_visited, tris_tested=rs.tris_tested, max_nodes_per_ray=rs.max_nodes_per_ray,
                 pool_peak=rs.reserved[0], kernel_us=rs.reserved[1], primary_phase_us=rs.reserved[2], pool_iters=rs.reserved[3])
        d["total"] = d["primary"] + d["shadow"] + d["reflection"] + d["refraction"]
        return d

    def render(self, params=None, out: np.ndarray | None = None, profile: bool = False):
        """One ``Producer`` iteration (pg1/simpleguidx11.cpp:95-118) into a host float32 [H,W,4] array."""
        p = self._params(params)
        if out is None:
            out = np.empty((self.height, self.width, 4), np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == self.width * self.height * 4
        rs = L.RenderStats()
        self._check(self.lib.pgrt_render(self.h, C.byref(p), _ptr(out), C.byref(rs), int(profile)))
        return out, self._stats(rs)

    def render_rgba8(self, params=None, out: np.ndarray | None = None, profile: bool = False):
        """The same frame as R8G8B8A8_UNORM [H,W,4] uint8 (``pgrt_render_rgba8``): what the reference's swap chain shows
        (pg1/simpleguidx11.cpp:229,290)."""
        p = self._params(params)
        if out is None:
            out = np.empty((self.height, self.width, 4), np.uint8)
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.size == self.width * self.height * 4
        rs = L.RenderStats()
        self._check(self.lib.pgrt_render_rgba8(self.h, C.byref(p), _ptr(out), C.byref(rs), int(profile)))
        return out, self._stats(rs)

    def render_accumulate(self, n_frames: int, params=None, out: np.ndarray | None = None):
        """Mean of ``n_frames`` finished frames with seeds seed, seed+1, ... (``pgrt_render_accumulate``; summed on the
        device in frame order).  The progressive loop the reference lacks (it re-renders from scratch each iteration)."""
        p = self._params(params)
        if out is None:
            out = np.empty((self.height, self.width, 4), np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == self.width * self.height * 4
        rs = L.RenderStats()
        self._check(self.lib.pgrt_render_accumulate(self.h, C.byref(p), int(n_frames), _ptr(out), C.byref(rs)))
        return out, self._stats(rs)

    def render_device(self, device_ptr: int, params=None, profile: bool = False) -> dict:
        p = self._params(params); rs = L.RenderStats()
        self._check(self.lib.pgrt_render_device(self.h, C.byref(p), C.c_void_p(device_ptr), C.byref(rs), int(profile)))
        return self._stats(rs)

    def render_host_ptr(self, host_ptr: int, params=None, profile: bool = False) -> dict:
        p = self._params(params); rs = L.RenderStats()
        self._check(self.lib.pgrt_render(self.h, C.byref(p), C.c_void_p(host_ptr), C.byref(rs), int(profile)))
        return self._stats(rs)

    # ---- pipelined frames (pgrt_render*_begin / pgrt_render_end): up to MAX_INFLIGHT frames in flight
    def render_begin(self, slot: int, params=None, *, host_ptr: int = 0, device_ptr: int = 0, shard_ptr: int = 0, frame_ptr: int = 0,
                     profile: int = 0, rgba8: bool = False):
        """Enqueue one frame in ``slot`` and return without waiting for the GPU.  Exactly one destination
        (``frame_ptr``: sharded context, tiles stored at their final place of a possibly peer-mapped full frame).
        ``rgba8``: the destination (``host_ptr`` or ``frame_ptr``) is an R8G8B8A8_UNORM frame."""
        p = self._params(params)
        if rgba8 and frame_ptr:
            rc = self.lib.pgrt_render_shard_to_frame_rgba8_begin(self.h, C.byref(p), C.c_void_p(frame_ptr), slot, int(profile))
        elif rgba8:
            rc = self.lib.pgrt_render_rgba8_begin(self.h, C.byref(p), C.c_void_p(host_ptr), slot, int(profile))
        elif frame_ptr:
            rc = self.lib.pgrt_render_shard_to_frame_begin(self.h, C.byref(p), C.c_void_p(frame_ptr), slot, int(profile))
        elif host_ptr:
            rc = self.lib.pgrt_render_begin(self.h, C.byref(p), C.c_void_p(host_ptr), slot, int(profile))
        elif device_ptr:
            rc = self.lib.pgrt_render_device_begin(self.h, C.byref(p), C.c_void_p(device_ptr), slot, int(profile))
        else:
            rc = self.lib.pgrt_render_shard_device_begin(self.h, C.byref(p), C.c_void_p(shard_ptr), slot, int(profile))
        self._check(rc)

    def render_end(self, slot: int) -> dict:
        rs = L.RenderStats()
        self._check(self.lib.pgrt_render_end(self.h, slot, C.byref(rs)))
        return self._stats(rs)

    # ---- frames shared between the processes of one box (CUDA IPC)
    def frame_alloc(self, nbytes: int) -> int:
        ptr = C.c_void_p()
        self._check(self.lib.pgrt_frame_alloc(self.h, nbytes, C.byref(ptr)))
        return int(ptr.value)

    def frame_free(self, ptr: int):
        self._check(self.lib.pgrt_frame_free(self.h, C.c_void_p(ptr)))

    def frame_export(self, ptr: int) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self.lib.pgrt_frame_export(self.h, C.c_void_p(ptr), buf))
        return buf.raw

    def frame_import(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        self._check(self.lib.pgrt_frame_import(self.h, handle, C.byref(ptr)))
        return int(ptr.value)

    def frame_unmap(self, ptr: int):
        self._check(self.lib.pgrt_frame_unmap(self.h, C.c_void_p(ptr)))

    def host_frame_register(self, host_ptr: int, nbytes: int) -> int:
        """Register page-aligned (shared) host memory with this rank's device; returns the address kernels store through."""
        ptr = C.c_void_p()
        self._check(self.lib.pgrt_host_frame_register(self.h, C.c_void_p(host_ptr), nbytes, C.byref(ptr)))
        return int(ptr.value)

    def host_frame_unregister(self, host_ptr: int):
        self._check(self.lib.pgrt_host_frame_unregister(self.h, C.c_void_p(host_ptr)))

    def flush_l2(self, slot: int, nbytes: int = 160 << 20, value: int = 0):
        """Measurement helper (``pgrt_debug_flush_l2``): evict L2 on the slot's stream before a timed frame."""
        self._check(self.lib.pgrt_debug_flush_l2(self.h, slot, nbytes, value))

    def frame_cycles(self, slot: int = 0):
        """``pgrt_debug_frame_cycles`` (library built with -DPGRT_FRAME_TIMING): warp cycles on primary chunks, on pool records, resident; warps."""
        out = (C.c_uint64 * 4)()
        self._check(self.lib.pgrt_debug_frame_cycles(self.h, slot, out))
        return [int(x) for x in out]

    def l2_bandwidth(self, nbytes: int = 32 << 20, iters: int = 50) -> float:
        """Measurement helper (``pgrt_debug_l2_bandwidth``): GB/s of an L2-resident buffer read from all SMs, L1 bypassed."""
        g = C.c_float()
        self._check(self.lib.pgrt_debug_l2_bandwidth(self.h, nbytes, iters, C.byref(g)))
        return float(g.value)

    def enable_peer_access(self, peer_device: int):
        self._check(self.lib.pgrt_enable_peer_access(self.h, peer_device))

    def slot_stream(self, slot: int) -> int:
        return int(self.lib.pgrt_slot_stream(self.h, slot) or 0)

    def stream_wait_slot(self, slot: int, cuda_stream_ptr: int):
        self._check(self.lib.pgrt_stream_wait_slot(self.h, slot, C.c_void_p(cuda_stream_ptr)))

    def slot_signal(self, slot: int, flag_ptr: int, value: int):
        """Frames begun on ``slot`` from now on store ``value`` in ``*flag_ptr`` when they finish (``pgrt_slot_signal``)."""
        self._check(self.lib.pgrt_slot_signal(self.h, slot, C.c_void_p(flag_ptr), value))

    def slot_signal_add(self, slot: int, counter_ptr: int):
        """Frames begun on ``slot`` from now on add 1 to ``*counter_ptr`` when they finish (``pgrt_slot_signal_add``)."""
        self._check(self.lib.pgrt_slot_signal_add(self.h, slot, C.c_void_p(counter_ptr)))

    def stream_wait_value32(self, cuda_stream_ptr: int, flag_ptr: int, value: int):
        self._check(self.lib.pgrt_stream_wait_value32(self.h, C.c_void_p(cuda_stream_ptr), C.c_void_p(flag_ptr), value))

    def stream_write_value32(self, cuda_stream_ptr: int, flag_ptr: int, value: int):
        self._check(self.lib.pgrt_stream_write_value32(self.h, C.c_void_p(cuda_stream_ptr), C.c_void_p(flag_ptr), value))

    def get_pixel(self, x: int, y: int, t: float = 0.0, params=None):
        """``Color4f get_pixel(x, y, t)`` (pg1/raytracer.cpp:396-437); ``t`` is ignored, as in the reference."""
        p = self._params(params)
        px = (C.c_float * 4)()
        self._check(self.lib.pgrt_get_pixel(self.h, C.byref(p), x, y, px))
        return tuple(px)

    def primary_ids(self, params=None):
        p = self._params(params)
        g = np.empty((self.height, self.width), np.uint32); pr = np.empty((self.height, self.width), np.uint32)
        self._check(self.lib.pgrt_primary_ids(self.h, C.byref(p), _ptr(g), _ptr(pr)))
        return g, pr

    # ---- sharding
    def set_shard(self, rank: int, n_ranks: int):
        self._check(self.lib.pgrt_set_shard(self.h, rank, n_ranks))

    def shard_pixels(self) -> int:
        return int(self.lib.pgrt_shard_pixels(self.h))

    def render_shard_device(self, device_ptr: int, params=None, profile: bool = False) -> dict:
        p = self._params(params); rs = L.RenderStats()
        self._check(self.lib.pgrt_render_shard_device(self.h, C.byref(p), C.c_void_p(device_ptr), C.byref(rs), int(profile)))
        return self._stats(rs)

    def untile(self, gathered_ptr: int, n_ranks: int, out_ptr: int, cuda_stream_ptr: int | None = None):
        if cuda_stream_ptr is None:
            self._check(self.lib.pgrt_untile(self.h, C.c_void_p(gathered_ptr), n_ranks, C.c_void_p(out_ptr)))
        else:
            self._check(self.lib.pgrt_untile_on_stream(self.h, C.c_void_p(gathered_ptr), n_ranks, C.c_void_p(out_ptr), C.c_void_p(cuda_stream_ptr)))

    # ---- rtcIntersect1 / rtcInterpolate0 batches
    def intersect(self, rayhits: np.ndarray) -> np.ndarray:
        rh = np.ascontiguousarray(rayhits.copy())
        assert rh.dtype == RAYHIT_DTYPE
        self._check(self.lib.pgrt_intersect(self.h, _ptr(rh), rh.shape[0]))
        return rh

    def interpolate(self, geom, prim, u, v, slot: int) -> np.ndarray:
        geom = np.ascontiguousarray(geom, np.uint32); prim = np.ascontiguousarray(prim, np.uint32)
        u = np.ascontiguousarray(u, np.float32); v = np.ascontiguousarray(v, np.float32)
        out = np.empty((geom.shape[0], 3 if slot == 0 else 2), np.float32)
        self._check(self.lib.pgrt_interpolate(self.h, _ptr(geom), _ptr(prim), _ptr(u), _ptr(v), geom.shape[0], slot, _ptr(out)))
        return out

    # ---- the two public query methods of the class (pg1/raytracer.h:31, :34)
    def trace(self, rays9, level: int = 0, params=None) -> np.ndarray:
        """``Color4f trace(RTCRay ray, int level)`` over a batch: rays9 [n, 9] = (org, tnear, dir, time, tfar) -> [n, 4]."""
        p = self._params(params)
        rays9 = np.ascontiguousarray(rays9, np.float32)
        rays = np.zeros((rays9.shape[0], 12), np.float32); rays[:, :9] = rays9      # RTCRay: mask, id, flags = 0
        out = np.empty((rays9.shape[0], 4), np.float32)
        self._check(self.lib.pgrt_trace(self.h, C.byref(p), _ptr(rays), rays.shape[0], int(level), _ptr(out)))
        return out

    def is_illuminated(self, light, hit, normal, params=None) -> np.ndarray:
        """``bool is_illuminated(LightSource light, Vector3 hit_position, Vector3 normal)`` over a batch -> bool [n]."""
        p = self._params(params)
        hit = np.ascontiguousarray(np.asarray(hit, np.float32).reshape(-1, 3)); n = hit.shape[0]
        light = np.ascontiguousarray(np.broadcast_to(np.asarray(light, np.float32).reshape(-1, 3), (n, 3)))
        normal = np.ascontiguousarray(np.broadcast_to(np.asarray(normal, np.float32).reshape(-1, 3), (n, 3)))
        out = np.empty(n, np.int32)
        self._check(self.lib.pgrt_is_illuminated(self.h, C.byref(p), _ptr(light), _ptr(hit), _ptr(normal), n, _ptr(out)))
        return out.astype(bool)

    # ---- device leaf functions (per-function parity)
    def mix_srgb(self, c0, c1, alpha):
        c0 = np.ascontiguousarray(c0, np.float32); c1 = np.ascontiguousarray(c1, np.float32); alpha = np.ascontiguousarray(alpha, np.float32)
        out = np.empty_like(c0)
        self._check(self.lib.pgrt_eval_mix_srgb(self.h, _ptr(c0), _ptr(c1), _ptr(alpha), c0.shape[0], _ptr(out)))
        return out

    def texture_get_texel(self, tex_id: int, uv):
        uv = np.ascontiguousarray(uv, np.float32); out = np.empty((uv.shape[0], 3), np.float32)
        self._check(self.lib.pgrt_eval_texture(self.h, tex_id, _ptr(uv), uv.shape[0], _ptr(out)))
        return out

    def env_get_texel(self, dirs):
        dirs = np.ascontiguousarray(dirs, np.float32); out = np.empty((dirs.shape[0], 4), np.float32)
        self._check(self.lib.pgrt_eval_envmap(self.h, _ptr(dirs), dirs.shape[0], _ptr(out)))
        return out

    def gamma(self, c, gamma_level=None):
        c = np.ascontiguousarray(c, np.float32); out = np.empty_like(c)
        g = self.gamma_level if gamma_level is None else gamma_level
        self._check(self.lib.pgrt_eval_gamma(self.h, _ptr(c), C.c_float(g), c.shape[0], _ptr(out)))
        return out

    def primary_rays(self, params=None):
        p = self._params(params)
        out = np.empty((self.width * self.height * p.sampling_width ** 2, 9), np.float32)
        self._check(self.lib.pgrt_eval_primary_rays(self.h, C.byref(p), _ptr(out)))
        return out

    def secondary_rays(self, items, refraction: bool):
        items = np.ascontiguousarray(items, np.float32); out = np.empty((items.shape[0], 9), np.float32)
        self._check(self.lib.pgrt_eval_secondary_rays(self.h, _ptr(items), items.shape[0], int(refraction), _ptr(out)))
        return out

    def level_stats(self) -> list:
        """Per-recursion-level counters of the last frame (``pgrt_last_level_stats``)."""
        out = []
        for l in range(64):
            ls = L.LevelStats()
            if self.lib.pgrt_last_level_stats(self.h, l, C.byref(ls)) != L.PGRT_OK:
                break
            out.append({k: getattr(ls, k) for k, _ in L.LevelStats._fields_})
        return out

    def kernel_launches(self) -> int:
        return int(self.lib.pgrt_kernel_launches(self.h))


def raytracer_for(scene: Scene, device: int = 0) -> Raytracer:
    """``raytrace_loop`` without the window (pg1/tutorials.cpp:181-200): construct + LoadScene."""
    c = scene.camera
    rt = Raytracer(c.width, c.height, c.fov_y, c.view_from, c.view_at, device=device)
    rt.LoadScene(scene)
    return rt


def to_srgb8(rgba: np.ndarray) -> np.ndarray:
    """What D3D11 does when the float texture reaches the R8G8B8A8_UNORM back buffer
    (pg1/simpleguidx11.cpp:229,290): q = round(clamp(c,0,1)*255), NaN -> 0."""
    c = np.nan_to_num(rgba[..., :3].astype(np.float32), nan=0.0, posinf=1.0, neginf=0.0)
    return np.floor(np.clip(c, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)
