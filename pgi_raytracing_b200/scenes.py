"""Seeded procedural inputs for the render loop: stand-in scenes, textures, env maps, OBJ/MTL writer.

The reference's named geometry is absent from its own tree (no
``data/6887_allied_avenger.obj``, no PalmTrees mesh, default env map missing) and ``/root/reference`` does
not exist on the GPU box, so every workload is generated here, deterministically, and labelled
"stand-in" wherever it is reported.  The generated scenes keep what the path depends on:

* the camera and light the reference hard-codes (``pg1/tutorials.cpp:186-191``, ``pg1/raytracer.cpp:66-68``),
* the five materials of ``data/6887_allied_avenger.mtl`` (values restated, including the malformed
  ``Ks 1.0. 1.0 1.0`` line whose parse is part of the contract),
* un-indexed per-corner positions / normals / uv per surface, surface order = geomID
  (``pg1/raytracer.cpp:71-125``).

All arrays are float32, C-contiguous: pos [T,3,3], nrm [T,3,3], uv [T,3,2].
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np

IOR_AIR = 1.000293  # pg1/material.h:15


# --------------------------------------------------------------------------------------- containers
@dataclass
class Material:
    """Fields of ``Material`` that the hot path reads (pg1/material.h:96-108)."""
    name: str = "default"
    diffuse: tuple = (0.4, 0.4, 0.4)     # pg1/material.cpp:13
    specular: tuple = (0.8, 0.8, 0.8)    # pg1/material.cpp:14
    shininess: float = 1.0               # pg1/material.cpp:19
    ior: float = 1.5                     # pg1/material.cpp:21
    type: int = 3                        # MTL "shader N" (uninitialised in the reference's default ctor)
    diffuse_tex: int = -1                # index into Scene.textures, -1 = none
    ambient: tuple = (0.1, 0.1, 0.1)
    map_kd: str = ""


@dataclass
class Image:
    """Raw top-down BGR(A) bytes exactly as ``Texture::Texture`` leaves them (pg1/texture.cpp:36-47)."""
    data: np.ndarray  # uint8 [height, pitch]
    width: int
    height: int
    pitch: int
    bpp: int

    @staticmethod
    def from_rgb(rgb: np.ndarray, alpha: np.ndarray | None = None) -> "Image":
        """rgb uint8 [H,W,3] (+ optional alpha [H,W]) -> BGR(A) rows padded to 4 bytes (FreeImage pitch)."""
        h, w, _ = rgb.shape
        bpp = 4 if alpha is not None else 3
        pitch = (w * bpp + 3) // 4 * 4
        buf = np.zeros((h, pitch), dtype=np.uint8)
        px = buf[:, : w * bpp].reshape(h, w, bpp)
        px[..., 0] = rgb[..., 2]
        px[..., 1] = rgb[..., 1]
        px[..., 2] = rgb[..., 0]
        if alpha is not None:
            px[..., 3] = alpha
        return Image(np.ascontiguousarray(buf), w, h, pitch, bpp)


@dataclass
class Mesh:
    name: str
    pos: np.ndarray
    nrm: np.ndarray
    uv: np.ndarray
    material: int

    @property
    def ntris(self) -> int:
        return int(self.pos.shape[0])


@dataclass
class Light:
    position: tuple = (-500.0, -100.0, 500.0)   # pg1/raytracer.cpp:67
    ambient: tuple = (1.0, 1.0, 1.0)
    diffuse: tuple = (1.0, 1.0, 1.0)
    specular: tuple = (1.0, 1.0, 1.0)


@dataclass
class Camera:
    width: int = 640                              # pg1/tutorials.cpp:186-191
    height: int = 480
    fov_y: float = float(np.float32(42.185) * np.float32(math.pi) / np.float32(180.0))  # deg2rad, pg1/mymath.h:27-30
    view_from: tuple = (-140.0, -175.0, 80.0)
    view_at: tuple = (0.0, 0.0, 40.0)


@dataclass
class Scene:
    name: str
    meshes: list = field(default_factory=list)
    materials: list = field(default_factory=list)
    textures: list = field(default_factory=list)   # list[Image]
    env: Image | None = None
    lights: list = field(default_factory=lambda: [Light()])
    camera: Camera = field(default_factory=Camera)

    @property
    def ntris(self) -> int:
        return sum(m.ntris for m in self.meshes)


# --------------------------------------------------------------------------------------- primitives
def _finish(pos, nrm, uv):
    return (np.ascontiguousarray(pos, dtype=np.float32), np.ascontiguousarray(nrm, dtype=np.float32),
            np.ascontiguousarray(uv, dtype=np.float32))


def rot_z(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)


def rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def transform(prim, R=None, t=(0, 0, 0)):
    pos, nrm, uv = prim
    if R is None:
        R = np.eye(3)
    p = pos.astype(np.float64) @ R.T + np.asarray(t, dtype=np.float64)
    n = nrm.astype(np.float64) @ R.T
    return _finish(p, n, uv)


def merge(prims):
    return _finish(np.concatenate([p[0] for p in prims]), np.concatenate([p[1] for p in prims]),
                   np.concatenate([p[2] for p in prims]))


def box(center, size):
    """Axis-aligned box, 12 triangles, flat normals, each face mapped to [0,1]^2."""
    cx, cy, cz = center
    hx, hy, hz = size[0] / 2, size[1] / 2, size[2] / 2
    faces = []  # (normal, 4 corners CCW seen from outside)
    faces.append(((1, 0, 0), [(hx, -hy, -hz), (hx, hy, -hz), (hx, hy, hz), (hx, -hy, hz)]))
    faces.append(((-1, 0, 0), [(-hx, hy, -hz), (-hx, -hy, -hz), (-hx, -hy, hz), (-hx, hy, hz)]))
    faces.append(((0, 1, 0), [(hx, hy, -hz), (-hx, hy, -hz), (-hx, hy, hz), (hx, hy, hz)]))
    faces.append(((0, -1, 0), [(-hx, -hy, -hz), (hx, -hy, -hz), (hx, -hy, hz), (-hx, -hy, hz)]))
    faces.append(((0, 0, 1), [(-hx, -hy, hz), (hx, -hy, hz), (hx, hy, hz), (-hx, hy, hz)]))
    faces.append(((0, 0, -1), [(-hx, hy, -hz), (hx, hy, -hz), (hx, -hy, -hz), (-hx, -hy, -hz)]))
    quv = [(0, 0), (1, 0), (1, 1), (0, 1)]
    pos, nrm, uv = [], [], []
    for n, c in faces:
        for tri in ((0, 1, 2), (0, 2, 3)):   # quad split 0-1-2 / 0-2-3 as pg1/objloader.cpp:455-471
            pos.append([c[i] for i in tri])
            nrm.append([n] * 3)
            uv.append([quv[i] for i in tri])
    pos = np.array(pos, dtype=np.float64) + np.array([cx, cy, cz])
    return _finish(pos, np.array(nrm), np.array(uv))


def cylinder(base, radius, height, segs=32, cap_top=True, cap_bottom=False, radius_top=None):
    """Cylinder / cone frustum along +z from ``base``; smooth side normals."""
    r1 = radius if radius_top is None else radius_top
    a = np.linspace(0.0, 2 * math.pi, segs + 1)
    c, s = np.cos(a), np.sin(a)
    slope = (radius - r1) / height if height != 0 else 0.0
    nz = slope
    nl = math.sqrt(1 + nz * nz)
    b0 = np.stack([radius * c[:-1], radius * s[:-1], np.zeros(segs)], -1)
    b1 = np.stack([radius * c[1:], radius * s[1:], np.zeros(segs)], -1)
    t0 = np.stack([r1 * c[:-1], r1 * s[:-1], np.full(segs, height)], -1)
    t1 = np.stack([r1 * c[1:], r1 * s[1:], np.full(segs, height)], -1)
    n0 = np.stack([c[:-1], s[:-1], np.full(segs, nz)], -1) / nl
    n1 = np.stack([c[1:], s[1:], np.full(segs, nz)], -1) / nl
    u0, u1 = a[:-1] / (2 * math.pi), a[1:] / (2 * math.pi)
    z0, z1 = np.zeros(segs), np.ones(segs)
    pos = [np.stack([b0, b1, t1], 1), np.stack([b0, t1, t0], 1)]
    nrm = [np.stack([n0, n1, n1], 1), np.stack([n0, n1, n0], 1)]
    uv = [np.stack([np.stack([u0, z0], -1), np.stack([u1, z0], -1), np.stack([u1, z1], -1)], 1),
          np.stack([np.stack([u0, z0], -1), np.stack([u1, z1], -1), np.stack([u0, z1], -1)], 1)]

    def cap(z, r, up):
        ctr = np.tile(np.array([0.0, 0.0, z]), (segs, 1))
        p0 = np.stack([r * c[:-1], r * s[:-1], np.full(segs, z)], -1)
        p1 = np.stack([r * c[1:], r * s[1:], np.full(segs, z)], -1)
        n = np.tile(np.array([0.0, 0.0, 1.0 if up else -1.0]), (segs, 3, 1))
        tri = np.stack([ctr, p0, p1], 1) if up else np.stack([ctr, p1, p0], 1)
        tuv = 0.5 + 0.5 * tri[..., :2] / max(r, 1e-9)
        return tri, n, tuv

    if cap_top and r1 > 0:
        tri, n, tuv = cap(height, r1, True)
        pos.append(tri); nrm.append(n); uv.append(tuv)
    if cap_bottom and radius > 0:
        tri, n, tuv = cap(0.0, radius, False)
        pos.append(tri); nrm.append(n); uv.append(tuv)
    pos = np.concatenate(pos) + np.asarray(base, dtype=np.float64)
    return _finish(pos, np.concatenate(nrm), np.concatenate(uv))


def ellipsoid(center, radii, nu=64, nv=32):
    """Closed UV ellipsoid, smooth normals, CCW outward."""
    th = np.linspace(0.0, 2 * math.pi, nu + 1)
    ph = np.linspace(0.0, math.pi, nv + 1)
    T, P = np.meshgrid(th, ph, indexing="xy")           # [nv+1, nu+1]
    unit = np.stack([np.sin(P) * np.cos(T), np.sin(P) * np.sin(T), np.cos(P)], -1)
    rad = np.asarray(radii, dtype=np.float64)
    pts = unit * rad
    nrm = unit / rad
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    uvs = np.stack([T / (2 * math.pi), 1.0 - P / math.pi], -1)
    i0 = (slice(0, nv), slice(0, nu)); i1 = (slice(0, nv), slice(1, nu + 1))
    j0 = (slice(1, nv + 1), slice(0, nu)); j1 = (slice(1, nv + 1), slice(1, nu + 1))

    def tris(a, b, c):
        return (np.stack([pts[a], pts[b], pts[c]], 2).reshape(-1, 3, 3), np.stack([nrm[a], nrm[b], nrm[c]], 2).reshape(-1, 3, 3),
                np.stack([uvs[a], uvs[b], uvs[c]], 2).reshape(-1, 3, 2))

    A = tris(i0, j0, j1)
    B = tris(i0, j1, i1)
    pos = np.concatenate([A[0], B[0]]); n = np.concatenate([A[1], B[1]]); uv = np.concatenate([A[2], B[2]])
    # drop the degenerate pole triangles
    e1 = pos[:, 1] - pos[:, 0]; e2 = pos[:, 2] - pos[:, 0]
    keep = np.linalg.norm(np.cross(e1, e2), axis=-1) > 1e-9
    pos = pos[keep] + np.asarray(center, dtype=np.float64)
    return _finish(pos, n[keep], uv[keep])


def stud_grid(origin, nx, ny, pitch=5.0, radius=1.5, height=1.7, segs=24):
    """nx*ny LEGO-like studs standing on z = origin[2]."""
    one = cylinder((0, 0, 0), radius, height, segs=segs, cap_top=True)
    ox, oy, oz = origin
    gx, gy = np.meshgrid(np.arange(nx) * pitch, np.arange(ny) * pitch, indexing="ij")
    offs = np.stack([gx.ravel() + ox, gy.ravel() + oy, np.full(gx.size, oz)], -1)
    pos = (one[0][None] + offs[:, None, None, :]).reshape(-1, 3, 3)
    nrm = np.tile(one[1][None], (offs.shape[0], 1, 1, 1)).reshape(-1, 3, 3)
    uv = np.tile(one[2][None], (offs.shape[0], 1, 1, 1)).reshape(-1, 3, 2)
    return _finish(pos, nrm, uv)


# --------------------------------------------------------------------------------------- images
def _value_noise(h, w, cells, rng):
    gy, gx = cells
    g = rng.random((gy + 2, gx + 2))
    y = np.linspace(0, gy, h, endpoint=False); x = np.linspace(0, gx, w, endpoint=False)
    y0 = y.astype(int); x0 = x.astype(int)
    fy = (y - y0)[:, None]; fx = (x - x0)[None, :]
    fy = fy * fy * (3 - 2 * fy); fx = fx * fx * (3 - 2 * fx)
    a = g[y0][:, x0]; b = g[y0][:, x0 + 1]; c = g[y0 + 1][:, x0]; d = g[y0 + 1][:, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def make_envmap(width=4000, height=2000, seed=7) -> Image:
    """Equirectangular stand-in for ``data/spherical_map_lakeside.jpg`` (absent): sky gradient, sun, clouds,
    a horizon band of 'windows' and a textured ground, so bilinear filtering has real texel variation."""
    rng = np.random.default_rng(seed)
    v = (np.arange(height) + 0.5) / height
    u = (np.arange(width) + 0.5) / width
    elev = (0.5 - v) * math.pi                                  # +pi/2 at the top row
    sky = np.stack([0.25 + 0.45 * (1 - np.sin(np.clip(elev, 0, None))), 0.45 + 0.35 * (1 - np.sin(np.clip(elev, 0, None))),
                    np.full(height, 0.92)], -1)                 # [H,3]
    img = np.repeat(sky[:, None, :], width, axis=1)
    clouds = _value_noise(height, width, (12, 24), rng) * 0.6 + _value_noise(height, width, (48, 96), rng) * 0.4
    cm = np.clip((clouds - 0.55) * 3.0, 0, 1)[..., None] * (elev > 0.02)[:, None, None]
    img = img * (1 - cm) + cm * np.array([0.97, 0.97, 0.98])
    # sun
    su, sv = 0.32, 0.27
    du = np.minimum(np.abs(u - su), 1 - np.abs(u - su))[None, :] * 2 * math.pi * np.cos(elev)[:, None]
    dv = (v - sv)[:, None] * math.pi
    d2 = du * du + dv * dv
    img += np.exp(-d2 / 0.0009)[..., None] * np.array([1.0, 0.95, 0.8]) * 1.5 + np.exp(-d2 / 0.03)[..., None] * 0.25
    # ground
    ground = (elev < 0)[:, None]
    gtex = 0.25 + 0.25 * _value_noise(height, width, (40, 160), rng)
    gcol = np.stack([gtex * 0.9, gtex * 1.1 + 0.05, gtex * 0.6], -1)
    img = np.where(ground[..., None], gcol, img)
    # horizon band of lit/unlit windows
    band = (np.abs(elev) < 0.10)[:, None]
    cols = ((u * 160).astype(int) % 2 == 0)[None, :]
    rows = ((v * 200).astype(int) % 2 == 0)[:, None]
    lit = rng.random((1, width // 25 + 1))[:, (np.arange(width) // 25)] > 0.45
    wcol = np.where((cols & rows & lit)[..., None], np.array([0.95, 0.85, 0.45]), np.array([0.18, 0.18, 0.22]))
    tower = (_value_noise(1, width, (1, 60), rng)[0] * 0.10)[None, :]
    bmask = band & (np.abs(elev)[:, None] < tower + 0.01)
    img = np.where(bmask[..., None], wcol, img)
    rgb = (np.clip(img, 0, 1) * 255.0 + 0.5).astype(np.uint8)
    return Image.from_rgb(rgb)


def make_texture(width, height, seed, style="print") -> Image:
    """Stand-ins for ``data/4150p04.jpg`` (640x640) and ``data/3069bp13.jpg`` (640x308): printed LEGO tiles."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:height, 0:width]
    base = np.full((height, width, 3), 0.93)
    base += (_value_noise(height, width, (8, 8), rng)[..., None] - 0.5) * 0.05
    if style == "print":
        cx, cy = width / 2, height / 2
        r = np.hypot(x - cx, y - cy) / (min(width, height) / 2)
        ring = (np.abs(r - 0.7) < 0.06) | (np.abs(r - 0.45) < 0.03)
        base[ring] = (0.1, 0.1, 0.12)
        wedge = (np.arctan2(y - cy, x - cx) % (math.pi / 4) < math.pi / 8) & (r < 0.42)
        base[wedge] = (0.8, 0.12, 0.1)
        base[(r < 0.12)] = (0.95, 0.8, 0.1)
    else:
        stripes = ((x // max(width // 16, 1)) % 2 == 0) & (y > height * 0.2) & (y < height * 0.8)
        base[stripes] = (0.1, 0.25, 0.7)
        for _ in range(12):
            bx, by = rng.integers(0, width - 40), rng.integers(0, max(height - 20, 1))
            base[by:by + 12, bx:bx + 36] = (0.05, 0.05, 0.05)
    rgb = (np.clip(base, 0, 1) * 255.0 + 0.5).astype(np.uint8)
    return Image.from_rgb(rgb)


# --------------------------------------------------------------------------------------- materials
def avenger_materials() -> list:
    """The five materials of ``data/6887_allied_avenger.mtl`` as ``LoadMTL`` parses them
    (pg1/objloader.cpp:53-208): Ks reads as (1.0, 0.8, 0.8) because ``Ks 1.0. 1.0 1.0`` stops sscanf after
    the first float and the ctor default 0.8 survives (pg1/material.cpp:14)."""
    ks = (1.0, 0.8, 0.8)
    return [
        Material("black_plastic", (0.3, 0.3, 0.3), ks, 32.0, 1.460, 3, -1, (0.03, 0.03, 0.03)),
        Material("green_plastic_transparent", (0.85, 1.0, 0.01), ks, 32.0, 1.5, 4, -1, (0.03, 0.03, 0.03)),
        Material("white_plastic", (0.95, 0.95, 0.95), ks, 32.0, 1.460, 3, -1, (0.03, 0.03, 0.03)),
        Material("white_plastic_4150p04", (0.95, 0.95, 0.95), ks, 32.0, 1.460, 3, 0, (0.03, 0.03, 0.03), "4150p04.jpg"),
        Material("white_plastic_3069bp13", (0.95, 0.95, 0.95), ks, 32.0, 1.460, 3, 1, (0.03, 0.03, 0.03), "3069bp13.jpg"),
    ]


AVENGER_MTL_TEXT = """# stand-in material library: same five materials, same field values and the same malformed Ks line
# as data/6887_allied_avenger.mtl of the reference (values restated, file regenerated)

newmtl black_plastic
\tNs 32
\td 1
\tTr 0
\tTf 1 1 1
\tillum 2
\tKa 0.03 0.03 0.03
\tKd 0.3 0.3 0.3
\tKs 1.0. 1.0 1.0
\tshader 3
\tNi 1.460

newmtl green_plastic_transparent
\tNs 32
\td 1
\tTr 0
\tTf 0.4 0.001 0.4
\tillum 2
\tKa 0.03 0.03 0.03
\tKd 0.85 1.0 0.01
\tKs 1.0. 1.0 1.0
\tshader 4
\tNi 1.5

newmtl white_plastic
\tNs 32
\td 1
\tTr 0
\tTf 1 1 1
\tillum 2
\tKa 0.03 0.03 0.03
\tKd 0.95 0.95 0.95
\tKs 1.0. 1.0 1.0
\tshader 3
\tNi 1.460

newmtl white_plastic_4150p04
\tNs 32
\td 1
\tTr 0
\tTf 1 1 1
\tillum 2
\tKa 0.03 0.03 0.03
\tKd 0.95 0.95 0.95
\tKs 1.0. 1.0 1.0
  \tmap_Kd 4150p04.jpg
\tshader 3
\tNi 1.460

newmtl white_plastic_3069bp13
\tNs 32
\td 1
\tTr 0
\tTf 1 1 1
\tillum 2
\tKa 0.03 0.03 0.03
\tKd 0.95 0.95 0.95
\tKs 1.0. 1.0 1.0
  \tmap_Kd 3069bp13.jpg
\tshader 3
\tNi 1.460
"""


# --------------------------------------------------------------------------------------- scenes
def avenger_proxy(seed: int = 6887, detail: float = 1.0, env_size=(4000, 2000), with_images: bool = True) -> Scene:
    """LEGO-like stand-in for ``6887_allied_avenger.obj`` (absent): 94 surfaces (the count the reference's
    screenshot shows), about 2e5 triangles at detail=1, inside ~[-100,100]^2 x [0,80],
    with one closed dielectric shell (the canopy, ``green_plastic_transparent``) and two textured tiles."""
    rng = np.random.default_rng(seed)
    BLACK, GREEN, WHITE, TEX0, TEX1 = 0, 1, 2, 3, 4
    segs = max(8, int(round(24 * detail)))
    meshes: list[Mesh] = []

    def add(name, prim, mat):
        meshes.append(Mesh(f"{name}_{len(meshes):02d}", prim[0], prim[1], prim[2], mat))

    def studs(n):
        return max(1, int(round(n * math.sqrt(detail))))

    # hull: three decks with stud grids
    add("deck", box((0, 0, 30), (130, 50, 4)), BLACK)
    add("deck_studs", stud_grid((-62.5, -22.5, 32), studs(26), studs(10), pitch=125 / max(studs(26) - 1, 1), segs=segs), WHITE)
    add("keel", box((0, 0, 24), (150, 22, 8)), WHITE)
    add("upper", box((-25, 0, 36), (60, 36, 8)), WHITE)
    add("upper_studs", stud_grid((-52.5, -15, 40), studs(12), studs(7), pitch=5.0, segs=segs), BLACK)
    # wings
    for side, sgn in (("l", 1), ("r", -1)):
        add(f"wing_{side}", box((-10, sgn * 60, 28), (60, 70, 3)), WHITE)
        add(f"wing_{side}_studs", stud_grid((-37.5, sgn * 60 - 32.5, 29.5), studs(16), studs(14), pitch=5.0, segs=segs), WHITE)
        add(f"wing_{side}_tip", box((-10, sgn * 97, 30), (40, 4, 10)), BLACK)
        add(f"wing_{side}_slope", transform(box((0, 0, 0), (30, 20, 3)), rot_x(sgn * 0.35), (-35, sgn * 38, 36)), BLACK)
        # engines (cylinders along x)
        eng = transform(cylinder((0, 0, 0), 8, 46, segs=2 * segs, cap_top=True, cap_bottom=True), rot_y(math.pi / 2), (-48, sgn * 36, 22))
        add(f"engine_{side}", eng, WHITE)
        noz = transform(cylinder((0, 0, 0), 6, 8, segs=2 * segs, cap_top=True, radius_top=9), rot_y(-math.pi / 2), (-48, sgn * 36, 22))
        add(f"nozzle_{side}", noz, BLACK)
        add(f"gun_{side}", transform(cylinder((0, 0, 0), 1.6, 50, segs=segs, cap_top=True), rot_y(math.pi / 2), (20, sgn * 80, 27)), BLACK)
    # canopy: one closed dielectric shell + interior
    nu = max(16, int(round(96 * math.sqrt(detail)))); nv = max(8, int(round(48 * math.sqrt(detail))))
    add("canopy", ellipsoid((30, 0, 44), (26, 15, 12), nu=nu, nv=nv), GREEN)
    add("seat", box((26, 0, 40), (10, 10, 8)), BLACK)
    add("pilot", cylinder((28, 0, 44), 3, 6, segs=segs, cap_top=True), WHITE)
    add("console", transform(box((0, 0, 0), (6, 14, 4)), rot_y(-0.5), (40, 0, 41)), WHITE)
    # textured tiles (uv outside [0,1] on the second, as tiled model UVs are)
    t0 = box((-25, 0, 40.6), (24, 24, 1.2))
    add("tile_4150p04", t0, TEX0)
    t1 = transform(box((0, 0, 0), (40, 19, 1.2)), rot_y(0.45), (70, 0, 36))
    t1 = (t1[0], t1[1], (t1[2] * np.float32(1.5) - np.float32(0.25)).astype(np.float32))
    add("tile_3069bp13", t1, TEX1)
    # nose
    add("nose", transform(cylinder((0, 0, 0), 11, 40, segs=2 * segs, cap_top=True, radius_top=2.5), rot_y(math.pi / 2), (65, 0, 27)), WHITE)
    add("tail_fin", transform(box((0, 0, 0), (30, 3, 26)), rot_y(0.4), (-70, 0, 48)), BLACK)
    add("tail_studs", stud_grid((-74, -10, 28), studs(4), studs(5), pitch=5.0, segs=segs), BLACK)
    # landing gear + ground plate so shadows/reflections have something to hit
    for k, (gx, gy) in enumerate(((45, 0), (-40, 30), (-40, -30))):
        add(f"gear_{k}", cylinder((gx, gy, 2), 2.5, 20, segs=segs, cap_top=False), BLACK)
        add(f"wheel_{k}", transform(cylinder((0, 0, 0), 5, 4, segs=2 * segs, cap_top=True, cap_bottom=True), rot_x(math.pi / 2), (gx, gy + 2, 5)), BLACK)
    add("pad", box((0, 0, -1), (190, 190, 2)), WHITE)
    add("pad_studs", stud_grid((-90, -90, 0), studs(37), studs(37), pitch=180 / max(studs(37) - 1, 1), segs=segs), WHITE)
    # greebles up to exactly 94 surfaces
    k = 0
    while len(meshes) < 94:
        gx = float(rng.uniform(-60, 55)); gy = float(rng.uniform(-22, 22)); gz = 32.0
        if abs(gx - 30) < 30 and abs(gy) < 17:
            gz = 24.0; gy = float(np.sign(gy) * (18 + abs(gy) * 0.2)) if gy != 0 else 19.0
        kind = k % 3
        if kind == 0:
            s = rng.uniform(2, 7, 3)
            prim = transform(box((0, 0, 0), tuple(s)), rot_z(float(rng.uniform(0, math.pi))), (gx, gy, gz + s[2] / 2))
        elif kind == 1:
            prim = cylinder((gx, gy, gz), float(rng.uniform(1, 3)), float(rng.uniform(2, 9)), segs=segs, cap_top=True)
        else:
            prim = merge([box((gx, gy, gz + 1), (5, 5, 2)), stud_grid((gx, gy, gz + 2), 1, 1, segs=segs)])
        add(f"greeble{k}", prim, int(rng.choice([BLACK, WHITE, WHITE])))
        k += 1
    sc = Scene("avenger_proxy(stand-in)", meshes, avenger_materials())
    if with_images:
        sc.textures = [make_texture(640, 640, seed + 1, "print"), make_texture(640, 308, seed + 2, "stripes")]
        sc.env = make_envmap(env_size[0], env_size[1], seed + 3)
    return sc


def triangle_soup(n: int, seed: int = 1, resolution=(3840, 2160)) -> Scene:
    """Config C5: centroids ~ U([-100,100]^3), two edge vectors ~ U([-1,1]^3),
    per-vertex normals = geometric normal, uv ~ U[0,1)^2, one Lambert material Kd=0.8, Philox counter RNG."""
    rng = np.random.Generator(np.random.Philox(seed))
    c = rng.uniform(-100, 100, (n, 3)).astype(np.float32)
    e1 = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    e2 = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    v0 = c - (e1 + e2) / np.float32(3)
    pos = np.stack([v0, v0 + e1, v0 + e2], 1)
    ng = np.cross(e1, e2)
    ln = np.linalg.norm(ng, axis=-1, keepdims=True)
    ng = np.where(ln > 0, ng / np.maximum(ln, 1e-30), np.array([0, 0, 1], dtype=np.float32)).astype(np.float32)
    nrm = np.repeat(ng[:, None, :], 3, axis=1)
    uv = rng.random((n, 3, 2), dtype=np.float32)
    mat = Material("soup_lambert", (0.8, 0.8, 0.8), (0.0, 0.0, 0.0), 1.0, 1.5, 3, -1)
    cam = Camera(resolution[0], resolution[1], Camera().fov_y, (-210.0, -262.5, 120.0), (0.0, 0.0, 60.0))
    sc = Scene(f"soup{n}", [Mesh("soup", *_finish(pos, nrm, uv), 0)], [mat], camera=cam)
    return sc


def palm_grove(n_palms: int = 300, seed: int = 11, fronds: int = 28, leaflets: int = 44, resolution=(1920, 1080),
               env_size=(4000, 2000), with_images: bool = True) -> Scene:
    """Config C4 stand-in ("PalmTrees"; the reference holds no such mesh, only a cube map): a seeded grove
    of palms with thin leaflet triangles (stresses SAH quality).  ~ n_palms * (fronds*leaflets*4 + trunk)."""
    rng = np.random.default_rng(seed)
    trunk_parts, leaf_parts = [], []
    for _ in range(n_palms):
        px, py = rng.uniform(-160, 160, 2)
        hgt = rng.uniform(28, 55)
        lean = rng.uniform(-0.15, 0.15, 2)
        nseg = 10
        for s in range(nseg):
            z0 = hgt * s / nseg
            r0 = 2.2 * (1 - 0.5 * s / nseg); r1 = 2.2 * (1 - 0.5 * (s + 1) / nseg)
            seg = cylinder((px + lean[0] * z0, py + lean[1] * z0, z0), r0, hgt / nseg, segs=10, cap_top=(s == nseg - 1), radius_top=r1)
            trunk_parts.append(seg)
        top = np.array([px + lean[0] * hgt, py + lean[1] * hgt, hgt])
        for f in range(fronds):
            az = 2 * math.pi * f / fronds + rng.uniform(-0.1, 0.1)
            droop = rng.uniform(0.6, 1.4)
            L = rng.uniform(16, 26)
            s = np.linspace(0.03, 1.0, leaflets + 1)
            # rachis curve: out along az, rising then drooping
            rad = L * s
            z = 6 * s - droop * 9 * s * s
            cx = top[0] + rad * math.cos(az); cy = top[1] + rad * math.sin(az); cz = top[2] + z
            ctr = np.stack([cx, cy, cz], -1)
            tang = np.gradient(ctr, axis=0); tang /= np.linalg.norm(tang, axis=-1, keepdims=True)
            side = np.cross(tang, np.array([0, 0, 1.0])); side /= np.linalg.norm(side, axis=-1, keepdims=True)
            up = np.cross(side, tang)
            wlen = 4.0 * np.sin(np.pi * s) ** 0.7 + 0.3
            for sgn in (1.0, -1.0):
                a = ctr[:-1]; b = ctr[1:]
                tip = 0.5 * (a + b) + sgn * side[:-1] * wlen[:-1, None] - up[:-1] * (0.35 * wlen[:-1, None]) + tang[:-1] * 0.8
                tri = np.stack([a, b, tip], 1) if sgn > 0 else np.stack([b, a, tip], 1)
                n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]); n /= np.maximum(np.linalg.norm(n, axis=-1, keepdims=True), 1e-12)
                uv = np.tile(np.array([[0, 0], [1, 0], [0.5, 1]], dtype=np.float32), (tri.shape[0], 1, 1))
                leaf_parts.append(_finish(tri, np.repeat(n[:, None, :], 3, 1), uv))
                # second triangle gives each leaflet some width along the rachis
                tip2 = tip + tang[:-1] * 1.2
                tri2 = np.stack([b, tip2, tip], 1) if sgn > 0 else np.stack([tip2, b, tip], 1)
                n2 = np.cross(tri2[:, 1] - tri2[:, 0], tri2[:, 2] - tri2[:, 0]); n2 /= np.maximum(np.linalg.norm(n2, axis=-1, keepdims=True), 1e-12)
                leaf_parts.append(_finish(tri2, np.repeat(n2[:, None, :], 3, 1), uv))
    mats = [Material("trunk", (0.45, 0.33, 0.2), (0, 0, 0), 1.0, 1.5, 3), Material("leaf", (0.2, 0.55, 0.15), (0, 0, 0), 1.0, 1.5, 3),
            Material("sand", (0.8, 0.72, 0.5), (0, 0, 0), 1.0, 1.5, 3)]
    meshes = [Mesh("trunks", *merge(trunk_parts), 0), Mesh("leaves", *merge(leaf_parts), 1), Mesh("ground", *box((0, 0, -1), (420, 420, 2)), 2)]
    cam = Camera(resolution[0], resolution[1], Camera().fov_y, (-230.0, -260.0, 70.0), (0.0, 0.0, 30.0))
    sc = Scene("palm_grove(stand-in)", meshes, mats, camera=cam)
    if with_images:
        sc.env = make_envmap(env_size[0], env_size[1], seed + 3)
    return sc


def single_triangle() -> Scene:
    """tutorial_1's scene (pg1/tutorials.cpp:39-41,61-69)."""
    pos = np.array([[[0, 0, 0], [2, 0, 0], [0, 3, 0]]], dtype=np.float32)
    nrm = np.array([[[0, 0, 1]] * 3], dtype=np.float32)
    uv = np.array([[[0, 1], [1, 1], [0, 0]]], dtype=np.float32)
    return Scene("tutorial_1", [Mesh("tri", pos, nrm, uv, 0)], [Material("m", type=3)])


def cornell_like(seed=3, n_boxes=6, env_size=(512, 256)) -> Scene:
    """Small deterministic test scene (a few hundred triangles) exercising every branch of ``trace``:
    Phong + texture, dielectric sphere and slab (TIR), env-map misses, shadowed and unshadowed hits."""
    rng = np.random.default_rng(seed)
    meshes = [Mesh("floor", *box((0, 0, -1), (160, 160, 2)), 2)]
    meshes.append(Mesh("glass_ball", *ellipsoid((10, -10, 30), (18, 18, 18), nu=24, nv=12), 1))
    meshes.append(Mesh("glass_slab", *transform(box((0, 0, 0), (30, 4, 30)), rot_z(0.5), (-35, -20, 22)), 1))
    meshes.append(Mesh("tile", *transform(box((0, 0, 0), (40, 40, 2)), rot_x(0.6), (30, 45, 30)), 3))
    meshes.append(Mesh("tile2", *transform(box((0, 0, 0), (30, 15, 2)), rot_y(-0.4), (-40, 40, 25)), 4))
    for i in range(n_boxes):
        c = rng.uniform(-60, 60, 3); c[2] = rng.uniform(5, 50)
        meshes.append(Mesh(f"b{i}", *transform(box((0, 0, 0), tuple(rng.uniform(6, 18, 3))), rot_z(float(rng.uniform(0, 3))), tuple(c)), int(i % 2) * 2))
    meshes.append(Mesh("pillar", *cylinder((-10, 30, 0), 6, 50, segs=16, cap_top=True), 0))
    sc = Scene("cornell_like", meshes, avenger_materials())
    sc.textures = [make_texture(64, 64, seed + 1, "print"), make_texture(61, 31, seed + 2, "stripes")]
    sc.env = make_envmap(env_size[0], env_size[1], seed + 3)
    return sc


# --------------------------------------------------------------------------------------- OBJ / MTL writer
def write_obj(scene: Scene, obj_path: str, mtl_name: str | None = None, mtl_text: str | None = None) -> None:
    """Write the scene as Wavefront OBJ in the dialect ``LoadOBJ`` accepts (pg1/objloader.cpp:210-507):
    ``v``/``vt``/``vn`` triples for every corner, one ``g`` per surface, ``usemtl`` after ``g``, faces as
    ``f v/vt/vn v/vt/vn v/vt/vn``.  Floats are printed with repr() so float32 values round-trip."""
    d = os.path.dirname(obj_path)
    if d:
        os.makedirs(d, exist_ok=True)
    mtl_name = mtl_name or (os.path.splitext(os.path.basename(obj_path))[0] + ".mtl")
    with open(obj_path, "w") as f:
        f.write(f"# stand-in scene {scene.name}\nmtllib {mtl_name}\n")
        base = 1
        for m in scene.meshes:
            n = m.ntris * 3
            p = m.pos.reshape(-1, 3); nn = m.nrm.reshape(-1, 3); t = m.uv.reshape(-1, 2)
            f.write("".join(f"v {float(a)!r} {float(b)!r} {float(c)!r}\n" for a, b, c in p))
            f.write("".join(f"vn {float(a)!r} {float(b)!r} {float(c)!r}\n" for a, b, c in nn))
            f.write("".join(f"vt {float(a)!r} {float(b)!r} 0\n" for a, b in t))
            f.write(f"g {m.name}\nusemtl {scene.materials[m.material].name}\n")
            idx = np.arange(base, base + n).reshape(-1, 3)
            f.write("".join(f"f {a}/{a}/{a} {b}/{b}/{b} {c}/{c}/{c}\n" for a, b, c in idx))
            base += n
    if mtl_text is None:
        lines = []
        for mt in scene.materials:
            lines += [f"newmtl {mt.name}", f"\tNs {mt.shininess}", f"\tKa {mt.ambient[0]} {mt.ambient[1]} {mt.ambient[2]}",
                      f"\tKd {mt.diffuse[0]} {mt.diffuse[1]} {mt.diffuse[2]}", f"\tKs {mt.specular[0]} {mt.specular[1]} {mt.specular[2]}"]
            if mt.map_kd:
                lines.append(f"\tmap_Kd {mt.map_kd}")
            lines += [f"\tshader {mt.type}", f"\tNi {mt.ior}", ""]
        mtl_text = "\n".join(lines)
    with open(os.path.join(d, mtl_name), "w") as f:
        f.write(mtl_text)
