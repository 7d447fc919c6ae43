"""Regenerates tests/golden/*.  Run in the build container (needs /root/reference for the two PNG fixtures):

    python tests/golden/make_golden.py

* test3_bgr.npy / test4_bgra.npy : the reference's own texel fixtures (data/test3.png, data/test4.png) decoded to the
  top-down BGR(A) byte layout Texture::Texture produces (pg1/texture.cpp:36-47).  tutorial_2 (pg1/tutorials.cpp:170-178)
  reads test4.png; its printout is the T2 known answer.
* avenger_mtl_parse.json : what LoadMTL (pg1/objloader.cpp:53-208) makes of data/6887_allied_avenger.mtl, obtained by
  running the same sscanf formats through glibc (ctypes) on the reference's file.
* cornell_*.npy : oracle renders of the small seeded test scene, so that any later change of the oracle is caught and
  the CUDA path has a committed target that does not need the oracle at run time.
"""
import ctypes
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/data"


def decode(path):
    from PIL import Image
    im = Image.open(path)
    a = np.asarray(im)
    bpp = a.shape[2]
    out = a.copy()
    out[..., 0], out[..., 2] = a[..., 2], a[..., 0]      # RGB(A) -> BGR(A)
    h, w = a.shape[:2]
    pitch = (w * bpp + 3) // 4 * 4
    buf = np.zeros((h, pitch), np.uint8)
    buf[:, :w * bpp] = out.reshape(h, w * bpp)
    return buf


def parse_mtl_like_reference(path):
    """LoadMTL's parsing rules restated with the real sscanf (pg1/objloader.cpp:106-191)."""
    libc = ctypes.CDLL("libc.so.6")
    mats = []
    cur = None
    for raw in open(path, "rb").read().split(b"\n"):
        if not raw or raw[:1] == b"#":
            continue
        if raw.startswith(b"newmtl"):
            name = ctypes.create_string_buffer(128)
            libc.sscanf(raw, b"%*s %s", name)
            # Material::Material() defaults, pg1/material.cpp:9-26 (type is left uninitialised there)
            cur = dict(name=name.value.decode(), ambient=[0.1] * 3, diffuse=[0.4] * 3, specular=[0.8] * 3, emission=[0.0] * 3,
                       shininess=1.0, ior=1.5, type=None, map_Kd=None)
            mats.append(cur)
            continue
        tmp = raw.strip(b" \t\r")
        f3 = (ctypes.c_float * 3)
        def scan3(key):
            v = f3(*[np.float32(x) for x in cur[key]])
            libc.sscanf(tmp, b"%*s %f %f %f", ctypes.byref(v, 0), ctypes.byref(v, 4), ctypes.byref(v, 8))
            cur[key] = [float(np.float32(x)) for x in v]
        if tmp.startswith(b"Ka"): scan3("ambient")
        if tmp.startswith(b"shader"):
            i = ctypes.c_int(); libc.sscanf(tmp, b"%*s %d", ctypes.byref(i)); cur["type"] = i.value
        if tmp.startswith(b"Ni"):
            f = ctypes.c_float(); libc.sscanf(tmp, b"%*s %f", ctypes.byref(f)); cur["ior"] = float(f.value)
        if tmp.startswith(b"Kd"): scan3("diffuse")
        if tmp.startswith(b"Ks"): scan3("specular")
        if tmp.startswith(b"Ke"): scan3("emission")
        if tmp.startswith(b"Ns"):
            f = ctypes.c_float(); libc.sscanf(tmp, b"%*s %f", ctypes.byref(f)); cur["shininess"] = float(f.value)
        if tmp.startswith(b"map_Kd"):
            s = ctypes.create_string_buffer(256); libc.sscanf(tmp, b"%*s %s", s); cur["map_Kd"] = s.value.decode()
    return mats


def main():
    np.save(os.path.join(HERE, "test3_bgr.npy"), decode(os.path.join(REF, "test3.png")))
    np.save(os.path.join(HERE, "test4_bgra.npy"), decode(os.path.join(REF, "test4.png")))
    with open(os.path.join(HERE, "avenger_mtl_parse.json"), "w") as f:
        json.dump(parse_mtl_like_reference(os.path.join(REF, "6887_allied_avenger.mtl")), f, indent=1)
    from pgi_raytracing_b200 import scenes
    from oracle.oracle import Oracle, make_params
    sc = scenes.cornell_like()
    sc.camera = scenes.Camera(64, 48, sc.camera.fov_y, sc.camera.view_from, sc.camera.view_at)
    orc = Oracle(sc)
    for tag, p in (("c1", dict(sampling_width=1, jitter=0, aperture=0.0)), ("dof3x3", dict(seed=5)),
                   ("lambert", dict(sampling_width=1, jitter=0, aperture=0.0, shader_mode=1)),
                   ("pinhole_depth3", dict(sampling_width=2, jitter=1, camera_mode=1, max_depth=3, gamma_level=0.4, seed=9))):
        rgba, g, pr, st = orc.render(make_params(**p), brute=True)
        np.savez_compressed(os.path.join(HERE, f"cornell_{tag}.npz"), rgba=rgba, geom=g, prim=pr,
                            rays=np.array([st["primary"], st["shadow"], st["reflection"], st["refraction"]], np.uint64),
                            params=json.dumps(p))
    print("golden written")


if __name__ == "__main__":
    main()
