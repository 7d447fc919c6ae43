"""The C++ host surface (pgi_raytracing_b200/host/): LoadOBJ / LoadMTL, Texture + image decode, PinHoleCamera, and --
on the GPU -- Raytracer over the C ABI.  Driven through the flat C test API of host/pg1_capi.cpp.

Reference behaviour cited: pg1/objloader.cpp:53-507 (loader), pg1/texture.cpp:5-54 (decode to top-down BGR, 4-byte pitch),
pg1/PinHoleCamera.cpp:5-105, pg1/raytracer.cpp:48-128 (LoadScene).
"""
import ctypes as C
import io
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from pgi_raytracing_b200 import scenes

PKG = os.path.join(ROOT, "pgi_raytracing_b200")
HOST_LIB = os.path.join(PKG, "libpg1_host.so")


@pytest.fixture(scope="module")
def host():
    subprocess.check_call(["make", "-C", os.path.join(PKG, "csrc"), "-s"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(PKG, "host"), "-s"], stdout=subprocess.DEVNULL)
    lib = C.CDLL(HOST_LIB)
    VP = C.c_void_p
    for name, res, args in [
        ("pg1_last_error", C.c_char_p, []), ("pg1_load_obj", VP, [C.c_char_p, C.c_int]), ("pg1_load_mtl", VP, [C.c_char_p, C.c_char_p]),
        ("pg1_free_scene", None, [VP]), ("pg1_scene_rc", C.c_int, [VP]), ("pg1_num_surfaces", C.c_int, [VP]), ("pg1_num_materials", C.c_int, [VP]),
        ("pg1_surface_triangles", C.c_int, [VP, C.c_int]), ("pg1_surface_name", C.c_char_p, [VP, C.c_int]), ("pg1_surface_material", C.c_int, [VP, C.c_int]),
        ("pg1_surface_data", None, [VP, C.c_int, VP, VP, VP]), ("pg1_material", C.c_char_p, [VP, C.c_int, VP]),
        ("pg1_load_image", VP, [C.c_char_p]), ("pg1_image_info", None, [VP, VP]), ("pg1_image_bytes", None, [VP, VP]), ("pg1_free_image", None, [VP]),
        ("pg1_write_ppm", C.c_int, [C.c_char_p, VP, C.c_int, C.c_int]), ("pg1_write_image", C.c_int, [C.c_char_p, VP, C.c_int, C.c_int]),
        ("pg1_camera_ray", None, [C.c_int, C.c_int, C.c_float, VP, VP, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, VP]),
        ("pg1_raytracer_create", VP, [C.c_int, C.c_int, C.c_float, VP, VP, C.c_char_p]), ("pg1_raytracer_destroy", None, [VP]),
        ("pg1_raytracer_load_scene", C.c_int, [VP, C.c_char_p, C.c_char_p]),
        ("pg1_raytracer_set", None, [VP, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float, C.c_uint]),
        ("pg1_raytracer_render", C.c_int, [VP, VP, VP]), ("pg1_raytracer_get_pixel", C.c_int, [VP, C.c_int, C.c_int, VP]),
        ("pg1_raytracer_counts", C.c_int, [VP, VP]),
    ]:
        fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
    return lib


def load_image(host, path):
    h = host.pg1_load_image(path.encode())
    assert h, host.pg1_last_error()
    info = (C.c_int * 4)(); host.pg1_image_info(h, info)
    w, hh, pitch, bpp = list(info)
    buf = np.zeros((hh, pitch), np.uint8); host.pg1_image_bytes(h, buf.ctypes.data)
    host.pg1_free_image(h)
    return buf, w, hh, pitch, bpp


def scene_arrays(host, h, i):
    n = host.pg1_surface_triangles(h, i)
    pos = np.zeros((n, 9), np.float32); nrm = np.zeros((n, 9), np.float32); uv = np.zeros((n, 6), np.float32)
    host.pg1_surface_data(h, i, pos.ctypes.data, nrm.ctypes.data, uv.ctypes.data)
    return pos, nrm, uv


def material(host, h, i):
    out = np.zeros(16, np.float32)
    name = host.pg1_material(h, i, out.ctypes.data).decode()
    return name, out


# ------------------------------------------------------------------------------------------------ image decode
def _test_picture(w, h, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 120 * np.sin(x / 17.0 + seed), 127 + 120 * np.cos(y / 11.0), (x * 3 + y * 5) % 256], -1)
    img += rng.normal(0, 12, img.shape)
    img[h // 3:h // 3 + 9, :, :] = 255; img[:, w // 2:w // 2 + 3, :] = 0          # hard edges
    return np.clip(img, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("size,subsampling,quality,restart", [((64, 48), 2, 90, 0), ((67, 45), 2, 75, 0), ((640, 308), 2, 85, 0),
                                                              ((33, 17), 0, 95, 0), ((200, 100), 0, 80, 5), ((129, 65), 2, 60, 3)])
def test_jpeg_decoder_matches_libjpeg(host, tmp_path, size, subsampling, quality, restart):
    """Baseline JPEG -> top-down BGR rows with a 4-byte pitch, byte for byte what libjpeg (inside FreeImage / Pillow)
    decodes: islow IDCT, fancy 4:2:0 up-sampling, fixed-point YCbCr."""
    from PIL import Image
    w, h = size
    p = str(tmp_path / "t.jpg")
    kw = dict(restart_marker_blocks=restart) if restart else {}
    Image.fromarray(_test_picture(w, h, w + h)).save(p, "JPEG", quality=quality, subsampling=subsampling, **kw)
    ref = np.asarray(Image.open(p).convert("RGB"))
    buf, ww, hh, pitch, bpp = load_image(host, p)
    assert (ww, hh, bpp, pitch) == (w, h, 3, (3 * w + 3) // 4 * 4)
    got = buf[:, :3 * w].reshape(h, w, 3)[..., ::-1]
    assert np.array_equal(got, ref)


def test_reference_textures_decode_like_libjpeg(host):
    """The reference's own JPEGs, when the mount is present (this container only)."""
    from PIL import Image
    d = "/root/reference/data"
    if not os.path.isdir(d):
        pytest.skip("reference data not mounted")
    for name in ("4150p04.jpg", "3069bp13.jpg", "spherical_map_windows.jpg", "PalmTrees/posx.jpg"):
        ref = np.asarray(Image.open(os.path.join(d, name)).convert("RGB"))
        buf, w, h, pitch, bpp = load_image(host, os.path.join(d, name))
        got = buf[:, :3 * w].reshape(h, w, 3)[..., ::-1]
        assert np.array_equal(got, ref), name


def test_ppm_and_bmp(host, tmp_path):
    from PIL import Image
    img = _test_picture(37, 21, 1)
    for ext, fmt in (("ppm", "PPM"), ("bmp", "BMP")):
        p = str(tmp_path / f"t.{ext}"); Image.fromarray(img).save(p, fmt)
        buf, w, h, pitch, bpp = load_image(host, p)
        assert np.array_equal(buf[:, :3 * w].reshape(h, w, 3)[..., ::-1], img)
        assert np.array_equal(buf, scenes.Image.from_rgb(img).data)          # same bytes the Python harness hands to the ABI
    assert host.pg1_load_image(str(tmp_path / "missing.jpg").encode()) is None


# ------------------------------------------------------------------------------------------------ loader
def test_load_obj_round_trip(host, tmp_path):
    """write_obj -> LoadOBJ: surfaces in group order, un-indexed corners bit-exact, materials by last usemtl."""
    sc = scenes.cornell_like()
    path = str(tmp_path / "scene.obj")
    scenes.write_obj(sc, path)
    h = host.pg1_load_obj(path.encode(), 0)
    try:
        assert host.pg1_scene_rc(h) == len(sc.meshes) == host.pg1_num_surfaces(h)
        assert host.pg1_num_materials(h) == len(sc.materials)
        for i, m in enumerate(sc.meshes):
            pos, nrm, uv = scene_arrays(host, h, i)
            assert host.pg1_surface_name(h, i).decode() == m.name
            assert host.pg1_surface_material(h, i) == m.material
            assert np.array_equal(pos, m.pos.reshape(-1, 9)) and np.array_equal(uv, m.uv.reshape(-1, 6))
            n = m.nrm.reshape(-1, 3).astype(np.float32)                       # normals are re-normalised on load (objloader.cpp:323)
            sq = (n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1] + n[:, 2] * n[:, 2]).astype(np.float32)
            rn = (np.float32(1) / np.sqrt(sq, dtype=np.float32)).astype(np.float32)
            assert np.array_equal(nrm.reshape(-1, 3), (n * rn[:, None]).astype(np.float32))
    finally:
        host.pg1_free_scene(h)


def test_mtl_parse_matches_the_reference_fixture(host, tmp_path):
    """tests/golden/avenger_mtl_parse.json was made with glibc sscanf on the reference's data/6887_allied_avenger.mtl
    (prefix key matching, partial parses keep Material() defaults: Ks 1.0. 1.0 1.0 -> (1.0, 0.8, 0.8))."""
    p = str(tmp_path / "a.mtl")
    open(p, "w").write(scenes.AVENGER_MTL_TEXT)
    h = host.pg1_load_mtl(p.encode(), (str(tmp_path) + "/").encode())
    try:
        with open(os.path.join(GOLDEN, "avenger_mtl_parse.json")) as f:
            gold = json.load(f)
        assert host.pg1_num_materials(h) == len(gold) == 5
        for i, g in enumerate(gold):
            name, v = material(host, h, i)
            assert name == g["name"]
            want = np.array(g["ambient"] + g["diffuse"] + g["specular"] + g["emission"] + [g["shininess"], g["ior"], g["type"]], np.float32)
            assert np.array_equal(v[:15], want), name
    finally:
        host.pg1_free_scene(h)


def test_loader_quirks(host, tmp_path):
    obj = tmp_path / "q.obj"
    (tmp_path / "q.mtl").write_text("newmtl a\nKd 1 0 0\nshader 4e1\nnewmtl b\n  Kd 0 1 0\nnewmtl a\nKd 0 0 1\nKdx 9 9 9\n# Kd 5 5 5\nnewmtl c\nNs 7\n  newmtl d\nshader x\nshader 3.9\n")
    obj.write_text("\n".join([
        "mtllib q.mtl", "v 0 0 0", "v 1 0 0", "v 1 1 0", "v 0 1 0", "vn 0 0 2", "vt 0.25 0.75 0", "vt 1 1",
        "f 1/1/1 2/1/1 3/2/1",                    # before any g: surface named ""
        "g first", "usemtl b", "f 1/1/1 2/1/1 3/1/1 4/2/1",   # quad -> (0,1,2) (0,2,3)
        "usemtl a",                               # the LAST usemtl before the flush wins
        "g renamed_without_faces", "g second", "usemtl nosuch",
        "f 1/1/1 2/1/1 3/1/1", "f 1//1 2//1 3//1",            # missing vt: skipped
        "f 1/1/1 2/1/1 9/1/1",                                # index out of range: skipped
        "f  1/1/1 2/1/1 3/1/1",                               # double space = 4 spaces with 3 corners: skipped (undefined in the reference)
        "", "g third", "usemtl c", "f 3/2/1 2/1/1 1/1/1  "]) + "\n")
    h = host.pg1_load_obj(str(obj).encode(), 0)
    try:
        assert host.pg1_num_materials(h) == 3                   # a, b, (the second "a" is dropped: the name exists), c (the last is always pushed)
        assert [material(host, h, i)[0] for i in range(3)] == ["a", "b", "c"]
        assert host.pg1_num_surfaces(h) == 4
        assert [host.pg1_surface_name(h, i).decode() for i in range(4)] == ["", "first", "second", "third"]
        assert [host.pg1_surface_triangles(h, i) for i in range(4)] == [1, 2, 1, 1]
        assert host.pg1_surface_material(h, 1) == 0            # "a": the last usemtl seen before the next g
        assert host.pg1_surface_material(h, 2) == -1           # no such material
        pos, nrm, uv = scene_arrays(host, h, 1)
        assert pos.reshape(2, 3, 3)[1].tolist() == [[0, 0, 0], [1, 1, 0], [0, 1, 0]]
        assert np.allclose(nrm.reshape(-1, 3), [0, 0, 1]) and uv.reshape(2, 3, 2)[1].tolist() == [[0.25, 0.75], [0.25, 0.75], [1, 1]]
        name, v = material(host, h, 0)
        assert v[3:6].tolist() == [1, 0, 0] and int(v[14]) == 4         # "shader 4e1" is read with %d
        assert int(material(host, h, 2)[1][14]) == 3                    # an indented "newmtl" starts nothing; "shader x" keeps, "shader 3.9" reads 3
    finally:
        host.pg1_free_scene(h)
    h = host.pg1_load_obj(str(tmp_path / "absent.obj").encode(), 0)
    assert host.pg1_scene_rc(h) == -1
    host.pg1_free_scene(h)


# ------------------------------------------------------------------------------------------------ camera
def test_host_camera_matches_the_oracle(host, oracle_mod, cornell):
    o = oracle_mod.Oracle(cornell)
    c = cornell.camera
    f = (C.c_float * 3)(*c.view_from); a = (C.c_float * 3)(*c.view_at)
    for mode in (1, 0):
        rays = o.primary_rays(oracle_mod.make_params(sampling_width=1, jitter=0, aperture=0.0, camera_mode=mode))
        out = np.zeros(9, np.float32)
        for (x, y) in ((0, 0), (17, 5), (c.width - 1, c.height - 1)):
            host.pg1_camera_ray(c.width, c.height, c.fov_y, f, a, float(x), float(y), 0 if mode == 1 else 1, 200.0, 0.0, 0.0, out.ctypes.data)
            ref = rays[y * c.width + x].copy()
            if mode == 0:
                ref[7] = 0.0                       # generate_ray leaves time = 0; get_pixel overwrites it with IOR_AIR (raytracer.cpp:416)
            else:
                ref[7] = 0.0
            assert np.array_equal(out[:8], ref[:8]), (mode, x, y)


# ------------------------------------------------------------------------------------------------ Raytracer on the GPU
def _write_scene_files(sc, d):
    """OBJ + MTL + textures (PPM) + env (PPM) for Raytracer::LoadScene."""
    from PIL import Image
    for i, t in enumerate(sc.textures):
        rgb = t.data[:, :t.width * t.bpp].reshape(t.height, t.width, t.bpp)[..., 2::-1]
        Image.fromarray(np.ascontiguousarray(rgb)).save(os.path.join(d, f"tex{i}.ppm"), "PPM")
    for m in sc.materials:
        m.map_kd = f"tex{m.diffuse_tex}.ppm" if m.diffuse_tex >= 0 else ""
    scenes.write_obj(sc, os.path.join(d, "scene.obj"))
    e = sc.env
    rgb = e.data[:, :e.width * e.bpp].reshape(e.height, e.width, e.bpp)[..., 2::-1]
    Image.fromarray(np.ascontiguousarray(rgb)).save(os.path.join(d, "env.ppm"), "PPM")


@pytest.mark.gpu
def test_cpp_raytracer_renders_the_same_frame_as_the_abi(host, tmp_path, cornell):
    """Raytracer(w,h,fov,from,at) + LoadScene(obj, bg) + RenderFrame / get_pixel in C++ == the same scene through the
    Python mirror (bit for bit: same ABI calls underneath, same bytes in)."""
    import pgi_raytracing_b200 as P
    sc = scenes.cornell_like(); sc.camera = cornell.camera
    _write_scene_files(sc, str(tmp_path))
    # the OBJ loader re-normalises normals: give the Python side the same arrays
    h = host.pg1_load_obj(str(tmp_path / "scene.obj").encode(), 0)
    for i, m in enumerate(sc.meshes):
        pos, nrm, uv = scene_arrays(host, h, i)
        m.nrm = nrm.reshape(m.nrm.shape)
    host.pg1_free_scene(h)
    c = sc.camera
    f = (C.c_float * 3)(*c.view_from); a = (C.c_float * 3)(*c.view_at)
    rt = host.pg1_raytracer_create(c.width, c.height, c.fov_y, f, a, b"threads=0,verbose=3")
    assert rt, host.pg1_last_error()
    try:
        assert host.pg1_raytracer_load_scene(rt, str(tmp_path / "scene.obj").encode(), str(tmp_path / "env.ppm").encode()) == 0, host.pg1_last_error()
        counts = (C.c_int * 2)(); host.pg1_raytracer_counts(rt, counts)
        assert list(counts) == [len(sc.meshes), len(sc.materials)]
        ref_rt = P.raytracer_for(sc)
        for params in (dict(), dict(sampling_width=1, jitter=0, aperture=0.0, max_depth=3)):
            p = P.default_params(**params)
            host.pg1_raytracer_set(rt, p.sampling_width, p.jitter, p.focal_distance, p.aperture, p.max_depth, p.gamma_level, p.seed)
            img = np.zeros((c.height, c.width, 4), np.float32); rays = (C.c_ulonglong * 4)()
            assert host.pg1_raytracer_render(rt, img.ctypes.data, rays) == 0, host.pg1_last_error()
            ref, st = ref_rt.render(params)
            assert np.array_equal(img, ref, equal_nan=True)
            assert list(rays) == [st["primary"], st["shadow"], st["reflection"], st["refraction"]]
            px = np.zeros(4, np.float32)
            assert host.pg1_raytracer_get_pixel(rt, 10, 7, px.ctypes.data) == 0
            assert np.array_equal(px, ref[7, 10], equal_nan=True)
    finally:
        host.pg1_raytracer_destroy(rt)


@pytest.mark.gpu
def test_pg1_b200_executable(tmp_path, cornell):
    """The headless counterpart of the reference's executable: loads OBJ + env, renders, writes a PPM."""
    from PIL import Image
    sc = scenes.cornell_like(); sc.camera = cornell.camera
    _write_scene_files(sc, str(tmp_path))
    exe = os.path.join(PKG, "pg1_b200")
    out = subprocess.run([exe, str(tmp_path / "scene.obj"), str(tmp_path / "env.ppm"), "--width", "96", "--height", "64", "--frames", "2",
                          "--out", str(tmp_path / "frame.ppm")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert f"Surfaces = {len(sc.meshes)}" in out.stdout and "Mrays/s" in out.stdout
    img = np.asarray(Image.open(str(tmp_path / "frame.ppm")))
    assert img.shape == (64, 96, 3) and img.std() > 5
    # the progressive loop and the README to-dos: 4 accumulated frames of the path-tracing mode with hard shadows, as a PNG
    out = subprocess.run([exe, str(tmp_path / "scene.obj"), str(tmp_path / "env.ppm"), "--width", "96", "--height", "64", "--frames", "1", "--depth", "4",
                          "--accumulate", "4", "--path-tracing", "--hard-shadows", "--out", str(tmp_path / "acc.png")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "accumulated 4 frames" in out.stdout
    acc = np.asarray(Image.open(str(tmp_path / "acc.png")))
    assert acc.shape == (64, 96, 3) and acc.std() > 5 and not np.array_equal(acc, img)


@pytest.mark.parametrize("mode,size,level", [("RGBA", (64, 32), 9), ("RGB", (37, 21), 6), ("L", (50, 7), 1), ("LA", (9, 40), 9), ("P", (33, 33), 9), ("RGB", (300, 200), 0)])
def test_png_decoder(host, tmp_path, mode, size, level):
    """8-bit non-interlaced PNG (zlib inflate incl. stored / fixed / dynamic blocks, all five filters) -> top-down BGR(A)."""
    from PIL import Image
    w, h = size
    rgb = _test_picture(w, h, w * 3 + h)
    rng = np.random.default_rng(w + h)
    alpha = rng.integers(0, 256, (h, w), dtype=np.uint8)
    if mode == "RGBA":
        im = Image.fromarray(np.dstack([rgb, alpha]), "RGBA")
    elif mode == "LA":
        im = Image.fromarray(np.dstack([rgb[..., 0], alpha]), "LA")
    elif mode == "L":
        im = Image.fromarray(rgb[..., 1], "L")
    elif mode == "P":
        im = Image.fromarray(rgb).quantize(64)
    else:
        im = Image.fromarray(rgb)
    p = str(tmp_path / "t.png")
    im.save(p, "PNG", compress_level=level)
    ref = np.asarray(Image.open(p).convert("RGBA" if mode in ("RGBA", "LA") else "RGB"))
    buf, ww, hh, pitch, bpp = load_image(host, p)
    assert (ww, hh, bpp) == (w, h, ref.shape[-1]) and pitch == (w * bpp + 3) // 4 * 4
    got = buf[:, :bpp * w].reshape(h, w, bpp)
    assert np.array_equal(got[..., 2::-1], ref[..., :3])
    if bpp == 4:
        assert np.array_equal(got[..., 3], ref[..., 3])


def test_tutorial_2_fixture_through_the_cpp_texture(host):
    """pg1/tutorials.cpp:170-178 loads data/test4.png; the committed golden holds its FreeImage-style BGRA bytes."""
    p = "/root/reference/data/test4.png"
    if not os.path.exists(p):
        pytest.skip("reference data not mounted")
    buf, w, h, pitch, bpp = load_image(host, p)
    assert (w, h, bpp) == (64, 32, 4)
    assert np.array_equal(buf, np.load(os.path.join(GOLDEN, "test4_bgra.npy")))


def test_parallel_loader_equals_single_thread(host, tmp_path, monkeypatch):
    """LoadOBJ cuts the file into one slice per thread; the result must not depend on the thread count (slices end at
    line ends, faces keep file order, groups and usemtl are replayed sequentially)."""
    sc = scenes.triangle_soup(12000, seed=3, resolution=(8, 8))
    m = sc.meshes[0]
    third = m.ntris // 3
    sc.meshes = [scenes.Mesh("a", m.pos[:third], m.nrm[:third], m.uv[:third], 0), scenes.Mesh("b", m.pos[third:2 * third], m.nrm[third:2 * third], m.uv[third:2 * third], 0),
                 scenes.Mesh("c", m.pos[2 * third:], m.nrm[2 * third:], m.uv[2 * third:], 0)]
    path = str(tmp_path / "soup.obj")
    scenes.write_obj(sc, path)
    assert os.path.getsize(path) > (1 << 20)                  # large enough for the loader to go parallel
    results = []
    for threads in ("1", "3", "7"):
        monkeypatch.setenv("PG1_LOADER_THREADS", threads)
        h = host.pg1_load_obj(path.encode(), 0)
        assert host.pg1_num_surfaces(h) == 3
        results.append([(host.pg1_surface_name(h, i).decode(),) + scene_arrays(host, h, i) for i in range(3)])
        host.pg1_free_scene(h)
    for other in results[1:]:
        for (n0, p0, q0, u0), (n1, p1, q1, u1) in zip(results[0], other):
            assert n0 == n1 and np.array_equal(p0, p1) and np.array_equal(q0, q1) and np.array_equal(u0, u1)
    for i, mesh in enumerate(sc.meshes):
        assert np.array_equal(results[0][i][1], mesh.pos.reshape(-1, 9)) and np.array_equal(results[0][i][3], mesh.uv.reshape(-1, 6))


@pytest.mark.parametrize("size", [(37, 21), (640, 480), (1, 1)])
def test_png_writer(host, tmp_path, size):
    """WritePNG (headless presentation): a valid PNG (PIL reads it; CRCs and Adler-32 are checked there) with
    the 8-bit quantisation of the D3D11 back buffer (round(clamp(c,0,1)*255), NaN -> 0); larger than one 64 KiB stored block
    at 640x480; and our own decoder reads it back to the same bytes."""
    from PIL import Image
    w, h = size
    rng = np.random.default_rng(w * h)
    rgba = rng.uniform(-0.2, 1.2, (h, w, 4)).astype(np.float32)
    rgba[0, 0, 0] = np.nan; rgba[-1, -1, 1] = np.inf
    p = str(tmp_path / "frame.png")
    assert host.pg1_write_image(p.encode(), rgba.ctypes.data, w, h) == 0
    c = rgba[..., :3].copy(); c[np.isnan(c)] = 0.0
    want = np.floor(np.clip(c, 0.0, 1.0) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
    im = Image.open(p); im.load()
    assert im.mode == "RGB" and im.size == (w, h)
    assert np.array_equal(np.asarray(im), want)
    buf, ww, hh, pitch, bpp = load_image(host, p)
    assert (ww, hh, bpp) == (w, h, 3)
    assert np.array_equal(buf[:, :3 * w].reshape(h, w, 3)[..., ::-1], want)
    ppm = str(tmp_path / "frame.ppm")
    assert host.pg1_write_image(ppm.encode(), rgba.ctypes.data, w, h) == 0          # by extension
    assert np.array_equal(np.asarray(Image.open(ppm)), want)


def test_jpeg_decoder_fuzz_against_libjpeg(host, tmp_path):
    """Random small JPEGs -- sizes 1..90, quality 1..100, 4:4:4 / 4:2:2 / 4:2:0, optimised Huffman tables, restart intervals,
    greyscale -- decoded bit-exactly as libjpeg-turbo does (through PIL).  Found by this fuzz: libjpeg switches from the
    triangle filter to plain replication when the down-sampled row has two samples or fewer (images up to 4 pixels wide)."""
    from PIL import Image
    rng = np.random.default_rng(11)
    sizes = [(w, h) for w in (1, 2, 3, 4, 5) for h in (1, 2, 3, 9)] + [(int(rng.integers(1, 90)), int(rng.integers(1, 90))) for _ in range(60)]
    for k, (w, h) in enumerate(sizes):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if k % 2 else _test_picture(w, h, k)
        grey = k % 7 == 3
        kw = dict(quality=int(rng.integers(1, 101)), optimize=bool(k % 3 == 0))
        if not grey:
            kw["subsampling"] = int(rng.choice([0, 1, 2]))
        if k % 5 == 1:
            kw["restart_marker_blocks"] = int(rng.choice([1, 2, 5]))
        p = str(tmp_path / f"f{k}.jpg")
        im = Image.fromarray(img[..., 0], "L") if grey else Image.fromarray(img)
        try:
            im.save(p, "JPEG", **kw)
        except TypeError:                      # an older Pillow without restart_marker_blocks
            kw.pop("restart_marker_blocks", None); im.save(p, "JPEG", **kw)
        ref = np.asarray(Image.open(p).convert("RGB"))
        buf, ww, hh, pitch, bpp = load_image(host, p)
        assert (ww, hh, bpp) == (w, h, 3), (k, w, h, kw)
        assert np.array_equal(buf[:, :3 * w].reshape(h, w, 3)[..., ::-1], ref), (k, w, h, kw, grey)


_MALFORMED_FUZZ = r"""
import ctypes as C, sys, numpy as np
lib = C.CDLL(sys.argv[1]); lib.pg1_load_image.restype = C.c_void_p; lib.pg1_load_image.argtypes = [C.c_char_p]
lib.pg1_free_image.argtypes = [C.c_void_p]; lib.pg1_image_info.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
seeds = [open(p, "rb").read() for p in sys.argv[3:]]
rng = np.random.default_rng(5)
out = sys.argv[2]
ok = bad = 0
for it in range(4000):
    d = bytearray(seeds[it % len(seeds)])
    hdr = min(len(d) - 4, 700)                   # markers, tables, frame and scan headers live here
    mode = it % 5
    if mode == 0:
        for _ in range(int(rng.integers(1, 6))): d[int(rng.integers(2, hdr))] = int(rng.integers(0, 256))
    elif mode == 1:
        d = d[: int(rng.integers(2, len(d)))]    # truncated anywhere
    elif mode == 2:
        i = int(rng.integers(2, hdr)); d[i:i + 2] = bytes([0xFF, int(rng.choice([0xC0, 0xC4, 0xDA, 0xDB, 0xDD]))])   # a marker where none belongs
    elif mode == 3:
        i = int(rng.integers(2, hdr)); d[i] = 0xFF; d[i + 1] = int(rng.integers(0xC0, 0xE0)); d[i + 2] = 0; d[i + 3] = int(rng.integers(0, 20))   # segment with a tiny length
    else:
        for _ in range(40): d[int(rng.integers(2, len(d)))] = int(rng.integers(0, 256))              # noise in the entropy-coded data too
    open(out, "wb").write(bytes(d))
    h = lib.pg1_load_image(out.encode())
    if h:
        info = (C.c_int * 4)(); lib.pg1_image_info(h, info)
        assert 0 < info[0] <= 65535 and 0 < info[1] <= 65535 and info[3] in (3, 4)
        lib.pg1_free_image(h); ok += 1
    else:
        bad += 1
print(ok, bad)
"""


def test_malformed_jpeg_files_are_rejected_not_trusted(host, tmp_path):
    """header fields come from the file.  4000 mutated / truncated JPEGs (bytes flipped in the marker segments,
    markers and short segment lengths planted, files cut anywhere, noise in the entropy-coded data) must each end in an
    error or an image -- in a child process, so that an out-of-bounds access fails this test instead of ending the run."""
    import subprocess
    import sys
    from PIL import Image
    rng = np.random.default_rng(3)
    seeds = []
    for k, (sub, q, grey) in enumerate([(0, 90, False), (2, 50, False), (1, 75, False), (0, 30, True)]):
        p = str(tmp_path / f"seed{k}.jpg")
        img = _test_picture(33 + 7 * k, 21 + 5 * k, k)
        kw = dict(quality=q, optimize=bool(k % 2))
        if grey:
            Image.fromarray(img[..., 0], "L").save(p, "JPEG", **kw)
        else:
            Image.fromarray(img).save(p, "JPEG", subsampling=sub, **kw)
        seeds.append(p)
    r = subprocess.run([sys.executable, "-c", _MALFORMED_FUZZ, HOST_LIB, str(tmp_path / "m.jpg")] + seeds, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    ok, bad = (int(x) for x in r.stdout.split())
    assert ok + bad == 4000 and bad > 500 and ok > 100, (ok, bad)


def test_loader_float_scan_equals_strtof_and_crlf(host, tmp_path):
    """The loader's in-place float scan (its own two-tier exact fast path, strtof behind it) against glibc strtof (what the
    reference's sscanf("%f") does): integers, fixed and scientific notation, leading '.', trailing '.', 16-19 digit
    mantissas, reprs of float32 values, 30-digit midpoints between adjacent floats -- bit for bit; a malformed
    token stops the line as sscanf does (the remaining components keep 0).  The same file with CRLF line ends loads identically."""
    import re
    libc = C.CDLL("libc.so.6"); libc.strtof.restype = C.c_float; libc.strtof.argtypes = [C.c_char_p, C.c_void_p]
    rng = np.random.default_rng(8)

    def tok():
        k = int(rng.integers(0, 10)); sign = str(rng.choice(["", "-", "+"], p=[0.6, 0.3, 0.1]))
        if k == 8: return repr(float(np.float32(rng.uniform(-100, 100) * 10.0 ** int(rng.integers(-6, 6)))))      # 16-17 digits: the x87 tier
        if k == 9: return sign + (str(int(rng.integers(10 ** 8, 10 ** 9))) + str(int(rng.integers(10 ** 9, 10 ** 10))))[: int(rng.integers(16, 20))] + "e-" + str(int(rng.integers(0, 28)))
        if k == 0: return sign + str(int(rng.integers(0, 10 ** int(rng.integers(1, 12)))))
        if k == 1: return sign + f"{rng.uniform(0, 1e3):.{int(rng.integers(0, 12))}f}"
        if k == 2: return sign + f"{rng.uniform(0, 1):.{int(rng.integers(1, 25))}f}"
        if k == 3: return sign + f"{rng.uniform(1, 10):.{int(rng.integers(0, 10))}f}e{int(rng.integers(-45, 39))}"
        if k == 4: return sign + f"{rng.uniform(1, 10):.{int(rng.integers(0, 17))}g}E{int(rng.integers(-10, 10)):+d}"
        if k == 5: return sign + "." + str(int(rng.integers(0, 10 ** 9)))
        if k == 6: return sign + str(int(rng.integers(0, 10 ** 6))) + "."
        f = np.float32(rng.uniform(1e-3, 1e3)); g = np.nextafter(f, np.float32(np.inf))
        return sign + f"{(float(f) + float(g)) / 2:.30f}"

    n = 6000
    def valid_tok():
        while True:
            t = tok()
            if re.fullmatch(r"[+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?", t):
                return t

    strs = [[valid_tok() for _ in range(3)] for _ in range(n)]
    strs[7] = ["1.5", "--2", "3"]                                   # malformed: sscanf stops here
    body = "".join("v " + " ".join(s) + "\n" for s in strs) + "vn 0 0 1\nvt 0 0\ng a\n" + "".join(f"f {i + 1}/1/1 {i + 2}/1/1 {i + 3}/1/1\n" for i in range(0, n - 2, 3))
    want = np.array([[libc.strtof(t.encode(), None) for t in s] for s in strs], np.float32)
    want[7] = [1.5, 0.0, 0.0]
    for name, text in (("lf.obj", body), ("crlf.obj", body.replace("\n", "\r\n"))):
        p = tmp_path / name
        p.write_bytes(text.encode())
        h = host.pg1_load_obj(str(p).encode(), 0)
        try:
            assert host.pg1_num_surfaces(h) == 1 and host.pg1_surface_name(h, 0).decode() == "a"
            pos, _, _ = scene_arrays(host, h, 0)
            assert np.array_equal(pos.reshape(-1, 3).view(np.uint32), want[: pos.size // 3].view(np.uint32)), name
        finally:
            host.pg1_free_scene(h)


def test_loader_flip_yz(host, tmp_path):
    """flip_yz (pg1/objloader.cpp:311-316, :327-333): "x y z" is read as (x, -z, y) for positions and normals (normals are
    normalised afterwards); texture coordinates are untouched."""
    p = tmp_path / "f.obj"
    p.write_text("v 1 2 3\nv 4 5 6\nv 7 8 10\nvn 0 3 4\nvt 0.25 0.5\ng a\nf 1/1/1 2/1/1 3/1/1\n")
    for flip, want_pos, want_n in ((0, [1, 2, 3, 4, 5, 6, 7, 8, 10], [0, 0.6, 0.8]), (1, [1, -3, 2, 4, -6, 5, 7, -10, 8], [0, -0.8, 0.6])):
        h = host.pg1_load_obj(str(p).encode(), flip)
        try:
            pos, nrm, uv = scene_arrays(host, h, 0)
            assert pos.reshape(-1).tolist() == [float(x) for x in want_pos]
            assert np.allclose(nrm.reshape(3, 3), want_n, atol=1e-7) and uv.reshape(3, 2).tolist() == [[0.25, 0.5]] * 3
        finally:
            host.pg1_free_scene(h)
