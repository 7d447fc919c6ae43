"""Host-side logic that needs no GPU: tile sharding arithmetic, stand-in scene generators, OBJ writer, bench helpers."""
import os

import numpy as np
import pytest

from pgi_raytracing_b200 import scenes
from pgi_raytracing_b200 import dist as D


@pytest.mark.parametrize("wh", [(640, 480), (1920, 1080), (3840, 2160), (100, 37), (33, 9), (5, 3)])
@pytest.mark.parametrize("n_ranks", [1, 2, 3, 8])
def test_tiles_partition_the_frame_exactly(wh, n_ranks):
    w, h = wh
    seen = np.zeros((h, w), np.int32)
    for r in range(n_ranks):
        x, y, valid = D.slot_pixels(w, h, r, n_ranks)
        assert x.shape[0] == D.shard_pixels(w, h, n_ranks)
        np.add.at(seen, (y[valid], x[valid]), 1)
    assert np.all(seen == 1)


def test_tile_untile_roundtrip_and_warp_coherence():
    w, h, n = 100, 37, 3
    frame = np.random.default_rng(0).random((h, w, 4)).astype(np.float32)
    g = np.concatenate([D.tile_numpy(frame, r, n) for r in range(n)])
    assert np.array_equal(D.untile_numpy(g, w, h, n), frame)
    # 32 consecutive slots cover an 8x4 pixel block (coherent primary rays within a warp)
    x, y, valid = D.slot_pixels(640, 480, 0, 1)
    assert np.ptp(x[:32]) == 7 and np.ptp(y[:32]) == 3
    # round-robin dealing: rank r of n owns tiles r, r+n, ...
    x1, y1, _ = D.slot_pixels(640, 480, 1, 4)
    assert (x1[0], y1[0]) == (32, 0) and (x1[256], y1[256]) == (32 * 5, 0)


def test_avenger_proxy_shape_and_determinism():
    a = scenes.avenger_proxy(detail=0.05, with_images=False)
    b = scenes.avenger_proxy(detail=0.05, with_images=False)
    assert len(a.meshes) == 94 and len(a.materials) == 5                    # the counts the reference's screenshot shows
    assert [m.type for m in a.materials] == [3, 4, 3, 3, 3]
    assert sum(1 for m in a.meshes if a.materials[m.material].type == 4) == 1   # one closed dielectric shell
    assert {a.materials[m.material].diffuse_tex for m in a.meshes} == {-1, 0, 1}
    for ma, mb in zip(a.meshes, b.meshes):
        assert np.array_equal(ma.pos, mb.pos) and np.array_equal(ma.nrm, mb.nrm) and np.array_equal(ma.uv, mb.uv)
    allp = np.concatenate([m.pos.reshape(-1, 3) for m in a.meshes])
    assert allp.min() > -110 and allp.max() < 110
    for m in a.meshes:
        assert m.pos.dtype == np.float32 and m.pos.shape == (m.ntris, 3, 3) and m.uv.shape == (m.ntris, 3, 2)
        assert np.allclose(np.linalg.norm(m.nrm, axis=-1), 1.0, atol=1e-4)


def test_full_detail_triangle_count():
    sc = scenes.avenger_proxy(with_images=False)
    assert 1.5e5 < sc.ntris < 2.5e5


def test_soup_matches_config_c5_statistics():
    sc = scenes.triangle_soup(20000, seed=1)
    p = sc.meshes[0].pos
    c = p.mean(axis=1)
    assert c.min() > -101.5 and c.max() < 101.5 and abs(c.mean()) < 2.0
    e = np.abs(p[:, 1] - p[:, 0])
    assert e.max() <= 1.0 and 0.45 < e.mean() < 0.55
    assert np.array_equal(scenes.triangle_soup(20000, seed=1).meshes[0].pos, p)
    assert not np.array_equal(scenes.triangle_soup(20000, seed=2).meshes[0].pos, p)


def test_images_are_bgr_topdown_with_freeimage_pitch():
    rgb = np.zeros((2, 3, 3), np.uint8); rgb[0, 1] = (10, 20, 30)
    im = scenes.Image.from_rgb(rgb)
    assert (im.width, im.height, im.bpp, im.pitch) == (3, 2, 3, 12)
    assert list(im.data[0, 3:6]) == [30, 20, 10]
    env = scenes.make_envmap(256, 128, seed=1)
    assert env.data.shape == (128, 768) and env.data.std() > 10
    assert np.array_equal(env.data, scenes.make_envmap(256, 128, seed=1).data)


def test_obj_writer_emits_the_dialect_loadobj_reads(tmp_path):
    sc = scenes.cornell_like()
    path = os.path.join(tmp_path, "scene.obj")
    scenes.write_obj(sc, path, mtl_text=scenes.AVENGER_MTL_TEXT)
    lines = open(path).read().splitlines()
    assert lines[1].startswith("mtllib scene.mtl")
    assert sum(1 for l in lines if l.startswith("g ")) == len(sc.meshes)
    assert sum(1 for l in lines if l.startswith("f ")) == sc.ntris
    assert sum(1 for l in lines if l.startswith("v ")) == 3 * sc.ntris
    f = next(l for l in lines if l.startswith("f "))
    assert f == "f 1/1/1 2/2/2 3/3/3"
    # float32 values survive the text round trip
    v = next(l for l in lines if l.startswith("v ")).split()[1:]
    assert np.array_equal(np.array(v, np.float64).astype(np.float32), sc.meshes[0].pos[0, 0])
    assert "Ks 1.0. 1.0 1.0" in open(os.path.join(tmp_path, "scene.mtl")).read()


def test_bench_contract_figures():
    import bench
    assert bench.algorithmic_bytes_per_ray(200_000) == 720.0        # N = 2e5 -> D = 6
    assert bench.algorithmic_bytes_per_ray(10_000_000) == 880.0     # N = 1e7 -> D = 8
    sc, p, desc = bench.workload("c1")
    assert (sc.camera.width, sc.camera.height) == (640, 480) and p["max_depth"] == 7 and p["sampling_width"] == 1


def test_byte_over_255_identity():
    """csrc/shading.cuh byte_over_255: q = b*r, q + fma(-q, 255, b)*r is the correctly rounded b / 255.0f for all bytes
    (float64 stands in for the two FMAs: every product here is exact in 53 bits)."""
    import numpy as np
    b = np.arange(256, dtype=np.float32)
    ref = (b / np.float32(255.0)).astype(np.float32)
    r = np.float32(0.003921568859368563)
    assert r == np.float32(1.0) / np.float32(255.0)
    q = (b * r).astype(np.float32)
    e = (b.astype(np.float64) - q.astype(np.float64) * 255.0).astype(np.float32)
    q2 = (q.astype(np.float64) + e.astype(np.float64) * np.float64(r)).astype(np.float32)
    assert np.array_equal(q2, ref) and np.count_nonzero(q != ref) > 50


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys, the
    BASELINE.json metric and unit, `impl: reference`, a cpu_baseline describing the run and an e2e object without copies."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    with open(os.path.join(root, "BASELINE.json")) as f:
        base = json.load(f)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["unit"] == "Mrays/s" and line["value"] > 0 and "workload" in line["config"]
    if isinstance(base.get("unit"), str):
        assert line["unit"] == base["unit"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_and_smoke_fail_loudly_without_a_gpu():
    """No silent CPU path: on a box without CUDA `bench.py` (our arm) exits non-zero with a message, and smoke() raises."""
    import os, subprocess, sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
    sys.path.insert(0, root)
    import __graft_entry__ as g
    with pytest.raises(Exception):
        g.smoke()


def test_tiling_properties_for_arbitrary_sizes():
    """Property test (hypothesis): for any frame size and rank count the shards partition the frame, have the same padded
    length on every rank, and tile -> gather order -> untile is the identity."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(w=st.integers(1, 300), h=st.integers(1, 200), n=st.integers(1, 9), seed=st.integers(0, 2 ** 31 - 1))
    def prop(w, h, n, seed):
        seen = np.zeros((h, w), np.int32)
        for r in range(n):
            x, y, valid = D.slot_pixels(w, h, r, n)
            assert x.shape[0] == D.shard_pixels(w, h, n) and x.shape[0] % 256 == 0
            np.add.at(seen, (y[valid], x[valid]), 1)
        assert np.all(seen == 1)
        frame = np.random.default_rng(seed).random((h, w, 4)).astype(np.float32)
        g = np.concatenate([D.tile_numpy(frame, r, n) for r in range(n)])
        assert np.array_equal(D.untile_numpy(g, w, h, n), frame)

    prop()


def test_host_side_look_at_a_flag_word_compares_like_the_stream_wait():
    """dist.ShardedRenderer looks at the "consumed" word from the host before it falls back to cuStreamWaitValue32 (an
    unsatisfied stream wait parks its hardware channel): same cyclic >= comparison, bounded spin, no CUDA involved."""
    import time
    sr = object.__new__(D.ShardedRenderer)
    sr.ctl = np.zeros(8, np.uint32)
    sr.ctl[2] = 5
    assert sr._host_sees(8, 5) and sr._host_sees(8, 3)
    t0 = time.perf_counter()
    assert not sr._host_sees(8, 6, spin_s=2e-3)                      # not reached: gives up after the spin, does not block
    assert 1e-3 < time.perf_counter() - t0 < 0.5
    sr.ctl[2] = 3                                                    # tags wrap at 2^32: 3 is "after" 0xFFFFFFFE
    assert sr._host_sees(8, 0xFFFFFFFE) and not sr._host_sees(8, 8, spin_s=1e-4)
    assert os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS") == "32"     # set by the package before any CUDA context exists (its default here)
