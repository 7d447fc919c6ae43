"""Developer check: CUDA path vs oracle on a few scenes, with numbers.  Run under gpurun."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pgi_raytracing_b200 import scenes, raytracer_for, default_params, to_srgb8
from oracle.oracle import Oracle, make_params, make_rayhits


def compare(name, sc, pdict, save=None):
    orc = Oracle(sc)
    rt = raytracer_for(sc)
    print(f"[{name}] tris={sc.ntris} build={rt.build_stats}")
    t0 = time.time(); ref, g0, p0, st0 = orc.render(make_params(**pdict), threads=0); t_cpu = time.time() - t0
    img, st = rt.render(pdict, profile=True)
    g1, p1 = rt.primary_ids(pdict)
    ids = np.mean((g0 == g1) & (p0 == p1))
    a, b = to_srgb8(ref).astype(int), to_srgb8(img).astype(int)
    d = np.abs(a - b).max(axis=-1)
    mse = np.mean((a - b) ** 2.0)
    psnr = 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
    exact = np.mean(np.all((ref == img) | (np.isnan(ref) & np.isnan(img)), axis=-1))
    print(f"[{name}] ids_equal={ids:.6f} le2lsb={np.mean(d <= 2):.6f} max_lsb={d.max()} psnr={psnr:.2f} bit_exact_px={exact:.6f}")
    print(f"[{name}] rays oracle={ {k: st0[k] for k in ('primary','shadow','reflection','refraction','total')} }")
    print(f"[{name}] rays gpu   ={ {k: st[k] for k in ('primary','shadow','reflection','refraction','total')} }")
    print(f"[{name}] gpu frame_ms={st['frame_ms']:.3f} trace_ms={st['trace_ms']:.3f} shade_ms={st['shade_ms']:.3f} launches={st['launches']} "
          f"Mrays/s={st['total'] / st['frame_ms'] / 1e3:.1f}  cpu_s={t_cpu:.2f} ({st0['total'] / t_cpu / 1e6:.2f} Mrays/s all cores)")
    if save:
        from PIL import Image
        Image.fromarray(to_srgb8(img)).save(save)
    return rt


if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    # T1
    sc = scenes.single_triangle()
    rt = raytracer_for(sc)
    rh = make_rayhits([[0.1, 0.2, 2.0]], [[0, 0, -1]], tnear=np.finfo(np.float32).tiny)
    out = rt.intersect(rh)
    print("T1", out["tfar"], out["u"], out["v"], out["geomID"], out["primID"], out["Ng_z"])
    print("T1 interp", rt.interpolate([0], [0], out["u"], out["v"], 0), rt.interpolate([0], [0], out["u"], out["v"], 1))
    sc = scenes.cornell_like()
    sc.camera = scenes.Camera(160, 120, sc.camera.fov_y, sc.camera.view_from, sc.camera.view_at)
    compare("cornell C1", sc, dict(sampling_width=1, jitter=0, aperture=0.0), "gpurun_out/cornell.png")
    compare("cornell 3x3 dof", sc, dict())
    sc = scenes.avenger_proxy()
    rt = compare("avenger C1", sc, dict(sampling_width=1, jitter=0, aperture=0.0), "gpurun_out/avenger.png")
    rt.set_camera(1920, 1080, sc.camera.fov_y, sc.camera.view_from, sc.camera.view_at)
    for depth in (7, 10):
        for _ in range(3):
            img, st = rt.render(dict(sampling_width=1, jitter=0, aperture=0.0, max_depth=depth), profile=True)
        print(f"avenger 1080p depth{depth}: frame_ms={st['frame_ms']:.3f} trace_ms={st['trace_ms']:.3f} shade_ms={st['shade_ms']:.3f} rays={st['total']} "
              f"Mrays/s={st['total'] / st['frame_ms'] / 1e3:.1f} launches={st['launches']}")
