"""CPU only: throughput of the OBJ loader (host/objloader.cpp) on the avenger stand-in written as OBJ, by thread count.
python tools/loader_bench.py [detail]   (detail 1.0 = 171 k triangles; 4.0 = about 2.7 M)"""
import ctypes as C, os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pgi_raytracing_b200 import scenes

detail = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
lib = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pgi_raytracing_b200", "libpg1_host.so"))
lib.pg1_load_obj.restype = C.c_void_p; lib.pg1_load_obj.argtypes = [C.c_char_p, C.c_int]
lib.pg1_free_scene.argtypes = [C.c_void_p]; lib.pg1_num_surfaces.argtypes = [C.c_void_p]; lib.pg1_surface_triangles.argtypes = [C.c_void_p, C.c_int]
sc = scenes.avenger_proxy(detail=detail, with_images=False)
d = tempfile.mkdtemp()
path = os.path.join(d, "scene.obj")
scenes.write_obj(sc, path)
mb = os.path.getsize(path) / 1e6
print(f"{sc.ntris} triangles, {len(sc.meshes)} surfaces, OBJ {mb:.1f} MB")
for threads in (1, 2, 4, 8, 16, 0):
    if threads > (os.cpu_count() or 1):
        continue
    if threads:
        os.environ["PG1_LOADER_THREADS"] = str(threads)
    else:
        os.environ.pop("PG1_LOADER_THREADS", None)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        h = lib.pg1_load_obj(path.encode(), 0)
        dt = time.perf_counter() - t0
        n = sum(lib.pg1_surface_triangles(h, i) for i in range(lib.pg1_num_surfaces(h)))
        lib.pg1_free_scene(h)
        best = min(best, dt)
    print(f"threads={threads or 'auto'}: {best * 1e3:.1f} ms  {mb / best:.0f} MB/s  {n / best / 1e6:.2f} M triangles/s")
