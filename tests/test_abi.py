"""The C-ABI library: loads, exports every symbol include/pgrt.h declares, struct layouts agree with the header,
and it refuses to work without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from conftest import ROOT
from pgi_raytracing_b200 import _lib as L

HEADER = os.path.join(ROOT, "include", "pgrt.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pgrt_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_and_loads():
    assert os.path.exists(L.LIB_PATH), "run __graft_entry__.build() first"
    lib = L.load()
    assert lib.pgrt_version().startswith(b"pgrt-b200")


def test_every_declared_symbol_is_exported_and_bound():
    lib = C.CDLL(L.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pgrt.h but not exported"
        assert n in L.SYMBOLS, f"{n} has no ctypes binding in _lib.SYMBOLS"
    assert sorted(L.SYMBOLS) == names


def test_struct_layouts_match_the_header_compiled_as_plain_c():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "pgrt.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(pgrt_material), sizeof(pgrt_light), sizeof(pgrt_render_params),
         sizeof(pgrt_build_stats), sizeof(pgrt_render_stats), sizeof(pgrt_rayhit));
  printf("%zu %zu %zu %zu\n", offsetof(pgrt_material, type), offsetof(pgrt_render_params, seed), offsetof(pgrt_render_stats, frame_ms), offsetof(pgrt_rayhit, geomID));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c"); exe = os.path.join(d, "t")
        open(c, "w").write(prog)
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    sizes = [int(x) for x in out[:6]]
    assert sizes == [C.sizeof(L.Material), C.sizeof(L.Light), C.sizeof(L.RenderParams), C.sizeof(L.BuildStats), C.sizeof(L.RenderStats), 80]
    assert [int(x) for x in out[6:]] == [L.Material.type.offset, L.RenderParams.seed.offset, L.RenderStats.frame_ms.offset, 72]


def test_default_params_are_the_reference_constants():
    """pg1/raytracer.cpp:398-400 (3x3, focal 200, aperture 5), :282 (depth 7), :450 (gamma 0.5)."""
    from pgi_raytracing_b200 import default_params
    p = default_params()
    assert (p.sampling_width, p.jitter, p.focal_distance, p.aperture, p.max_depth, p.gamma_level) == (3, 1, 200.0, 5.0, 7, 0.5)
    assert (p.camera_mode, p.shader_mode) == (0, 0)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pgi_raytracing_b200 import Raytracer, PgrtError
    with pytest.raises(PgrtError) as e:
        Raytracer(64, 48, 0.7, (0, 0, 1), (0, 0, 0))
    assert e.value.code == L.PGRT_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pgi_raytracing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "pg_oracle" not in txt and "libpg_oracle" not in txt, f
