"""B200-native replacement of the per-pixel render loop of rddrdhd/PGI_RayTracing.

``csrc/``   hand-written sm_100a CUDA kernels + the C ABI (``include/pgrt.h``) -> ``libpgrt_b200.so``
``api.py``  host-side mirror of the reference's ``Raytracer`` class over that ABI
``scenes.py`` seeded stand-in scenes / textures / env maps (the reference's named assets are absent)
"""
from .api import Raytracer, PgrtError, default_params, raytracer_for, to_srgb8, RAYHIT_DTYPE  # noqa: F401
from . import scenes  # noqa: F401
