"""B200-native replacement of the per-pixel render loop of rddrdhd/PGI_RayTracing.

``csrc/``   hand-written sm_100a CUDA kernels + the C ABI (``include/pgrt.h``) -> ``libpgrt_b200.so``
``api.py``  host-side mirror of the reference's ``Raytracer`` class over that ABI
``scenes.py`` seeded stand-in scenes / textures / env maps (the reference's named assets are absent)
"""
import os as _os

# Every frame slot of a context renders on its own stream, up to 32 of them in flight: with the driver's default of 8 hardware
# channels several of those streams share one, and a stream that waits (a completion flag, an event) holds back its
# channel-mates.  Read by the driver when the CUDA context is created, so it only helps when this import comes first.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .api import Raytracer, PgrtError, default_params, raytracer_for, to_srgb8, RAYHIT_DTYPE  # noqa: F401
from . import scenes  # noqa: F401
