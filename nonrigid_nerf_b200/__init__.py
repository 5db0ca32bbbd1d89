"""nonrigid_nerf_b200 -- B200-native (sm_100a) implementation of NR-NeRF's per-ray volumetric
rendering hot path behind the reference's Python entry points.  See DESIGN.md / INTEGRATION.md."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
