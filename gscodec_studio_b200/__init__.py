"""gscodec_studio_b200 -- B200-native (sm_100a) rasterization hot path of GSCodec Studio / gsplat.

Drop-in for the reference's `gsplat.rasterization()` + `gsplat.compression_simulation` call path:
the same Python API on top of hand-written CUDA kernels reached through the C-ABI library
`libgsb200.so` (include/gsb200.h).  Nothing here falls back to PyTorch or the CPU: without the built
library (or without a CUDA device) the ops raise.
"""
from .cuda._wrapper import (  # noqa: F401
    fully_fused_projection,
    isect_offset_encode,
    isect_tiles,
    persp_proj,
    proj,
    quat_scale_to_covar_preci,
    rasterize_to_indices_in_range,
    rasterize_to_pixels,
    spherical_harmonics,
    world_to_cam,
)
from .cuda._wrapper_2dgs import (  # noqa: F401
    accumulate_2dgs,
    fully_fused_projection_2dgs,
    rasterize_to_indices_in_range_2dgs,
    rasterize_to_pixels_2dgs,
)
from .rendering import rasterization  # noqa: F401
from .rendering_2dgs import rasterization_2dgs, rasterization_2dgs_inria_wrapper  # noqa: F401
from . import distributed  # noqa: F401
from .version import __version__  # noqa: F401

__all__ = [
    "rasterization",
    "fully_fused_projection",
    "isect_tiles",
    "isect_offset_encode",
    "rasterize_to_pixels",
    "spherical_harmonics",
    "quat_scale_to_covar_preci",
    "proj",
    "persp_proj",
    "world_to_cam",
    "rasterize_to_indices_in_range",
    "rasterization_2dgs",
    "rasterization_2dgs_inria_wrapper",
    "fully_fused_projection_2dgs",
    "rasterize_to_pixels_2dgs",
    "rasterize_to_indices_in_range_2dgs",
    "accumulate_2dgs",
    "distributed",
]
