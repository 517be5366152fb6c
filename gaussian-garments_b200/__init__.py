"""gaussian-garments_b200 -- B200-native differentiable Gaussian-splat rasterizer.

This package holds only what the hot path of eth-ait/Gaussian-Garments needs
(SURVEY.md section 8): the host-side mirror of the rasterizer interface that
`gaussian_renderer/__init__.py:16,39-54,103-111` of the reference binds to, the
ctypes loader of the C-ABI library built from `csrc/`, synthetic scenes/cameras
that follow the reference's conventions, and the one-view-per-GPU gradient
exchange.

The directory name carries a hyphen (task contract), so it is imported through
`diff_gaussian_rasterization_depth_alpha` (the drop-in name the reference
imports) or `gg_import()` in `__graft_entry__.py`, both of which register it in
`sys.modules` as `gaussian_garments_b200`.
"""

from .rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_gaussians,
    mark_visible,
)
from . import cameras, scenes, densify, ply_io  # noqa: F401
from .mesh_binding import bind_to_mesh, FusedMeshBinding  # noqa: F401
from .losses import photometric_loss  # noqa: F401
from .visibility import cast_rays_from_point, visible_mask, visible_mask_multi  # noqa: F401

__version__ = "0.1.0"
