"""Drop-in import name for the reference.

/root/reference/gaussian_renderer/__init__.py:16 does
    from diff_gaussian_rasterization_depth_alpha import GaussianRasterizationSettings, GaussianRasterizer
Putting this repository's root on PYTHONPATH makes that import resolve here, so
`gaussian_renderer.render()` / `doll_render()`, `s2_registration.py`, `s3_appearance.py` and
`inference.py` call the B200-native rasterizer unchanged.

The implementation lives in the hyphen-named package directory `gaussian-garments_b200/`
(task contract); it is registered in sys.modules as `gaussian_garments_b200`.
"""
import importlib.util
import os
import sys

_ALIAS = "gaussian_garments_b200"


def _load_impl():
    if _ALIAS in sys.modules:
        return sys.modules[_ALIAS]
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gaussian-garments_b200")
    spec = importlib.util.spec_from_file_location(_ALIAS, os.path.join(root, "__init__.py"),
                                                  submodule_search_locations=[root])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_ALIAS] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception:
        sys.modules.pop(_ALIAS, None)
        raise
    return mod


_impl = _load_impl()
GaussianRasterizationSettings = _impl.GaussianRasterizationSettings
GaussianRasterizer = _impl.GaussianRasterizer
rasterize_gaussians = _impl.rasterize_gaussians
mark_visible = _impl.mark_visible

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians", "mark_visible"]
