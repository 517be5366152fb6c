"""Adaptive density control for mesh-bound Gaussians ("next" row N4 of SURVEY.md 8f) -- host logic.

The consumer of the rasterizer's `means2D` gradient side channel.  Mirrors, for any object that carries the
reference's attribute names, what the template stage of s2_registration.py (:312-322) does every
`densification_interval` iterations:

    add_densification_stats(model, viewspace_points, visibility_filter)     scene/gaussian_model.py:410-412
    densify_and_prune(model, max_grad, min_opacity, extent, max_screen_size) scene/gaussian_model.py:390-406
        densify_and_clone   scene/mesh_gaussian_model.py:184-203  (small splats: duplicate, same face)
        densify_and_split   scene/mesh_gaussian_model.py:155-182  (large splats: N samples of the splat's own normal
                                                                   distribution, local scale / (0.8 N), same face)
        prune_points        scene/mesh_gaussian_model.py:130-153  (never takes a face's last Gaussian)
    reset_opacity(model)                                                     scene/gaussian_model.py:211-214

plus the Adam-state surgery those need (rows of exp_avg / exp_avg_sq follow the rows of their parameter; the mesh's
"vertex" group is left alone: scene/gaussian_model.py:257-315).

Everything is a row operation on the per-Gaussian table (six parameter tensors + binding + three statistics), so it
is written once as `_Rows`: keep-rows and append-rows, applied to parameters, optimizer state and side arrays alike.
All tensors stay on their device; nothing synchronises except the two size-determining mask counts torch needs.

Multi-GPU: parameters are replicated, so every rank must take IDENTICAL decisions and draw IDENTICAL
samples.  Decisions are functions of the (all-reduced) gradient statistics; the only randomness is the split's
torch.normal, drawn from an explicit generator that `rank_consistent_generator(seed, iteration, device)` seeds the
same way on every rank.  `state_fingerprint` + `assert_rank_consistent` verify it with one tiny all-reduce.
"""

from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

PARAM_ATTRS = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity",
               "scaling": "_scaling", "rotation": "_rotation"}          # optimizer group name -> model attribute
STAT_ATTRS = ("xyz_gradient_accum", "denom", "max_radii2D")
SKIP_GROUPS = ("vertex",)                                               # mesh.v: not a per-Gaussian table


def inverse_sigmoid(x: torch.Tensor) -> torch.Tensor:
    return torch.log(x / (1 - x))                                        # utils/general_utils.py:22-23


def rotation_matrices(q: torch.Tensor) -> torch.Tensor:
    """[N,4] wxyz (normalised here) -> [N,3,3]; entries of utils/general_utils.py:88-110 (build_rotation)."""
    q = q / q.norm(dim=1, keepdim=True)
    r, x, y, z = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1).reshape(-1, 3, 3)


class _Rows:
    """The per-Gaussian table of `model`: parameters (with their Adam moments), binding, statistics."""

    def __init__(self, model):
        self.m = model

    # -- optimizer plumbing ---------------------------------------------------------------------------
    def _groups(self):
        opt = getattr(self.m, "optimizer", None)
        if opt is None:
            return {}
        return {g["name"]: g for g in opt.param_groups if g.get("name") in PARAM_ATTRS and g.get("name") not in SKIP_GROUPS}

    def _swap(self, name: str, new_value: torch.Tensor, moments):
        """Install `new_value` as the parameter of group `name`; `moments(old_state_tensor)` maps each Adam moment."""
        attr = PARAM_ATTRS[name]
        groups = self._groups()
        new_param = nn.Parameter(new_value.requires_grad_(True))
        if name in groups:
            g = groups[name]
            opt = self.m.optimizer
            old = g["params"][0]
            state = opt.state.pop(old, None)
            if state is not None:
                for k in ("exp_avg", "exp_avg_sq"):
                    if k in state:
                        state[k] = moments(state[k])
                opt.state[new_param] = state
            g["params"][0] = new_param
        setattr(self.m, attr, new_param)

    # -- row operations -------------------------------------------------------------------------------
    def keep(self, keep_mask: torch.Tensor):
        for name, attr in PARAM_ATTRS.items():
            cur = getattr(self.m, attr)
            self._swap(name, cur.detach()[keep_mask], lambda s: s[keep_mask])
        for attr in STAT_ATTRS:
            if getattr(self.m, attr, None) is not None and torch.is_tensor(getattr(self.m, attr)) and getattr(self.m, attr).numel():
                setattr(self.m, attr, getattr(self.m, attr)[keep_mask])

    def append(self, new: Dict[str, torch.Tensor]):
        for name, attr in PARAM_ATTRS.items():
            cur, ext = getattr(self.m, attr), new[name]
            self._swap(name, torch.cat((cur.detach(), ext.detach()), dim=0),
                       lambda s, ext=ext: torch.cat((s, torch.zeros_like(ext)), dim=0))
        n = self.m._xyz.shape[0]
        dev = self.m._xyz.device
        # the reference restarts the statistics after every densification (scene/gaussian_model.py:343-345)
        self.m.xyz_gradient_accum = torch.zeros((n, 1), device=dev)
        self.m.denom = torch.zeros((n, 1), device=dev)
        self.m.max_radii2D = torch.zeros((n,), device=dev)


def _has_binding(model) -> bool:
    return getattr(model, "binding", None) is not None


def _bind_new(model, new_binding: torch.Tensor):
    model.binding = torch.cat((model.binding, new_binding))
    model.binding_counter.scatter_add_(0, new_binding, torch.ones_like(new_binding, dtype=model.binding_counter.dtype))


# --------------------------------------------------------------------------------------------------------
def add_densification_stats(model, viewspace_point_tensor: torch.Tensor, update_filter: torch.Tensor):
    """Accumulate |d loss / d screen-space mean| of the Gaussians that were visible in this view; the gradient is the
    rasterizer's means2D side channel (GaussianRasterizer backward; scene/gaussian_model.py:410-412)."""
    g = viewspace_point_tensor.grad if viewspace_point_tensor.grad is not None else viewspace_point_tensor
    model.xyz_gradient_accum[update_filter] += torch.norm(g[update_filter, :2], dim=-1, keepdim=True)
    model.denom[update_filter] += 1


def prune_points(model, mask: torch.Tensor):
    """Remove the rows where `mask` is True -- except that a mesh face never loses its last Gaussian: among the rows
    marked on a face, they are only removed if at least one unmarked row stays bound to it."""
    mask = mask.clone()
    if _has_binding(model):
        marked_faces = model.binding[mask]
        marked_per_face = torch.zeros_like(model.binding_counter)
        marked_per_face.scatter_add_(0, marked_faces, torch.ones_like(marked_faces, dtype=model.binding_counter.dtype))
        face_keeps_one = (model.binding_counter - marked_per_face) > 0
        mask[mask.clone()] = face_keeps_one[marked_faces]
    keep = ~mask
    if _has_binding(model):
        gone = model.binding[mask]
        model.binding_counter.scatter_add_(0, gone, -torch.ones_like(gone, dtype=model.binding_counter.dtype))
    _Rows(model).keep(keep)
    if _has_binding(model):
        model.binding = model.binding[keep]
    return mask


def densify_and_clone(model, grads: torch.Tensor, grad_threshold: float, scene_extent: float):
    sel = (torch.norm(grads, dim=-1) >= grad_threshold) & \
          (model.get_scaling.max(dim=1).values <= model.percent_dense * scene_extent)
    new = {name: getattr(model, attr).detach()[sel] for name, attr in PARAM_ATTRS.items()}
    if _has_binding(model):
        _bind_new(model, model.binding[sel])
    _Rows(model).append(new)
    return sel


def densify_and_split(model, grads: torch.Tensor, grad_threshold: float, scene_extent: float, N: int = 2,
                      generator: Optional[torch.Generator] = None):
    n0 = model.get_xyz.shape[0]
    dev = model._xyz.device
    padded = torch.zeros(n0, device=dev)
    padded[:grads.shape[0]] = grads.squeeze()
    world_scale = model.get_scaling.detach()
    sel = (padded >= grad_threshold) & (world_scale.max(dim=1).values > model.percent_dense * scene_extent)
    stds = world_scale[sel].repeat(N, 1)
    samples = torch.normal(mean=torch.zeros_like(stds), std=stds, generator=generator)
    R = rotation_matrices(model._rotation.detach()[sel]).repeat(N, 1, 1)
    new_xyz = torch.bmm(R, samples.unsqueeze(-1)).squeeze(-1) + model.get_xyz.detach()[sel].repeat(N, 1)
    if _has_binding(model):
        local_scale = world_scale[sel] / model.face_scaling.detach()[model.binding[sel]]      # back to the face frame
    else:
        local_scale = world_scale[sel]
    new = {"xyz": new_xyz, "scaling": torch.log(local_scale.repeat(N, 1) / (0.8 * N)),
           "rotation": model._rotation.detach()[sel].repeat(N, 1),
           "f_dc": model._features_dc.detach()[sel].repeat(N, 1, 1),
           "f_rest": model._features_rest.detach()[sel].repeat(N, 1, 1),
           "opacity": model._opacity.detach()[sel].repeat(N, 1)}
    if _has_binding(model):
        _bind_new(model, model.binding[sel].repeat(N))
    _Rows(model).append(new)
    grown = torch.cat((sel, torch.zeros(N * int(sel.sum()), device=dev, dtype=torch.bool)))
    prune_points(model, grown)
    return sel


def densify_and_prune(model, max_grad: float, min_opacity: float, extent: float, max_screen_size,
                      generator: Optional[torch.Generator] = None):
    grads = model.xyz_gradient_accum / model.denom
    grads[grads.isnan()] = 0.0
    densify_and_clone(model, grads, max_grad, extent)
    densify_and_split(model, grads, max_grad, extent, generator=generator)
    mask = (model.get_opacity < min_opacity).squeeze()
    if max_screen_size:
        mask = mask | (model.max_radii2D > max_screen_size) | (model.get_scaling.max(dim=1).values > 0.1 * extent)
    prune_points(model, mask)


def reset_opacity(model):
    cur = model.get_opacity.detach()
    new = inverse_sigmoid(torch.min(cur, torch.ones_like(cur) * 0.01))
    _Rows(model)._swap("opacity", new, lambda s: torch.zeros_like(new))


# --------------------------------------------------------------------------------------------------------
def rank_consistent_generator(seed: int, iteration: int, device) -> torch.Generator:
    """Same stream on every rank for the densification of `iteration` (ranks hold replicated parameters)."""
    g = torch.Generator(device=device)
    g.manual_seed((int(seed) * 1_000_003 + int(iteration)) % (2 ** 63 - 1))
    return g


def state_fingerprint(model) -> torch.Tensor:
    """[count, sum(binding), sum|xyz|, sum|scaling|] in float64 -- cheap to all-reduce, changes on any divergence."""
    b = model.binding.double().sum() if _has_binding(model) else torch.zeros((), dtype=torch.float64, device=model._xyz.device)
    return torch.stack([torch.tensor(float(model._xyz.shape[0]), dtype=torch.float64, device=model._xyz.device), b,
                        model._xyz.detach().double().abs().sum(), model._scaling.detach().double().abs().sum()])


def assert_rank_consistent(model, group=None):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    f = state_fingerprint(model)
    lo, hi = f.clone(), f.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    if not torch.equal(lo, hi):
        raise RuntimeError(f"gaussian-garments_b200: ranks diverged after densification (min {lo.tolist()}, max {hi.tolist()})")
