"""Checkpoint interchange with the reference (row N4 of SURVEY.md 8f): the per-frame Gaussian PLY + `binding.pkl`.

The reference writes / reads these with `plyfile` (not in this image):
    MeshGaussianModel.save_ply(path, save_local)   scene/mesh_gaussian_model.py:251-283
    MeshGaussianModel.load_ply(path)               scene/mesh_gaussian_model.py:289-342
    GaussianModel.construct_list_of_attributes     scene/gaussian_model.py:176-191
The file is a standard binary little-endian PLY with ONE element `vertex` whose properties are all float32, in this
order: x y z nx ny nz f_dc_0..2 f_rest_0..(3(M-1)-1) opacity scale_0..2 rot_0..3 .  SH blocks are stored
channel-major (`[N,K,3].transpose(1,2).flatten(1)`), normals are zeros.  `binding.pkl` next to a local-frame PLY is the
pickled `binding` tensor.  Written here with numpy only, so files go both ways between the two code bases.
"""

from __future__ import annotations

import os
import pickle
from typing import Dict, Optional

import numpy as np
import torch


def attribute_names(n_dc: int, n_rest: int, n_scale: int = 3, n_rot: int = 4):
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(n_dc)] + [f"f_rest_{i}" for i in range(n_rest)] + ["opacity"]
    names += [f"scale_{i}" for i in range(n_scale)] + [f"rot_{i}" for i in range(n_rot)]
    return names


def write_ply(path: str, columns: np.ndarray, names):
    """columns [N, len(names)] float32 -> binary little-endian PLY, one `vertex` element."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    columns = np.ascontiguousarray(columns, dtype="<f4")
    assert columns.ndim == 2 and columns.shape[1] == len(names)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {columns.shape[0]}"]
    header += [f"property float {n}" for n in names] + ["end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(columns.tobytes())


def read_ply(path: str) -> Dict[str, np.ndarray]:
    """Reads the `vertex` element of a binary-little-endian or ascii PLY whose vertex properties are scalar."""
    with open(path, "rb") as f:
        raw = f.read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    lines = raw[:end].decode("ascii").splitlines()
    fmt = next(l.split()[1] for l in lines if l.startswith("format"))
    types = {"float": "f4", "float32": "f4", "double": "f8", "float64": "f8", "uchar": "u1", "uint8": "u1", "int": "i4",
             "int32": "i4", "uint": "u4", "short": "i2", "ushort": "u2", "char": "i1"}
    n, props, in_vertex = 0, [], False
    for l in lines:
        p = l.split()
        if p[:1] == ["element"]:
            in_vertex = p[1] == "vertex"
            if in_vertex:
                n = int(p[2])
        elif p[:1] == ["property"] and in_vertex:
            if p[1] == "list":
                raise ValueError("list properties on the vertex element are not supported")
            props.append((p[2], types[p[1]]))
    if fmt == "ascii":
        rows = np.loadtxt(raw[end:].decode("ascii").splitlines()[:n], ndmin=2)
        return {name: rows[:, i].astype(t) for i, (name, t) in enumerate(props)}
    order = "<" if fmt == "binary_little_endian" else ">"
    dt = np.dtype([(name, order + t) for name, t in props])
    arr = np.frombuffer(raw, dtype=dt, count=n, offset=end)
    return {name: np.ascontiguousarray(arr[name]) for name, _ in props}


def save_ply(model, path: str, save_local: bool = False, world: Optional[dict] = None, valid=None):
    """MeshGaussianModel.save_ply.  save_local=True writes the face-frame parameters (`_xyz`, `_scaling`, `_rotation`)
    and `binding.pkl` beside the file; otherwise the world-frame ones -- taken from `world` (dict xyz / scaling (log) /
    rotation, e.g. the outputs of FusedMeshBinding.world() with scaling logged) or from model.get_xyz / get_scaling /
    get_rotation.  `valid`: optional index list / mask of the Gaussians to keep (find_valid_gaussians)."""
    np_ = lambda t: t.detach().cpu().numpy()
    if save_local:
        xyz, scale, rot = np_(model._xyz), np_(model._scaling), np_(model._rotation)
    elif world is not None:
        xyz, scale, rot = np_(world["xyz"]), np_(world["scaling"]), np_(world["rotation"])
    else:
        xyz, scale, rot = np_(model.get_xyz), np_(torch.log(model.get_scaling)), np_(model.get_rotation)
    f_dc = np_(model._features_dc.detach().transpose(1, 2).flatten(start_dim=1).contiguous())
    f_rest = np_(model._features_rest.detach().transpose(1, 2).flatten(start_dim=1).contiguous())
    cols = np.concatenate((xyz, np.zeros_like(xyz), f_dc, f_rest, np_(model._opacity), scale, rot), axis=1)
    sel = slice(None) if valid is None else valid
    write_ply(path, cols[sel], attribute_names(f_dc.shape[1], f_rest.shape[1], scale.shape[1], rot.shape[1]))
    if save_local:
        with open(os.path.join(os.path.dirname(os.path.abspath(path)), "binding.pkl"), "wb") as f:
            pickle.dump(model.binding[sel], f)


def load_ply(model, path: str, device="cuda"):
    """MeshGaussianModel.load_ply: fills _xyz/_rotation/_features_dc/_features_rest/_opacity/_scaling (no grad),
    active_sh_degree, max_radii2D and -- when `binding.pkl` sits next to the file -- binding."""
    d = read_ply(path)
    col = lambda names: np.stack([np.asarray(d[n], dtype=np.float32) for n in names], axis=1)
    T = lambda a: torch.tensor(a, dtype=torch.float, device=device)
    rest = sorted((k for k in d if k.startswith("f_rest_")), key=lambda s: int(s.split("_")[-1]))
    assert len(rest) == 3 * (model.max_sh_degree + 1) ** 2 - 3, "SH degree of the file does not match the model"
    scales = sorted((k for k in d if k.startswith("scale_")), key=lambda s: int(s.split("_")[-1]))
    rots = sorted((k for k in d if k.startswith("rot")), key=lambda s: int(s.split("_")[-1]))
    n = d["x"].shape[0]
    model._xyz = T(col(["x", "y", "z"]))
    model._opacity = T(col(["opacity"]))
    model._features_dc = T(col(["f_dc_0", "f_dc_1", "f_dc_2"]).reshape(n, 3, 1)).transpose(1, 2).contiguous()
    model._features_rest = T(col(rest).reshape(n, 3, -1)).transpose(1, 2).contiguous()
    model._scaling = T(col(scales))
    model._rotation = T(col(rots))
    model.active_sh_degree = model.max_sh_degree
    model.max_radii2D = torch.zeros(n, device=device)
    bpath = os.path.join(os.path.dirname(os.path.abspath(path)), "binding.pkl")
    if os.path.exists(bpath):
        with open(bpath, "rb") as f:
            b = pickle.load(f)
        model.binding = (b if torch.is_tensor(b) else torch.as_tensor(np.asarray(b))).to(device)
    return model
