"""Row N4 (densify / prune / Adam-state surgery / rank-consistent RNG / PLY + binding.pkl IO), CPU side.

tests/golden/densify.npz was produced by the REFERENCE'S OWN code (MeshGaussianModel.densify_and_prune, prune_points,
reset_opacity, add_densification_stats: tests/golden/make_densify_golden.py); the replay below goes through
gaussian-garments_b200/densify.py and must reproduce every tensor -- parameters, Adam moments, binding,
binding_counter, statistics -- exactly.
"""
import os
import socket
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as h
from gaussian_garments_b200 import densify as D
from gaussian_garments_b200 import ply_io

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity",
         "scaling": "_scaling", "rotation": "_rotation"}


class Model:
    """Attribute surface of MeshGaussianModel that densification touches (test double; the product functions are
    duck-typed and work on the reference's own model objects)."""

    def __init__(self, z, tag="s0", device="cpu"):
        T = lambda k: torch.tensor(z[k]).to(device)
        P = torch.nn.Parameter
        for k, a in NAMES.items():
            setattr(self, a, P(T(f"{tag}__{k}")))
        self.mesh = types.SimpleNamespace(v=P(T(f"{tag}__vertex")), f=T("mesh_f"))
        self.binding, self.binding_counter = T(f"{tag}__binding"), T(f"{tag}__binding_counter")
        self.face_scaling, self.face_center, self.face_orien_mat = T("face_scaling"), T("face_center"), T("face_orien_mat")
        self.xyz_gradient_accum, self.denom = T(f"{tag}__xyz_gradient_accum"), T(f"{tag}__denom")
        self.max_radii2D = T(f"{tag}__max_radii2D")
        self.percent_dense = float(z["percent_dense"])
        groups = [{"params": [getattr(self, a)], "lr": 1e-3, "name": k} for k, a in NAMES.items()]
        groups.append({"params": [self.mesh.v], "lr": 1e-4, "name": "vertex"})
        self.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
        for k, a in NAMES.items():
            p = getattr(self, a)
            self.optimizer.state[p] = {"step": torch.tensor(1.0), "exp_avg": T(f"{tag}__{k}__exp_avg"),
                                       "exp_avg_sq": T(f"{tag}__{k}__exp_avg_sq")}
        self.optimizer.state[self.mesh.v] = {"step": torch.tensor(1.0), "exp_avg": T(f"{tag}__vertex__exp_avg"),
                                             "exp_avg_sq": T(f"{tag}__vertex__exp_avg").square()}

    get_opacity = property(lambda s: torch.sigmoid(s._opacity))
    get_scaling = property(lambda s: torch.exp(s._scaling) * s.face_scaling[s.binding])
    get_xyz = property(lambda s: torch.bmm(s.face_orien_mat[s.binding], s._xyz[..., None]).squeeze(-1)
                       * s.face_scaling[s.binding] + s.face_center[s.binding])


def _check(m, z, tag):
    for k, a in NAMES.items():
        p = getattr(m, a)
        assert torch.equal(p.detach().cpu(), torch.tensor(z[f"{tag}__{k}"])), (tag, k)
        st = m.optimizer.state[p]
        assert torch.equal(st["exp_avg"].cpu(), torch.tensor(z[f"{tag}__{k}__exp_avg"])), (tag, k, "exp_avg")
        assert torch.equal(st["exp_avg_sq"].cpu(), torch.tensor(z[f"{tag}__{k}__exp_avg_sq"])), (tag, k, "exp_avg_sq")
        assert p.requires_grad and any(g["params"][0] is p for g in m.optimizer.param_groups if g["name"] == k)
    assert torch.equal(m.binding.cpu(), torch.tensor(z[f"{tag}__binding"]))
    assert torch.equal(m.binding_counter.cpu(), torch.tensor(z[f"{tag}__binding_counter"]))
    for a in ("xyz_gradient_accum", "denom", "max_radii2D"):
        assert torch.equal(getattr(m, a).cpu(), torch.tensor(z[f"{tag}__{a}"])), (tag, a)
    # the mesh's "vertex" group is never touched
    assert torch.equal(m.mesh.v.detach().cpu(), torch.tensor(z[f"{tag}__vertex"]))
    assert torch.equal(m.optimizer.state[m.mesh.v]["exp_avg"].cpu(), torch.tensor(z[f"{tag}__vertex__exp_avg"]))


def _replay(m, z):
    N = m._xyz.shape[0]
    dev = m._xyz.device
    for k in range(2):
        vs = torch.zeros(N, 3, device=dev)
        vs.grad = torch.tensor(z[f"view{k}_grad"]).to(dev)
        D.add_densification_stats(m, vs, torch.tensor(z[f"view{k}_filter"]).to(dev))
    m.max_radii2D = torch.tensor(z["max_radii2D_in"]).to(dev)
    yield "s1"
    mg, mo, ext, mss = (float(v) for v in z["dp_args"])
    torch.manual_seed(int(z["split_seed"]))
    D.densify_and_prune(m, mg, mo, ext, int(mss))
    yield "s2"
    D.prune_points(m, torch.tensor(z["prune_mask"]).to(dev))
    yield "s3"
    D.reset_opacity(m)
    yield "s4"


def test_replay_equals_the_reference_models_own_densification():
    z = np.load(os.path.join(HERE, "golden", "densify.npz"))
    m = Model(z)
    sizes = {}
    for tag in _replay(m, z):
        _check(m, z, tag)
        sizes[tag] = m._xyz.shape[0]
    assert sizes["s2"] > sizes["s1"] > 0 and sizes["s3"] < sizes["s2"]          # it did grow, then shrink
    assert int(m.binding_counter.min()) >= 1                                      # no face was stripped bare
    assert int(m.binding_counter.sum()) == m._xyz.shape[0]


def test_prune_never_takes_the_last_gaussian_of_a_face():
    z = np.load(os.path.join(HERE, "golden", "densify.npz"))
    m = Model(z)
    everything = torch.ones(m._xyz.shape[0], dtype=torch.bool)
    D.prune_points(m, everything)                     # ask for all of them: every face keeps its Gaussians
    assert m._xyz.shape[0] == int(z["s0__xyz"].shape[0]) and int(m.binding_counter.min()) >= 1
    some = torch.zeros(m._xyz.shape[0], dtype=torch.bool)
    some[::3] = True                                  # one of three per face: removable
    D.prune_points(m, some)
    assert m._xyz.shape[0] == int(z["s0__xyz"].shape[0]) * 2 // 3
    assert torch.equal(m.binding_counter, torch.full_like(m.binding_counter, 2))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.load(os.path.join(HERE, "golden", "densify.npz"))
    m = Model(z)
    torch.manual_seed(1234 + rank)                    # ranks' GLOBAL generators differ on purpose
    N = m._xyz.shape[0]
    # per-rank view statistics are all-reduced before the decision (replicated parameters, sharded views)
    vs = torch.zeros(N, 3)
    vs.grad = torch.randn(N, 3) * 0.002
    D.add_densification_stats(m, vs, torch.rand(N) > 0.3)
    dist.all_reduce(m.xyz_gradient_accum)
    dist.all_reduce(m.denom)
    gen = D.rank_consistent_generator(seed=31359, iteration=500, device="cpu")
    D.densify_and_prune(m, 0.0002, 0.05, 2.0, 20, generator=gen)
    D.assert_rank_consistent(m)
    torch.save({k: getattr(m, a).detach() for k, a in NAMES.items()} | {"binding": m.binding}, os.path.join(out_dir, f"r{rank}.pt"))
    if rank == 1:                                     # a rank that drew from its own generator would be caught
        m2 = Model(z)
        m2._xyz.data[0, 0] += 1.0
    else:
        m2 = Model(z)
    try:
        D.assert_rank_consistent(m2)
        ok = False
    except RuntimeError:
        ok = True
    assert ok
    dist.destroy_process_group()


def test_two_ranks_take_identical_densification_decisions(tmp_path):
    port = _free_port()
    mp.spawn(_rank_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert a["xyz"].shape[0] != 180                   # something was actually densified


def test_ply_and_binding_pkl_roundtrip(tmp_path):
    """scene/mesh_gaussian_model.py:251-283 (save_ply, save_local=True -> local_point_cloud.ply + binding.pkl) and
    :289-342 (load_ply): same property list / order / dtype, readable back bit for bit."""
    z = np.load(os.path.join(HERE, "golden", "densify.npz"))
    m = Model(z, "s2")
    path = tmp_path / "frame_0001" / "local_point_cloud.ply"
    ply_io.save_ply(m, str(path), save_local=True)
    head = open(path, "rb").read(2048).split(b"end_header\n")[0].decode().splitlines()
    assert head[0] == "ply" and head[1] == "format binary_little_endian 1.0"
    assert f"element vertex {m._xyz.shape[0]}" in head
    props = [l.split()[-1] for l in head if l.startswith("property float")]
    expect = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(45)] + \
             ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]
    assert props == expect                                        # construct_list_of_attributes (gaussian_model.py:176-191)
    assert os.path.exists(tmp_path / "frame_0001" / "binding.pkl")
    m2 = types.SimpleNamespace(max_sh_degree=3, mesh=types.SimpleNamespace(v=m.mesh.v))
    ply_io.load_ply(m2, str(path), device="cpu")
    for a in ("_xyz", "_opacity", "_scaling", "_rotation", "_features_dc", "_features_rest"):
        assert torch.equal(getattr(m2, a), getattr(m, a).detach()), a
    assert torch.equal(m2.binding, m.binding) and m2.active_sh_degree == 3
    assert m2._features_dc.shape[1:] == (1, 3) and m2._features_rest.shape[1:] == (15, 3)
    # world-frame export (save_local=False): xyz / log-scale / rotation are the WORLD quantities the facade renders
    wpath = tmp_path / "frame_0001" / "point_cloud.ply"
    world = dict(xyz=m.get_xyz.detach(), scaling=torch.log(m.get_scaling.detach()), rotation=m._rotation.detach())
    ply_io.save_ply(m, str(wpath), save_local=False, world=world)
    data = ply_io.read_ply(str(wpath))
    assert np.array_equal(np.stack([data["x"], data["y"], data["z"]], 1), world["xyz"].numpy())
    assert np.array_equal(data["scale_1"], world["scaling"][:, 1].numpy())
