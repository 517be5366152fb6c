"""CPU tests of the N>1 host logic: world_size-2 gloo run of the flat gradient bucket
(the rasterizer itself needs a GPU; here each rank's 'backward' is a deterministic stand-in that
writes into the bucket's sinks the same way gg_backward does)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as h

gg = h.gg
from gaussian_garments_b200.dist import GradBucket, shard_views  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_view_grads(params, view):
    g = torch.Generator().manual_seed(1000 + view)
    return [torch.randn(p.shape, generator=g) for p in params]


def _worker(rank, world, port, n_views, out_dir):
    sys.path.insert(0, h.ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.zeros(50, 3, requires_grad=True), torch.zeros(50, 16, 3, requires_grad=True),
              torch.zeros(50, 1, requires_grad=True)]
    bucket = GradBucket(params, world)
    steps = n_views // world
    mine = shard_views(n_views, rank, world)
    results = []
    for s in range(steps):
        bucket.zero()
        grads = _fake_view_grads(params, mine[s])
        for i, g in enumerate(grads):          # what backward does: write straight into the sink views,
            v = bucket.view(i)                 # which autograd's AccumulateGrad then adopts as .grad
            v.copy_(g)
            params[i].grad = v
        bucket.all_reduce()
        for i, p in enumerate(params):
            assert p.grad is not None and p.grad.data_ptr() == bucket.view(i).data_ptr()
        results.append(bucket.flat.clone())
    torch.save(results, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_bucket_allreduce_equals_single_rank_mean(tmp_path):
    world, n_views = 2, 4
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_views, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    params = [torch.zeros(50, 3), torch.zeros(50, 16, 3), torch.zeros(50, 1)]
    ref_bucket = GradBucket([p.requires_grad_(True) for p in params], 1, register=False)
    for s in range(n_views // world):
        assert torch.equal(r0[s], r1[s])                      # every rank holds the same reduced bucket
        views = [shard_views(n_views, r, world)[s] for r in range(world)]
        expect = torch.zeros_like(ref_bucket.flat)
        for v in views:
            for i, g in enumerate(_fake_view_grads(params, v)):
                o = ref_bucket.offsets[i]
                expect[o:o + g.numel()] += g.reshape(-1) / world
        assert torch.allclose(r0[s], expect, atol=1e-6)


def _split_worker(rank, world, port, out_dir):
    sys.path.insert(0, h.ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params = [torch.zeros(20, 3, requires_grad=True), torch.zeros(20, 16, 3, requires_grad=True),
              torch.zeros(20, 1, requires_grad=True)]
    bucket = GradBucket(params, world, deferred=(1,))          # the SH block is exchanged separately (side stream on GPU)
    assert bucket.impl == "process_group" and bucket.offsets[1] == bucket.split > bucket.offsets[2] > bucket.offsets[0] == 0
    bucket.zero()
    for i, g in enumerate(_fake_view_grads(params, rank)):
        bucket.view(i).copy_(g)
        params[i].grad = bucket.view(i)
    local = bucket.flat.clone()
    bucket.all_reduce()
    bucket.wait()
    chk = bucket.check_against_gather(local)
    assert chk["ok"] and chk["elements"] == bucket.numel, chk
    torch.save(bucket.flat.clone(), os.path.join(out_dir, f"split{rank}.pt"))
    dist.destroy_process_group()


def test_split_exchange_immediate_and_deferred_blocks_equal_the_mean(tmp_path):
    port = _free_port()
    mp.spawn(_split_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = torch.load(tmp_path / "split0.pt"), torch.load(tmp_path / "split1.pt")
    assert torch.equal(a, b)
    params = [torch.zeros(20, 3), torch.zeros(20, 16, 3), torch.zeros(20, 1)]
    ref = GradBucket([p.requires_grad_(True) for p in params], 1, register=False, deferred=(1,))
    expect = torch.zeros_like(ref.flat)
    for r in range(2):
        for i, g in enumerate(_fake_view_grads(params, r)):
            o = ref.offsets[i]
            expect[o:o + g.numel()] += g.reshape(-1) / 2
    assert torch.allclose(a, expect, atol=1e-6)


def test_shard_views_partitions_all_views():
    for world in (1, 2, 4, 8):
        seen = sorted(v for r in range(world) for v in shard_views(32, r, world))
        assert seen == list(range(32))


def test_bucket_views_are_aligned_and_disjoint():
    params = [torch.zeros(7, 3, requires_grad=True), torch.zeros(7, 16, 3, requires_grad=True)]
    b = GradBucket(params, 1, register=False)
    assert b.offsets[1] % 64 == 0 and b.offsets[1] >= 21
    b.view(0).fill_(1.0)
    b.view(1).fill_(2.0)
    assert float(b.view(0).sum()) == 21.0 and float(b.view(1).sum()) == 2.0 * 7 * 48


def test_bucket_adopt_copies_foreign_grads_and_zero_fills_missing():
    """adopt(): a .grad that autograd cloned (not a bucket view) is copied in; a missing .grad becomes zeros."""
    params = [torch.zeros(5, 3, requires_grad=True), torch.zeros(5, 1, requires_grad=True)]
    b = GradBucket(params, 2, register=False)
    b.flat.fill_(7.0)
    params[0].grad = torch.full((5, 3), 2.0)          # foreign tensor
    params[1].grad = None
    b.adopt()
    assert params[0].grad.data_ptr() == b.view(0).data_ptr() and float(b.view(0).sum()) == 30.0
    assert params[1].grad.data_ptr() == b.view(1).data_ptr() and float(b.view(1).abs().sum()) == 0.0


def test_grad_sink_registry_roundtrip():
    """Sinks live on the tensor OBJECT (never keyed by address: ADVICE r1), are usable once per arming and only
    for leaves, and disappear with unregister()."""
    from gaussian_garments_b200 import rasterizer
    params = [torch.zeros(4, 3, requires_grad=True)]
    b = GradBucket(params, 1)
    try:
        ref = rasterizer._sink_ref(params[0])
        assert ref is not None and ref() is params[0]
        ent = params[0]._gg_sink
        assert ent.flat is b.flat and ent.shape == (4, 3) and ent.armed is False
        b.zero()
        assert ent.armed is True
        assert rasterizer._sink_ref(torch.zeros(4, 3, requires_grad=True)) is None      # unknown tensor: no sink
        assert rasterizer._sink_ref(params[0][:2]) is None                              # a view is not the leaf
        # a new tensor that happens to reuse the parameter's storage address must not inherit the sink
        alias = torch.zeros(4, 3, requires_grad=True)
        alias.data = params[0].data
        assert alias.data_ptr() == params[0].data_ptr() and rasterizer._sink_ref(alias) is None
    finally:
        b.unregister()
    assert rasterizer._sink_ref(params[0]) is None
