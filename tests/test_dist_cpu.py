"""world_size-2 gloo test of the data-parallel host logic: batch sharding + one all-reduce(SUM) of the flat gradient
arena + 1/world prescale reproduces the gradient of the global-batch step (DDP-mean semantics, trainer.py:154-155,439),
using the CPU oracle for the per-rank forward/backward."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restatement as R
from open_musiclm_b200.dist_utils import BucketReducer, allreduce_sum_, grad_prescale, plan_buckets, rank_seed


def _cfg():
    return R.coarse_cfg(dim=64, depth=1, heads=2, codebook=32, n_clap_q=2, n_coarse_q=3, ce_weights=[0.0, 0.0, 1.0])


def _tokens(B, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randint(0, 32, s, generator=g).numpy() for s in [(B, 2), (B, 5), (B, 4, 3)]]


def _flat_grads(cfg, sd, toks):
    leaf = {k: (v.clone().requires_grad_(True) if not k.endswith("beta") else v) for k, v in sd.items()}
    loss = R.loss_and_logits(cfg, leaf, toks)[0]
    loss.backward()
    names = [k for k in leaf if not k.endswith("beta")]
    return torch.cat([(leaf[k].grad if leaf[k].grad is not None else torch.zeros_like(leaf[k])).reshape(-1) for k in names])


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg = _cfg()
    sd = R.init_state(cfg, seed=0)                     # same seed on every rank -> identical replicas
    toks = _tokens(4, 99)                              # global batch; rank r takes rows [2r, 2r+2)
    shard = [t[2 * rank:2 * rank + 2] for t in toks]
    flat = _flat_grads(cfg, sd, shard)
    allreduce_sum_(flat)
    flat *= grad_prescale()
    if rank == 0:
        torch.save(flat, out)
    dist.destroy_process_group()


def test_allreduce_sum_prescale_equals_global_batch_gradient(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    cfg = _cfg()
    ref = _flat_grads(cfg, R.init_state(cfg, seed=0), _tokens(4, 99))
    err = float((got - ref).norm() / ref.norm())
    assert err < 1e-5, err


def test_rank_seeds_differ_and_single_process_is_identity():
    assert rank_seed(0, 0) != rank_seed(0, 1) and rank_seed(3, 2) == rank_seed(3, 2)
    x = torch.arange(5.0)
    assert torch.equal(allreduce_sum_(x.clone()), x) and grad_prescale() == 1.0


def _toy_layout(depth=5):
    """Arena layout with the engine's ordering: [embeddings | logit heads | layer matrices | rel-pos || 1-D params], 64-aligned."""
    names = ["embeddings.0.weight", "embeddings.1.weight", "logit_weights.0", "logit_weights.1"]
    for l in range(depth):
        names += [f"transformer.layers.{l}.0.to_q.weight", f"transformer.layers.{l}.0.to_kv.weight", f"transformer.layers.{l}.0.to_out.0.weight",
                  f"transformer.layers.{l}.2.1.weight", f"transformer.layers.{l}.2.2.ds_conv.weight", f"transformer.layers.{l}.2.6.weight"]
    names += ["transformer.rel_pos_bias.net.0.0.weight", "transformer.rel_pos_bias.net.3.weight", "start_tokens.0", "start_tokens.1"]
    for l in range(depth):
        names += [f"transformer.layers.{l}.0.q_scale", f"transformer.layers.{l}.0.norm.gamma", f"transformer.layers.{l}.2.4.gamma"]
    names += ["transformer.rel_pos_bias.net.0.0.bias", "transformer.norm.gamma"]
    layout, sizes, off = {}, {}, 0
    for i, n in enumerate(names):
        sizes[n] = 37 + 101 * (i % 7)
        layout[n] = off
        off = (off + sizes[n] + 63) // 64 * 64
    return layout, sizes, off


def test_bucket_plan_tiles_the_arena_in_backward_order():
    layout, sizes, total = _toy_layout(5)
    plan = plan_buckets(layout, sizes, total, 5, min_elems=700)       # small buckets: layers get merged in pairs
    trig = [t for t, _ in plan]
    assert trig[0] == "heads" and trig[-1] == "tail" and trig[1:-1] == sorted(trig[1:-1], key=lambda t: -int(t[5:]))
    assert trig[-2] == "layer0"
    covered = sorted(s for _, sl in plan for s in sl)
    assert covered[0][0] == 0 and covered[-1][1] == total and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    # a layer's matrices are reduced no earlier than that layer's own trigger
    for l in range(5):
        first = layout[f"transformer.layers.{l}.0.to_q.weight"]
        owner = next(t for t, sl in plan if any(lo <= first < hi for lo, hi in sl))
        assert owner.startswith("layer") and int(owner[5:]) <= l
    # every slice is a multiple of 64 elements (parameters start on 64-element boundaries, the padding behind the last one
    # of a span belongs to it): equal parts for any power-of-two world size up to 64 (reduce-scatter mode)
    assert all((hi - lo) % 64 == 0 for _, sl in plan for lo, hi in sl)
    one = plan_buckets(layout, sizes, total, 5, min_elems=1 << 30)    # everything merged: heads, one layer bucket, tail
    assert [t for t, _ in one] == ["heads", "layer0", "tail"]


def _bucket_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    layout, sizes, total = _toy_layout(4)
    plan = plan_buckets(layout, sizes, total, 4, min_elems=500)
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(total, generator=g)
    red = BucketReducer(flat, plan, None, side_stream=None)
    red.begin()
    red.fire("heads")
    for l in reversed(range(4)):
        red.fire(f"layer{l}")          # triggers of merged-away layers are ignored
    red.fire("tail")
    red.join()
    if rank == 0:
        torch.save(flat, out)
    dist.destroy_process_group()


def test_bucketed_allreduce_equals_one_allreduce(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "flat.pt")
    mp.spawn(_bucket_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    _, _, total = _toy_layout(4)
    ref = sum(torch.randn(total, generator=torch.Generator().manual_seed(100 + r)) for r in range(2))
    assert torch.equal(got, ref)


def _shard_worker(rank, world, port, out):
    """Sharded update on the toy arena: reduce-scatter buckets -> each rank owns its part of every slice -> a stand-in
    'optimiser' (p -= 0.1 g) on the parts only -> all-gather of the parameters."""
    from open_musiclm_b200.dist_utils import all_gather_, shard_of
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    layout, sizes, total = _toy_layout(4)
    plan = plan_buckets(layout, sizes, total, 4, min_elems=500)
    flat = torch.randn(total, generator=torch.Generator().manual_seed(100 + rank))
    params = torch.arange(total, dtype=torch.float32) / total          # identical on every rank
    red = BucketReducer(flat, plan, None, side_stream=None, scatter=True)
    red.begin()
    red.fire("heads")
    for l in reversed(range(4)):
        red.fire(f"layer{l}")
    red.fire("tail")
    red.join()
    parts = [shard_of(lo, hi, world, rank) for lo, hi in red.slices()]
    assert sum(b - a for a, b in parts) * world == total               # the parts of all ranks tile the arena
    for a, b in parts:
        params[a:b] -= 0.1 * flat[a:b]
    for lo, hi in red.slices():
        all_gather_(params[lo:hi], world, rank)
    torch.save(params, out + f".{rank}")
    dist.destroy_process_group()


def test_sharded_update_equals_replicated_update(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "p.pt")
    mp.spawn(_shard_worker, args=(2, port, out), nprocs=2, join=True)
    _, _, total = _toy_layout(4)
    gsum = sum(torch.randn(total, generator=torch.Generator().manual_seed(100 + r)) for r in range(2))
    ref = torch.arange(total, dtype=torch.float32) / total - 0.1 * gsum
    p0, p1 = torch.load(out + ".0"), torch.load(out + ".1")
    assert torch.equal(p0, p1) and torch.equal(p0, ref)
