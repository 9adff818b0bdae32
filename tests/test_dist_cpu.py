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
from open_musiclm_b200.dist_utils import allreduce_sum_, grad_prescale, rank_seed


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
