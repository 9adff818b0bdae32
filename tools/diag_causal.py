"""Diagnostic: causality of the full-size coarse forward (perturb late tokens, report which logit positions move)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import open_musiclm_b200 as O
from oracle import restatement as R
torch.manual_seed(0)
m = O.create_coarse_transformer(dim=1024, depth=6, heads=8, num_coarse_quantizers=3, attn_dropout=0.0, ff_dropout=0.1).cuda().eval()
cfg = R.coarse_cfg(ce_weights=[0.0, 0.0, 1.0])
g = torch.Generator().manual_seed(1234)
clap, sem, coarse = (torch.randint(0, 1024, (16, 12), generator=g), torch.randint(0, 1024, (16, 197), generator=g), torch.randint(0, 1024, (16, 270, 3), generator=g))
def run(c, s, a):
    ids, mask, _ = R.prepare_ids(cfg, [c.numpy(), s.numpy(), a.numpy()], True, None)
    with torch.no_grad():
        out = m(all_token_ids=[torch.from_numpy(np.ascontiguousarray(i)).cuda() for i in ids], self_attn_mask=torch.from_numpy(mask).cuda(), return_only_final_seq_logits=True)
    return out[-1].float().clone()
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
base = run(clap, sem, coarse); base2 = run(clap, sem, coarse)
print("determinism rel", rel(base2, base))
pert = coarse.clone(); pert[:, 135:] = (pert[:, 135:] + 7) % 1024
moved = run(clap, sem, pert)
d = (moved - base).abs().amax(dim=(0, 2))          # per position
nz = torch.nonzero(d > 1e-3).flatten()
print("first positions with diff", nz[:20].tolist(), "count below 406:", int((nz < 406).sum()))
print("per-position max diff around cut", [(p, round(float(d[p]), 5)) for p in (0, 1, 100, 300, 380, 390, 400, 403, 404, 405, 406, 407, 410)])
per_b = (moved[:, :406] - base[:, :406]).abs().amax(dim=(1, 2))
print("per-batch max diff below cut", [round(float(x), 4) for x in per_b])
