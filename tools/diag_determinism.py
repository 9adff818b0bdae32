"""Run the training-layout forward twice on the same input and report the first activation buffer that differs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import open_musiclm_b200 as O
torch.manual_seed(0)
m = O.create_coarse_transformer(dim=1024, depth=6, heads=8, num_coarse_quantizers=3, attn_dropout=0.0, ff_dropout=0.1).cuda().eval()
tr = O.HotPathTrainer(m, cross_entropy_loss_weights=[0.0, 0.0, 1.0], use_cuda_graph=False)
g = torch.Generator().manual_seed(1234)
toks = [torch.randint(0, 1024, (16, 12), generator=g).cuda(), torch.randint(0, 1024, (16, 197), generator=g).cuda(), torch.randint(0, 1024, (16, 270, 3), generator=g).cuda()]
eng = tr.eng
def snap():
    tr.eng.arena_g.zero_()
    tr._micro_batch(toks, False, 0, True)
    torch.cuda.synchronize()
    ws = next(iter(eng._ws.values())) if hasattr(eng, "_ws") else None
    return ws
ws = snap()
if ws is None:
    cands = [v for v in vars(eng).values() if isinstance(v, dict)]
    print("engine dict attrs:", [k for k, v in vars(eng).items() if isinstance(v, dict)])
    sys.exit(0)
keys = ["table", "x", "xn", "xraw", "q_raw", "kv_raw", "qn", "kvn", "o", "lse", "xn2", "u", "hn", "st_i", "logits"]
def grab():
    out = {}
    for k in keys:
        v = ws.get(k)
        if v is None: continue
        out[k] = [t.clone() for t in v] if isinstance(v, list) else [v.clone()]
    return out
a = grab()
snap()
b = grab()
for k in keys:
    if k not in a: continue
    for i, (p, q) in enumerate(zip(a[k], b[k])):
        d = (p.double() - q.double()).abs().max().item()
        n = (p != q).float().mean().item()
        if d > 0:
            print(f"{k}[{i}] max abs diff {d:.4g}  frac differing {n:.4g}  (max |value| {p.double().abs().max().item():.4g})")
print("done")
