import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from open_musiclm_b200 import lib
B, N, h = 16, 1024, 8
M = B * N
qn = F.normalize(torch.randn(M, h, 64, device="cuda"), dim=-1).reshape(M, h * 64).bfloat16()
kvn = torch.randn(M, 128, device="cuda").bfloat16()
table = (torch.randn(h, 1, device="cuda") * 0.05 * torch.arange(N, device="cuda")[None]).contiguous()
km = (torch.rand(B, N, device="cuda") > 0.15).to(torch.uint8); km[:, 0] = 1
out = torch.empty(M, h * 64, device="cuda", dtype=torch.bfloat16); lse = torch.empty(B, N * h, device="cuda")
lib.attn_fwd_tc(qn, kvn, table, km, out, lse, B, N, h)
d_o = torch.randn(M, h * 64, device="cuda").bfloat16()
dqn = torch.zeros(M, h * 64, device="cuda"); dkvn = torch.zeros(M, 128, device="cuda"); dtab = torch.zeros_like(table); dsum = torch.empty(M * h, device="cuda")
for _ in range(3): lib.attn_bwd(qn, kvn, d_o, out, lse, table, km, dsum, dqn, dkvn, dtab, B, N, h)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): lib.attn_bwd(qn, kvn, d_o, out, lse, table, km, dsum, dqn, dkvn, dtab, B, N, h)
e1.record(); torch.cuda.synchronize()
print(f"attn_bwd: {e0.elapsed_time(e1)/10*1000:.1f} us")
Ns = (N + 127) // 128 * 128
ds = torch.empty(B, N * h, Ns, device="cuda", dtype=torch.bfloat16)
for _ in range(3): lib.attn_bwd_tc(qn, kvn, d_o, out, lse, table, km, dsum, ds, dqn, dkvn, dtab, B, N, h)
e0.record()
for _ in range(10): lib.attn_bwd_tc(qn, kvn, d_o, out, lse, table, km, dsum, ds, dqn, dkvn, dtab, B, N, h)
e1.record(); torch.cuda.synchronize()
print(f"attn_bwd_tc: {e0.elapsed_time(e1)/10*1000:.1f} us")
