"""One forward + one backward attention launch at the cfg2 shape (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from open_musiclm_b200 import lib
B, N, h = 16, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 8
M = B * N
torch.manual_seed(0)
qn = F.normalize(torch.randn(M, h, 64, device="cuda"), dim=-1).reshape(M, h * 64).bfloat16()
kvn = torch.randn(M, 128, device="cuda").bfloat16()
table = (torch.randn(h, 1, device="cuda") * 0.05 * torch.arange(N, device="cuda")[None]).contiguous()
km = (torch.rand(B, N, device="cuda") > 0.15).to(torch.uint8); km[:, 0] = 1
out = torch.empty(M, h * 64, device="cuda", dtype=torch.bfloat16); lse = torch.empty(B, N * h, device="cuda")
d_o = torch.randn(M, h * 64, device="cuda").bfloat16()
dqn = torch.zeros(M, h * 64, device="cuda"); dkvn = torch.zeros(M, 128, device="cuda"); dtab = torch.zeros_like(table)
dsum = torch.empty(M * h, device="cuda")
for _ in range(2):
    lib.attn_fwd_tc(qn, kvn, table, km, out, lse, B, N, h)
    lib.attn_bwd_tc(qn, kvn, d_o, out, lse, table, km, dsum, dqn, dkvn, dtab, B, N, h)
torch.cuda.synchronize()
