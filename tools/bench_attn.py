"""Runs the attention kernels at the cfg2 shape (for ncu captures):  python tools/bench_attn.py [fwd|bwd]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from open_musiclm_b200 import lib  # noqa: E402

B, N, h = 16, 1024, 8
M = B * N
torch.manual_seed(0)
qn = F.normalize(torch.randn(M, h, 64, device="cuda"), dim=-1).reshape(M, h * 64).bfloat16()
kvn = torch.randn(M, 128, device="cuda").bfloat16()
table = (torch.randn(h, 1, device="cuda") * 0.05 * torch.arange(N, device="cuda")[None]).contiguous()
key_mask = (torch.rand(B, N, device="cuda") > 0.15).to(torch.uint8)
key_mask[:, 0] = 1
out = torch.empty(M, h * 64, device="cuda", dtype=torch.bfloat16)
lse2 = torch.empty(B, N * h, device="cuda")
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
for _ in range(3):
    lib.attn_fwd_tc(qn, kvn, table, key_mask, out, lse2, B, N, h)
    if which == "bwd":
        d_o = torch.randn(M, h * 64, device="cuda").bfloat16()
        dqn = torch.zeros(M, h * 64, device="cuda"); dkvn = torch.zeros(M, 128, device="cuda")
        dtab = torch.zeros_like(table); dsum = torch.empty(M * h, device="cuda")
        lib.attn_bwd(qn, kvn, d_o, out, lse2, table, key_mask, dsum, dqn, dkvn, dtab, B, N, h)
torch.cuda.synchronize()
