"""Runs one GEMM shape of the training step a few times (for `ncu -k regex:gemm_bf16_kernel`).
  python tools/gemm_once.py out|ffn_down|q|kv|d_hn|d_xn2|... [bn]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from open_musiclm_b200 import lib
M, d, HD, Fp = 16384, 1024, 512, 2816
shapes = {"q": ("fwd", M, HD, d), "kv": ("fwd", M, 128, d), "out": ("fwd_res", M, d, HD), "ffn_upg": ("fwd", M, 2 * Fp, d), "ffn_down": ("fwd_res", M, d, Fp),
          "d_hn": ("dgrad", M, Fp, d), "d_xn2": ("dgrad", M, d, 2 * Fp), "d_o": ("dgrad", M, HD, d), "d_xn": ("dgrad", M, d, HD), "d_xraw": ("dgrad", M, d, 128),
          "dW2": ("wgrad", d, Fp, M), "dW1": ("wgrad", 2 * Fp, d, M)}
name = sys.argv[1]
bn = int(sys.argv[2]) if len(sys.argv) > 2 else 256
splits = int(sys.argv[3]) if len(sys.argv) > 3 else 1
kind, m, n, k = shapes[name]
bf = lambda *s: torch.randn(*s, device="cuda").bfloat16()
if kind in ("fwd", "fwd_res"):
    a, b = bf(m, k), bf(n, k)
    out = torch.empty(m, n, device="cuda", dtype=torch.float32 if kind == "fwd_res" else torch.bfloat16)
    res = torch.randn(m, n, device="cuda") if kind == "fwd_res" else None
    fn = lambda: lib.gemm(a, b, out, addend=res, block_n=bn)
elif kind == "dgrad":
    a, b = bf(m, k), bf(k, n)
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    fn = lambda: lib.gemm(a, b, out, b_mn=True, M=m, N=n, K=k, block_n=bn)
else:
    a, b = bf(k, m), bf(k, n)
    out = torch.zeros(m, n, device="cuda")
    fn = (lambda: lib.gemm(a, b, out, a_mn=True, b_mn=True, M=m, N=n, K=k, splits=splits, block_n=bn)) if splits > 1 else \
         (lambda: lib.gemm(a, b, out, a_mn=True, b_mn=True, M=m, N=n, K=k, addend=out, block_n=bn))
for _ in range(3):
    fn()
torch.cuda.synchronize()
