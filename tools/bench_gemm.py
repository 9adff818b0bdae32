"""Times every GEMM shape of the cfg2 training step (forward / dgrad / wgrad) for both tile widths."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from open_musiclm_b200 import lib
from open_musiclm_b200.engine import Engine as E
M, d, HD, Fp, F = 16384, 1024, 512, 2816, 2730
dev = "cuda"
def t(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
bf = lambda *s: torch.randn(*s, device=dev).bfloat16()
rows = []
# (name, kind, m, n, k)
shapes = [("q", "fwd", M, HD, d), ("kv", "fwd", M, 128, d), ("out", "fwd_res", M, d, HD), ("ffn_up", "fwd", M, 2 * Fp, d), ("ffn_down", "fwd_res", M, d, Fp),
          ("d_hn", "dgrad", M, Fp, d), ("d_xn2", "dgrad", M, d, 2 * Fp), ("d_o", "dgrad", M, HD, d), ("d_xn(q)", "dgrad", M, d, HD), ("d_xraw", "dgrad", M, d, 128),
          ("dW2", "wgrad", d, Fp, M), ("dW1", "wgrad", 2 * Fp, d, M), ("dWo", "wgrad", d, HD, M), ("dWq", "wgrad", HD, d, M), ("dWkv", "wgrad", 128, d, M)]
total = {}
kinds = os.environ.get("BENCH_GEMM_KINDS", "fwd,fwd_res,dgrad,wgrad").split(",")
for name, kind, m, n, k in shapes:
    if kind not in kinds:
        continue
    fl = 2.0 * m * n * k
    if kind in ("fwd", "fwd_res"):
        a, b = bf(m, k), bf(n, k)
        out = torch.empty(m, n, device=dev, dtype=torch.float32 if kind == "fwd_res" else torch.bfloat16)
        res = torch.randn(m, n, device=dev) if kind == "fwd_res" else None
        for bn in (128, 256):
            ms = t(lambda: lib.gemm(a, b, out, addend=res, block_n=bn))
            rows.append((name, kind, m, n, k, f"bn={bn}", ms, fl / ms / 1e9))
    elif kind == "dgrad":
        a, b = bf(m, k), bf(k, n)
        out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
        for bn in (128, 256):
            ms = t(lambda: lib.gemm(a, b, out, b_mn=True, M=m, N=n, K=k, block_n=bn))
            rows.append((name, kind, m, n, k, f"bn={bn}", ms, fl / ms / 1e9))
    else:
        a, b = bf(k, m), bf(k, n)
        out = torch.zeros(m, n, device=dev)
        for bn in (128, 256):
            for s in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32):
                kb = (k + 63) // 64
                if s > kb // 8: continue
                if s == 1:
                    ms = t(lambda: lib.gemm(a, b, out, a_mn=True, b_mn=True, M=m, N=n, K=k, addend=out, block_n=bn))
                else:
                    ms = t(lambda: lib.gemm(a, b, out, a_mn=True, b_mn=True, M=m, N=n, K=k, splits=s, block_n=bn))
                rows.append((name, kind, m, n, k, f"bn={bn} s={s}", ms, fl / ms / 1e9))
best = {}
for r in rows:
    print(f"{r[0]:9s} {r[1]:8s} {r[2]:6d}x{r[3]:5d}x{r[4]:6d} {r[5]:14s} {r[6]*1000:8.1f} us {r[7]:7.0f} TF/s")
    if r[0] not in best or r[6] < best[r[0]][6]: best[r[0]] = r
print("---- best per shape, per-layer sum")
tot = 0
for k, r in best.items():
    print(f"{r[0]:9s} {r[5]:14s} {r[6]*1000:8.1f} us {r[7]:7.0f} TF/s")
    tot += r[6]
print(f"sum of best per layer: {tot*1000:.1f} us -> x6 layers = {tot*6:.2f} ms")
