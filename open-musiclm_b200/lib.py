"""ctypes binding of libomlm_b200.so (the C ABI declared in include/omlm_b200.h).

There is no fallback: if the shared library is missing or a call fails, this module raises.
PyTorch is used only for device memory and streams; every tensor is passed as a raw pointer.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libomlm_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "omlm_b200.h")

_lib = None


class OmlmError(RuntimeError):
    pass


def header_symbols():
    """Every function name declared in include/omlm_b200.h."""
    with open(HEADER_PATH) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(omlm_[a-z0-9_]+)\s*\(", src)))


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OmlmError(
                f"{LIB_PATH} not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.omlm_last_error.restype = ctypes.c_char_p
    return _lib


def _check(rc, name):
    if rc != 0:
        msg = load().omlm_last_error().decode(errors="replace")
        raise OmlmError(f"{name} failed (rc={rc}): {msg}")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


_I = ctypes.c_int
_L = ctypes.c_long
_F = ctypes.c_float


def call(name, *args):
    """Call an int-returning entry point; raise with omlm_last_error() on failure."""
    fn = getattr(load(), name)
    rc = fn(*args)
    _check(rc, name)


def device_check():
    call("omlm_device_check")


def gemm(a, b, out, *, a_mn=False, b_mn=False, M=None, N=None, K=None, addend=None, alpha=1.0,
         splits=1, row_split=0, row_valid=0, n_valid=0, block_n=128, max_ctas=0):
    """out[m,n] = alpha * sum_k A(m,k) B(n,k) (+ addend).  a: [M,K] (or [K,M] if a_mn); b: [N,K] (or [K,N] if b_mn)."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.stride(-1) == 1 and b.stride(-1) == 1 and out.stride(-1) == 1
    if M is None:
        M = a.shape[1] if a_mn else a.shape[0]
    if K is None:
        K = a.shape[0] if a_mn else a.shape[1]
    if N is None:
        N = b.shape[1] if b_mn else b.shape[0]
    out_f32 = 1 if out.dtype == torch.float32 else 0
    assert out_f32 or out.dtype == torch.bfloat16
    call("omlm_gemm_bf16", _p(a), _I(int(a_mn)), _L(a.stride(0)), _p(b), _I(int(b_mn)), _L(b.stride(0)),
         _I(M), _I(N), _I(K), _p(out), _I(out_f32), _L(out.stride(0)),
         _p(addend), _L(addend.stride(0) if addend is not None else 0), _F(alpha), _I(splits),
         _I(row_split), _I(row_valid), _I(n_valid), _I(block_n), _I(max_ctas), _stream())
    return out
