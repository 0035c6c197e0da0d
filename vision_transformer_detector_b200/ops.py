"""Operator-level entry points of the CUDA library (the Keras layers the path is built from), on
torch CUDA tensors.  Used by the parity tests to check each kernel against the oracle in isolation.
torch is only the tensor carrier here: every result comes from include/vitdet_b200.h functions."""
from __future__ import annotations

import ctypes as C

from . import _capi

ACT = {None: 0, "none": 0, "mish": 1, "gelu": 2}


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(t):
    import torch
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _f32c(t):
    import torch
    assert t.is_cuda, "operator entry points take CUDA tensors"
    return t.to(torch.float32).contiguous()


def dense(a, kernel, bias=None, resid=None, act=None, mode="bf16"):
    """keras.layers.Dense (+ activation + residual add): act(a @ kernel + bias) + resid."""
    import torch
    a, kernel = _f32c(a), _f32c(kernel)
    M, K = a.shape
    K2, N = kernel.shape
    assert K == K2
    bias = _f32c(bias) if bias is not None else None
    resid = _f32c(resid) if resid is not None else None
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _capi.check(_capi.load().vitdet_op_dense(_p(a), _p(kernel), _p(bias), _p(resid), _p(out), M, K, N, ACT[act],
                                                 _capi.MODES[mode], _stream(a)))
    return out


def layernorm(x, gamma, beta, eps=1e-3):
    """keras.layers.LayerNormalization(axis=-1)."""
    import torch
    x, gamma, beta = _f32c(x), _f32c(gamma), _f32c(beta)
    M, D = x.shape
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _capi.check(_capi.load().vitdet_op_layernorm(_p(x), _p(gamma), _p(beta), _p(y), M, D, float(eps), _stream(x)))
    return y


def attention(q, k, v, mode="bf16"):
    """softmax(q k^T / sqrt(d)) v per (image, head); q, k, v: [B, T, H, d]."""
    import torch
    q, k, v = _f32c(q), _f32c(k), _f32c(v)
    B, T, H, d = q.shape
    out = torch.empty_like(q)
    with torch.cuda.device(q.device):
        _capi.check(_capi.load().vitdet_op_attention(_p(q), _p(k), _p(v), _p(out), B, T, H, d, _capi.MODES[mode], _stream(q)))
    return out


def patchify(images, patch_size):
    """tf.image.extract_patches(SAME) + Reshape: [B,H,W,3] -> [B, T, 3 p^2]."""
    import torch
    images = _f32c(images)
    B, H, W, ch = images.shape
    assert ch == 3
    p = int(patch_size)
    T = (-(-H // p)) * (-(-W // p))
    out = torch.empty((B, T, 3 * p * p), dtype=torch.float32, device=images.device)
    with torch.cuda.device(images.device):
        _capi.check(_capi.load().vitdet_op_patchify(_p(images), B, H, W, p, _p(out), _stream(images)))
    return out
