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


def dense_ex(a, kernel, bias=None, resid=None, act=None, pos=None, ln=None, store_bf16=False, pair=None):
    """The tensor-core Dense with the epilogue extras the forward pass uses: `pos` = per-row scalar (position embedding,
    period = len(pos)), `ln` = (gamma, beta, eps) for the fused LayerNorm of the output row (returns (out, ln_out)),
    `store_bf16` = the bf16-output epilogue (bias + activation on packed pairs, TMA store), `pair` = None (auto) /
    False (single-CTA kernel) / True (CTA-pair kernel)."""
    import torch
    a, kernel = _f32c(a), _f32c(kernel)
    M, K = a.shape
    K2, N = kernel.shape
    assert K == K2
    bias = _f32c(bias) if bias is not None else None
    resid = _f32c(resid) if resid is not None else None
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    ex = _capi.DenseEx()
    keep = []
    if pos is not None:
        pos = _f32c(pos).reshape(-1); keep.append(pos)
        ex.pos, ex.pos_period = pos.data_ptr(), pos.numel()
    ln_out = None
    if ln is not None:
        g, b, eps = ln
        g, b = _f32c(g), _f32c(b); keep += [g, b]
        ln_out = torch.empty((M, N), dtype=torch.float32, device=a.device)
        ex.ln_gamma, ex.ln_beta, ex.ln_eps, ex.ln_out = g.data_ptr(), b.data_ptr(), float(eps), ln_out.data_ptr()
    ex.store_bf16 = 1 if store_bf16 else 0
    ex.pair = -1 if pair is None else (1 if pair else 0)
    with torch.cuda.device(a.device):
        _capi.check(_capi.load().vitdet_op_dense_ex(_p(a), _p(kernel), _p(bias), _p(resid), _p(out), M, K, N, ACT[act], C.byref(ex), _stream(a)))
    return (out, ln_out) if ln is not None else out


def mlp_tail(a, layers, x, act="mish", ln=None):
    """The fused last three Dense+activation layers of an encoder block + residual (+ LayerNorm): layers = three
    (kernel, bias) pairs in Keras layout; x = residual stream [M, N2]; returns the new x (and LN(x) with ln=(gamma, beta, eps))."""
    import torch
    a = _f32c(a)
    (w0, b0), (w1, b1), (w2, b2) = [(_f32c(k), _f32c(b)) for k, b in layers]
    x = _f32c(x).clone()
    M, K0 = a.shape
    N0, N1, N2 = w0.shape[1], w1.shape[1], w2.shape[1]
    g = b = None
    eps = 1e-3
    ln_out = None
    if ln is not None:
        g, b, eps = _f32c(ln[0]), _f32c(ln[1]), float(ln[2])
        ln_out = torch.empty_like(x)
    with torch.cuda.device(a.device):
        _capi.check(_capi.load().vitdet_op_mlp_tail(_p(a), _p(w0), _p(b0), _p(w1), _p(b1), _p(w2), _p(b2), _p(x), _p(g), _p(b), eps,
                                                    _p(ln_out), M, K0, N0, N1, N2, ACT[act], _stream(a)))
    return (x, ln_out) if ln is not None else x


def head_slots(x, kernel, bias, mode="bf16"):
    """mlp_head first stage: Dense(D -> S) per token + Reshape((S, -1)).  x: [images, tokens, D] -> [images, S, tokens]."""
    import torch
    x, kernel, bias = _f32c(x), _f32c(kernel), _f32c(bias)
    B, T, D = x.shape
    S = kernel.shape[1]
    out = torch.empty((B, S, T), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _capi.check(_capi.load().vitdet_op_head_slots(_p(x), _p(kernel), _p(bias), _p(out), B, T, D, S, _capi.MODES[mode], _stream(x)))
    return out
