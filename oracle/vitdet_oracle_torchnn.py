"""CPU ORACLE, second restatement — test infrastructure, not product code.

The same forward pass as vitdet_oracle.forward (reference det.py:239-495), written a second time against a DIFFERENT
set of primitives so that a misreading of one library semantic in the first restatement does not silently repeat here:

    tf.image.extract_patches(SAME) + Reshape      -> F.pad + F.unfold on NCHW (channel-major columns, re-ordered)
    keras.layers.MultiHeadAttention                -> torch.nn.MultiheadAttention with its packed in-projection
    keras.layers.LayerNormalization                -> torch.nn.LayerNorm(eps=1e-3)
    tfa.activations.mish / tfa.layers.GELU()       -> F.mish / F.gelu(approximate="tanh")
    keras.layers.Dense                             -> torch.nn.Linear

No einsum strings, no hand-written softmax.  tests/test_oracle.py requires it to agree with vitdet_oracle.forward to
1e-9 in float64.  PARITY UNPINNED against TensorFlow like the first restatement (TF is not installable here): what this
file adds is independence of the arithmetic, not of the reading of Keras' defaults (epsilon, q-scaling, SAME split).

torch.nn.MultiheadAttention projects embed_dim -> embed_dim, while Keras projects D -> H*key_dim -> D.  The Keras layer is
therefore embedded in a torch layer of embed_dim E = H*key_dim: the D-wide input is zero-padded to E columns, the packed
in_proj_weight (3E, E) holds the Keras q/k/v kernels in its first D columns, out_proj.weight (E, E) holds the Keras output
kernel in its first D rows, and the first D output columns are the Keras result.  torch scales q by head_dim**-0.5 after
the in-projection bias, which is where Keras applies 1/sqrt(key_dim) (det.py:364-369 semantics).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

SLOTS = 17


def _get(cfg, name):
    return cfg[name] if isinstance(cfg, dict) else getattr(cfg, name)


def _kname(base, i):
    return base if i == 0 else f"{base}_{i}"


def _linear(w, name, dtype):
    k = torch.as_tensor(np.asarray(w[name + "/kernel"]), dtype=dtype)        # keras (in, units)
    lin = torch.nn.Linear(k.shape[0], k.shape[1], dtype=dtype)
    with torch.no_grad():
        lin.weight.copy_(k.t())
        lin.bias.copy_(torch.as_tensor(np.asarray(w[name + "/bias"]), dtype=dtype))
    return lin


def _mha(w, name, D, H, d, dtype):
    E = H * d
    m = torch.nn.MultiheadAttention(E, H, bias=True, batch_first=True, dtype=dtype)
    with torch.no_grad():
        m.in_proj_weight.zero_(); m.in_proj_bias.zero_(); m.out_proj.weight.zero_(); m.out_proj.bias.zero_()
        for s, sel in enumerate(("query", "key", "value")):
            k = torch.as_tensor(np.asarray(w[f"{name}/{sel}/kernel"]), dtype=dtype).reshape(D, E)     # (D, H, d) -> (D, H*d)
            m.in_proj_weight[s * E:(s + 1) * E, :D] = k.t()
            m.in_proj_bias[s * E:(s + 1) * E] = torch.as_tensor(np.asarray(w[f"{name}/{sel}/bias"]), dtype=dtype).reshape(E)
        ko = torch.as_tensor(np.asarray(w[f"{name}/attention_output/kernel"]), dtype=dtype).reshape(E, D)   # (H, d, D) -> (H*d, D)
        m.out_proj.weight[:D, :] = ko.t()
        m.out_proj.bias[:D] = torch.as_tensor(np.asarray(w[f"{name}/attention_output/bias"]), dtype=dtype)
    return m


def _ln(w, name, D, dtype):
    ln = torch.nn.LayerNorm(D, eps=1e-3, dtype=dtype)
    with torch.no_grad():
        ln.weight.copy_(torch.as_tensor(np.asarray(w[name + "/gamma"]), dtype=dtype))
        ln.bias.copy_(torch.as_tensor(np.asarray(w[name + "/beta"]), dtype=dtype))
    return ln


def forward(weights: dict, cfg, images, dtype=torch.float64) -> np.ndarray:
    p = int(_get(cfg, "patch_size")); D = int(_get(cfg, "embedding_dim"))
    H, d = int(_get(cfg, "encoder_num_heads")), int(_get(cfg, "encoder_key_dim"))
    L, qn = int(_get(cfg, "encoder_repeat_times")), int(_get(cfg, "encoder_mlp_quantities"))
    n_head = int(_get(cfg, "mlp_head_dense_layers_quantity")) * int(_get(cfg, "mlp_head_dense_mish_block_repeats"))
    act = F.mish if _get(cfg, "use_mish") else (lambda t: F.gelu(t, approximate="tanh"))
    with torch.no_grad():
        img = torch.as_tensor(np.asarray(images), dtype=dtype).permute(0, 3, 1, 2)          # NHWC -> NCHW
        B, C, Hi, Wi = img.shape
        gh, gw = math.ceil(Hi / p), math.ceil(Wi / p)
        ph, pw = gh * p - Hi, gw * p - Wi
        img = F.pad(img, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))                    # SAME: the smaller half first
        cols = F.unfold(img, kernel_size=p, stride=p)                                       # (B, C*p*p, T), rows ordered (c, r, s)
        T = cols.shape[-1]
        patches = cols.reshape(B, C, p, p, T).permute(0, 4, 2, 3, 1).reshape(B, T, p * p * C)   # -> (row, col, channel) per token
        x = _linear(weights, "linear_projection", dtype)(patches)
        x = x + torch.as_tensor(np.asarray(weights["position_encoding/position_embedding/embeddings"]), dtype=dtype).reshape(1, T, 1)
        for i in range(L):
            y = _ln(weights, _kname("layer_normalization", 2 * i), D, dtype)(x)
            m = _mha(weights, _kname("multi_head_attention", i), D, H, d, dtype)
            ypad = F.pad(y, (0, H * d - D))
            att, _ = m(ypad, ypad, ypad, need_weights=False)
            x = x + att[..., :D]
            y = _ln(weights, _kname("layer_normalization", 2 * i + 1), D, dtype)(x)
            for j in range(qn):
                y = act(_linear(weights, f"MLP_{i + 1}_{j + 1}", dtype)(y))
            x = x + y
        y = _linear(weights, "dense", dtype)(x)                      # (B, T, 17)
        y = y.contiguous().view(B, SLOTS, T)                         # Reshape((17, -1)): same memory, new shape
        for k in range(1, n_head + 1):
            y = act(_linear(weights, _kname("dense", k), dtype)(y))
        return _linear(weights, "MLP_Head_no_Sigmoid", dtype)(y).numpy()
