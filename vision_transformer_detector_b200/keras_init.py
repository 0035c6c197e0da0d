"""Keras default initialisers, restated without TensorFlow, so that
`create_vision_transformer_detector()` returns a random-initialised model exactly as the reference's
Keras layers would produce one (same distributions; the random streams differ, of course).

  Dense / EinsumDense kernels : GlorotUniform, limit = sqrt(6 / (fan_in + fan_out)); for rank > 2
      kernels Keras' `_compute_fans` uses receptive_field = prod(shape[:-2]),
      fan_in = shape[-2] * rf, fan_out = shape[-1] * rf  (MHA q/k/v (D,H,d); output (H,d,D))
  biases                      : zeros
  LayerNormalization          : gamma = 1, beta = 0
  Embedding                   : uniform(-0.05, 0.05)      (PositionEncoding, det.py:148-151)
"""
from __future__ import annotations

import numpy as np


def compute_fans(shape) -> tuple[float, float]:
    shape = tuple(int(s) for s in shape)
    if len(shape) < 1:
        return 1.0, 1.0
    if len(shape) == 1:
        return float(shape[0]), float(shape[0])
    if len(shape) == 2:
        return float(shape[0]), float(shape[1])
    rf = 1
    for s in shape[:-2]:
        rf *= s
    return float(shape[-2] * rf), float(shape[-1] * rf)


def glorot_uniform(rng: np.random.Generator, shape) -> np.ndarray:
    fan_in, fan_out = compute_fans(shape)
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


def init_weight(rng: np.random.Generator, name: str, shape, *, spread: bool = False) -> np.ndarray:
    """Initial value of the Keras variable `name` (no ':0' suffix) of the given shape.

    spread=True gives the "set B" used by the parity tests and the bench: the Keras-default kernels,
    but non-trivial biases U(-0.1, 0.1), LayerNorm gamma U(0.5, 1.5) / beta U(-0.2, 0.2), and the last
    Dense ('MLP_Head_no_Sigmoid') scaled x25 with bias U(-1, 1), so that every parameter kind takes part
    in the arithmetic and the logits leave the sigmoid ~ 0.5 knife-edge that the default initialisation
    sits on (|logit| ~ 0.05 there) without saturating."""
    leaf = name.rsplit("/", 1)[-1]
    last = name.startswith("MLP_Head_no_Sigmoid/")
    if leaf == "kernel":
        w = glorot_uniform(rng, shape)
        return (w * 25.0).astype(np.float32) if (spread and last) else w
    if leaf == "bias":
        if spread:
            lim = 1.0 if last else 0.1
            return rng.uniform(-lim, lim, size=shape).astype(np.float32)
        return np.zeros(shape, np.float32)
    if leaf == "gamma":
        return rng.uniform(0.5, 1.5, size=shape).astype(np.float32) if spread else np.ones(shape, np.float32)
    if leaf == "beta":
        return rng.uniform(-0.2, 0.2, size=shape).astype(np.float32) if spread else np.zeros(shape, np.float32)
    if leaf == "embeddings":
        return rng.uniform(-0.05, 0.05, size=shape).astype(np.float32)
    raise ValueError(f"unknown variable kind: {name}")
