"""Shared helpers of the test-suite (seeded inputs, tiny configurations, error metrics)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import vitdet_oracle as oracle  # noqa: E402  (test infrastructure)

TOL_FP32 = 1e-3     # north_star: <= 1e-3 relative in the fp32-accumulate mode
TOL_BF16 = 2e-2     # north_star: <= 2e-2 in bf16


def rel_err(got, ref) -> float:
    """max|got - ref| / max|ref| (SURVEY §8(d) tolerance bookkeeping)."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


def tiny_config(**over):
    from vision_transformer_detector_b200 import DetectorConfig
    kw = dict(input_shape=(60, 130, 3), patch_size=17, embedding_dim=28, encoder_num_heads=2, encoder_key_dim=40,
              encoder_mlp_quantities=3, encoder_repeat_times=2, mlp_head_last_units=8,
              mlp_head_dense_layers_quantity=2, mlp_head_dense_mish_block_repeats=1, use_mish=True)
    kw.update(over)
    return DetectorConfig(**kw)


def images(cfg, batch, seed=1234):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1, 1, size=(batch, *cfg.input_shape)).astype(np.float32)   # util.py:443-447 range


def build_model(cfg, weights, compute_mode="bf16"):
    from vision_transformer_detector_b200 import VisionTransformerDetector
    m = VisionTransformerDetector(cfg, seed=None, compute_mode=compute_mode)
    m.set_weights(weights)
    return m
