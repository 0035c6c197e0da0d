"""Diagnostic: where do the bf16-store epilogue and the bf16-faithful oracle differ by more than one ulp?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import vitdet_oracle as o
from vision_transformer_detector_b200 import ops
R = o.bf16_round
def ulp(x):
    ax = np.maximum(np.abs(np.asarray(x, np.float64)), 2.0 ** -120)
    return 2.0 ** (np.floor(np.log2(ax)) - 7)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
for (M, K, N), pair in (((1296, 28, 3584), False), ((130, 200, 300), False), ((640, 3584, 1792), False), ((4096, 3584, 1792), True)):
    rng = np.random.default_rng(1)
    a = rng.normal(size=(M, K)).astype(np.float32)
    w = (rng.normal(size=(K, N)) * (1.5 / np.sqrt(K))).astype(np.float32)
    b = rng.normal(size=(N,)).astype(np.float32)
    pre = R(a) @ R(w) + b.astype(np.float64)
    ref = R(pre)
    got = ops.dense_ex(t(a), t(w), t(b), act=None, store_bf16=True, pair=pair).cpu().numpy().astype(np.float64)
    f32 = ops.dense_ex(t(a), t(w), t(b), act=None, store_bf16=False, pair=pair).cpu().numpy().astype(np.float64)
    d = (got - ref) / ulp(ref)
    print(f"shape {(M, K, N)} pair {pair}: exact {np.mean(got == ref):.4f}  |d|>1: {np.mean(np.abs(d) > 1.0001):.5f}  max|d| {np.abs(d).max():.2f}  f32-epilogue rel err {np.abs(f32 - pre).max() / np.abs(pre).max():.2e}")
    bad = np.argwhere(np.abs(d) > 1.0001)
    print("   bad columns mod 32 histogram:", np.bincount(bad[:, 1] % 32, minlength=32).tolist() if len(bad) else None)
    print("   bad rows mod 32 histogram:", np.bincount(bad[:, 0] % 32, minlength=32).tolist() if len(bad) else None)
    for r, c in bad[:6]:
        print(f"   [{r},{c}] pre {pre[r, c]:+.6f} ref {ref[r, c]:+.6f} got {got[r, c]:+.6f} f32epi {f32[r, c]:+.6f} R(f32epi) {R(f32[r, c]):+.6f}")
    # is `got` the rounding of the f32 epilogue's value?
    print("   got == R(f32 epilogue):", np.mean(got == R(f32)))
