"""GPU parity tests, operator level: each CUDA kernel, called through the C ABI, against the oracle
on the same seeded inputs.  Tolerances: north_star's 1e-3 (fp32 mode) and 2e-2 (bf16 mode),
as max|gpu - ref| / max|ref| against the float64 oracle."""
import os
import numpy as np
import pytest

from _util import TOL_BF16, TOL_FP32, oracle, rel_err

pytestmark = pytest.mark.gpu


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


ACTS = {None: lambda x: x, "mish": oracle.mish, "gelu": oracle.gelu_tanh}

# (M, K, N): every distinct Dense shape class of the default model plus ragged ones
DENSE_SHAPES = [
    (1296, 867, 28),      # linear_projection: K tail (867 = 13*64 + 35), N < 16-multiple
    (1296, 28, 3584),     # MLP_i_1: K < one k-block
    (640, 3584, 1792),    # MLP_i_2: the big one, 7 N-tiles of 256
    (300, 1792, 896),     # block_n 224, M tail
    (257, 448, 224), (129, 112, 56), (1296, 56, 28),
    (1296, 28, 1536),     # fused QKV width
    (1296, 512, 28),      # attention_output
    (34, 1296, 8704),     # head, M = 2 images x 17 slots
    (34, 272, 136),
    (1, 64, 16), (127, 70, 17), (128, 16, 8), (130, 200, 300),
    # M >= 1024 and N >= 128: the CTA-pair (cta_group::2) kernel; M tails on one or both CTAs of the last pair
    (4096, 3584, 1792), (1024, 64, 128), (1100, 200, 300), (1300, 28, 3584), (2592, 896, 448),
]


@pytest.mark.parametrize("mode,tol", [("bf16", TOL_BF16), ("fp32", TOL_FP32)])
@pytest.mark.parametrize("shape", DENSE_SHAPES)
def test_dense_matches_oracle(shape, mode, tol):
    from vision_transformer_detector_b200 import ops
    M, K, N = shape
    rng = np.random.default_rng(hash(shape) % 2**31)
    a = rng.normal(size=(M, K)).astype(np.float32)
    w = (rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)
    b = rng.normal(size=(N,)).astype(np.float32)
    ref = a.astype(np.float64) @ w.astype(np.float64) + b
    got = ops.dense(_t(a), _t(w), _t(b), mode=mode).cpu().numpy()
    assert got.shape == (M, N)
    assert rel_err(got, ref) < tol


@pytest.mark.parametrize("mode,tol", [("bf16", TOL_BF16), ("fp32", TOL_FP32)])
@pytest.mark.parametrize("act", ["mish", "gelu", None])
def test_dense_epilogue_activation_and_residual(act, mode, tol):
    from vision_transformer_detector_b200 import ops
    M, K, N = 700, 224, 112
    rng = np.random.default_rng(7)
    a = rng.normal(size=(M, K)).astype(np.float32)
    w = (rng.normal(size=(K, N)) * (2.0 / np.sqrt(K))).astype(np.float32)
    b = rng.normal(size=(N,)).astype(np.float32)
    r = rng.normal(size=(M, N)).astype(np.float32)
    ref = ACTS[act](a.astype(np.float64) @ w.astype(np.float64) + b) + r          # Dense -> act -> add (det.py:388-412)
    got = ops.dense(_t(a), _t(w), _t(b), resid=_t(r), act=act, mode=mode).cpu().numpy()
    assert rel_err(got, ref) < tol


def test_dense_fp32_activation_is_accurate_over_a_wide_range():
    """Mish / GELU at large |x| (softplus overflow guard) in the exact mode."""
    from vision_transformer_detector_b200 import ops
    x = np.linspace(-60, 60, 4 * 128, dtype=np.float32).reshape(4, 128)
    eye = np.eye(128, dtype=np.float32)
    for act in ("mish", "gelu"):
        got = ops.dense(_t(x), _t(eye), None, act=act, mode="fp32").cpu().numpy()
        ref = ACTS[act](x.astype(np.float64))
        assert np.abs(got - ref).max() < 1e-4 * 60


@pytest.mark.parametrize("M,D", [(1296 * 2, 28), (1000, 768), (77, 30), (5, 4), (300, 130)])
def test_layernorm_matches_oracle(M, D):
    from vision_transformer_detector_b200 import ops
    rng = np.random.default_rng(M + D)
    x = (rng.normal(size=(M, D)) * 3 + 1).astype(np.float32)
    g = rng.normal(size=(D,)).astype(np.float32)
    b = rng.normal(size=(D,)).astype(np.float32)
    ref = oracle.layer_norm(x.astype(np.float64), g.astype(np.float64), b.astype(np.float64), 1e-3)
    got = ops.layernorm(_t(x), _t(g), _t(b), 1e-3).cpu().numpy()
    assert rel_err(got, ref) < 1e-5


@pytest.mark.parametrize("mode,tol", [("bf16", TOL_BF16), ("fp32", TOL_FP32)])
@pytest.mark.parametrize("B,T,H,d", [(2, 1296, 8, 40), (1, 4096, 2, 40), (2, 1600, 3, 64), (3, 100, 2, 40), (1, 64, 1, 8), (2, 65, 2, 24),
                                     (2, 300, 3, 96), (1, 1296, 2, 128), (2, 65, 2, 72), (1, 129, 5, 104)])      # heads wider than 64: two boxes per head
def test_attention_matches_oracle(B, T, H, d, mode, tol):
    from vision_transformer_detector_b200 import ops
    rng = np.random.default_rng(B * T + d)
    q, k, v = (rng.normal(size=(B, T, H, d)).astype(np.float32) for _ in range(3))
    ref = oracle.attention_core(q.astype(np.float64), k.astype(np.float64), v.astype(np.float64))
    got = ops.attention(_t(q), _t(k), _t(v), mode=mode).cpu().numpy()
    assert rel_err(got, ref) < tol


def test_attention_is_invariant_to_a_key_shift():
    """softmax(q.(k + c)) with c orthogonal... property: adding a constant vector u to every key adds the
    same q.u to every score of a row, which softmax ignores."""
    from vision_transformer_detector_b200 import ops
    rng = np.random.default_rng(3)
    B, T, H, d = 1, 1296, 8, 40
    q, k, v = (rng.normal(size=(B, T, H, d)).astype(np.float32) for _ in range(3))
    u = rng.normal(size=(1, 1, H, d)).astype(np.float32) * 0.5
    a = ops.attention(_t(q), _t(k), _t(v), mode="fp32").cpu().numpy()
    b = ops.attention(_t(q), _t(k + u), _t(v), mode="fp32").cpu().numpy()
    assert rel_err(b, a) < 1e-4


@pytest.mark.parametrize("B,H,W,p", [(2, 608, 608, 17), (1, 1024, 1024, 16), (3, 60, 130, 17), (1, 17, 17, 17), (2, 33, 50, 8)])
def test_patchify_is_bit_exact(B, H, W, p):
    from vision_transformer_detector_b200 import ops
    rng = np.random.default_rng(H + W)
    img = rng.uniform(-1, 1, size=(B, H, W, 3)).astype(np.float32)
    ref = oracle.extract_patches(img, p)
    got = ops.patchify(_t(img), p).cpu().numpy()
    assert np.array_equal(got, ref)


def test_shared_memory_opt_in_grows_with_later_launches():
    """A kernel whose dynamic shared-memory size depends on the model (the wide slot projection: 17 x D floats) must work
    when a SMALL model runs first and a larger one later in the same process: the opt-in above 48 KB is remembered per
    (kernel, device) with its size and raised on demand (launch.h).  Own process, so that no earlier test has set it."""
    import subprocess
    import sys
    code = r"""
import numpy as np, torch
from vision_transformer_detector_b200 import ops
rng = np.random.default_rng(0)
for D in (64, 768, 1024):
    x = rng.normal(size=(1, 40, D)).astype(np.float32)
    w = (rng.normal(size=(D, 17)) / np.sqrt(D)).astype(np.float32)
    b = rng.normal(size=(17,)).astype(np.float32)
    ref = (x.astype(np.float64) @ w.astype(np.float64) + b).reshape(1, 17, 40)
    got = ops.head_slots(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda(), mode="fp32").cpu().numpy()
    assert np.abs(got - ref).max() < 1e-5 * np.abs(ref).max(), D
print("ok")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
