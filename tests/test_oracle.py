"""CPU tests of the oracle itself: pinned against the reference's golden vectors, and checked for
internal consistency (float64 vs float32 vs the torch restatement, structural properties)."""
import json
import os

import numpy as np
import pytest

from _util import oracle, rel_err, tiny_config
from vision_transformer_detector_b200 import random_weights

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "threshold_vectors.json")


def test_threshold_rule_matches_reference_golden_vectors():
    g = json.load(open(GOLDEN))
    slots = np.array([r["slot"] for r in g["reference"]], np.float32)
    cid, cc, keep = oracle.threshold(slots, strict=True)
    assert keep.tolist() == [r["keep_strict"] for r in g["reference"]]
    assert cid.tolist() == [r["class_id"] for r in g["reference"]]
    # the confidences the reference's test docstrings quote: 79.255 -> 0.49, 79.3 -> 0.4
    assert abs(cc[4] - 0.49) < 1e-4 and abs(cc[7] - 0.4) < 1e-4


def test_threshold_rule_boundaries():
    g = json.load(open(GOLDEN))
    slots = np.array([r["slot"] for r in g["derived"]], np.float32)
    cid, _, keep_s = oracle.threshold(slots, strict=True)
    _, _, keep_v = oracle.threshold(slots, strict=False)
    assert keep_s.tolist() == [r["keep_strict"] for r in g["derived"]]
    assert keep_v.tolist() == [r["keep_visualise"] for r in g["derived"]]
    assert cid.tolist() == [r["class_id"] for r in g["derived"]]


def test_shape_facts_of_the_default_model():
    """Facts pinned by the reference's notebook artefacts: 1296 tokens of 867,
    245 weight tensors."""
    cfg = oracle.default_config()
    table = oracle.weight_table(cfg)
    assert len(table) == 245
    d = dict(table)
    assert d["linear_projection/kernel"] == (867, 28)
    assert d["position_encoding/position_embedding/embeddings"] == (1296, 1)
    assert d["MLP_1_1/kernel"] == (28, 3584) and d["MLP_1_2/kernel"] == (3584, 1792) and d["MLP_8_8/kernel"] == (56, 28)
    assert d["dense/kernel"] == (28, 17) and d["dense_1/kernel"] == (1296, 8704) and d["dense_7/kernel"] == (272, 136)
    assert d["MLP_Head_no_Sigmoid/kernel"] == (136, 6)
    assert d["multi_head_attention_7/query/kernel"] == (28, 8, 40) and d["multi_head_attention/attention_output/kernel"] == (8, 40, 28)
    assert "layer_normalization_15/gamma" in d and "layer_normalization_16/gamma" not in d


def test_extract_patches_same_padding():
    """608 = 36*17 - 4: two zero rows/cols on every side; depth order (row, col, channel)."""
    rng = np.random.default_rng(0)
    img = rng.uniform(-1, 1, (1, 608, 608, 3))
    p = oracle.extract_patches(img, 17)
    assert p.shape == (1, 1296, 867)
    first = p[0, 0].reshape(17, 17, 3)
    assert np.all(first[:2] == 0) and np.all(first[:, :2] == 0)
    assert np.array_equal(first[2:, 2:], img[0, :15, :15])
    last = p[0, -1].reshape(17, 17, 3)
    assert np.all(last[-2:] == 0) and np.all(last[:, -2:] == 0)
    assert np.array_equal(last[:15, :15], img[0, -15:, -15:])
    # token (1, 2) of the grid, element (r=3, c=5, ch=1)
    tok = p[0, 1 * 36 + 2].reshape(17, 17, 3)
    assert tok[3, 5, 1] == img[0, 17 * 1 + 3 - 2, 17 * 2 + 5 - 2, 1]
    # no padding when the size divides
    q = oracle.extract_patches(rng.uniform(-1, 1, (2, 32, 48, 3)), 16)
    assert q.shape == (2, 6, 768)


def test_head_reshape_is_a_flat_reinterpretation():
    """det.py:461: Reshape((17, -1)) of the (T, 17) Dense output — slot i reads flat [i*T, (i+1)*T)."""
    T = 12
    y = np.arange(T * 17, dtype=np.float64).reshape(1, T, 17)
    z = y.reshape(1, 17, -1)
    # slot 0 is made of tokens 0.. interleaved with the 17 outputs, not of column 0
    assert z[0, 1, 0] == T and z[0, 0, 1] == y[0, 0, 1] and z[0, 0, 11] == y[0, 0, 11] and z[0, 2, 0] == y[0, 1, 7]


def test_activations_and_layernorm():
    x = np.linspace(-20, 20, 101)
    assert np.allclose(oracle.mish(x), x * np.tanh(np.log1p(np.exp(x))), atol=1e-12)
    assert abs(oracle.gelu_tanh(np.array([1.0]))[0] - 0.841192) < 1e-5
    r = np.random.default_rng(1).normal(size=(5, 28))
    y = oracle.layer_norm(r, np.ones(28), np.zeros(28))
    assert np.allclose(y.mean(-1), 0, atol=1e-12)
    assert np.allclose(y.var(-1), r.var(-1) / (r.var(-1) + 1e-3), atol=1e-12)   # epsilon 1e-3, biased variance


def test_transform_predictions():
    l = np.array([[[0.0, 0.0, 0.0, 0.0, 0.0, 0.0], [50.0, -50.0, 2.0, -2.0, 1.0, -1.0]]])
    d = oracle.transform_predictions(l)
    assert np.allclose(d[0, 0], [0.5, 39.5, 304, 304, 304, 304])
    s = 1 / (1 + np.exp(-np.array([2.0, -2.0, 1.0, -1.0])))
    assert np.allclose(d[0, 1, 2:], s * 608) and d[0, 1, 0] > 0.999999 and d[0, 1, 1] < 1e-15
    d2 = oracle.transform_predictions(l, image_size=(100, 200))
    assert np.allclose(d2[0, 0, 2:], [100, 50, 50, 100])          # x,w scale by width; y,h by height


@pytest.mark.parametrize("use_mish", [True, False])
def test_f64_f32_torch_agree_on_a_tiny_model(use_mish):
    cfg = tiny_config(use_mish=use_mish)
    w = random_weights(cfg, seed=3, spread=True)
    rng = np.random.default_rng(5)
    img = rng.uniform(-1, 1, (3, *cfg.input_shape)).astype(np.float32)
    l64 = oracle.forward(w, cfg, img, np.float64)
    l32 = oracle.forward(w, cfg, img, np.float32)
    lt = oracle.forward_torch_f32(w, cfg, img)
    assert l64.shape == (3, 17, 6)
    assert rel_err(l32, l64) < 1e-4
    assert rel_err(lt, l64) < 1e-4


def test_attention_core_equals_mha_without_projections():
    rng = np.random.default_rng(2)
    B, T, H, d, D = 2, 9, 3, 8, 12
    x = rng.normal(size=(B, T, D))
    wq, wk, wv = (rng.normal(size=(D, H, d)) for _ in range(3))
    bq, bk, bv = (rng.normal(size=(H, d)) for _ in range(3))
    wo, bo = rng.normal(size=(H, d, D)), rng.normal(size=(D,))
    full = oracle.multi_head_attention(x, wq, bq, wk, bk, wv, bv, wo, bo)
    q = np.einsum("abc,cde->abde", x, wq) + bq
    k = np.einsum("abc,cde->abde", x, wk) + bk
    v = np.einsum("abc,cde->abde", x, wv) + bv
    o = oracle.attention_core(q, k, v)
    assert np.allclose(np.einsum("abcd,cde->abe", o, wo) + bo, full, atol=1e-12)


def test_iou_calculator_matches_reference_known_answers():
    """tests.py:170-195 (IoU 0.64 -> AP 0.3) and tests.py:223-248 (IoU 0.49 -> AP 0)."""
    g = json.load(open(GOLDEN))
    for row in g["iou"]:
        a = np.array([[1.0, 79.0, *row["a"]]], np.float32)
        b = np.array([[1.0, 79.0, *row["b"]]], np.float32)
        assert abs(float(oracle.iou_calculator(a, b)[0]) - row["iou"]) < 1e-6
    same = np.array([[3.0, 4.0, 5.0, 6.0]], np.float64)
    assert abs(oracle.iou_calculator(same, same)[0] - 1.0) < 1e-9
    apart = np.array([[100.0, 100.0, 5.0, 6.0]], np.float64)
    assert oracle.iou_calculator(same, apart)[0] == 0.0
    touching = np.array([[9.0, 4.0, 5.0, 6.0]], np.float64)      # shares an edge: strict comparison -> no overlap
    assert oracle.iou_calculator(same, touching)[0] == 0.0


def test_resize_with_pad_geometry_and_preprocess():
    """Structural checks of the input-side restatement (TF semantics; parity unpinned): float32 size arithmetic,
    centred zero padding mapped to -1, identity when no resize is needed, value range."""
    assert oracle.resize_with_pad_geometry(608, 608, 608, 608) == (608, 608, 0, 0)
    assert oracle.resize_with_pad_geometry(304, 608, 608, 608) == (304, 608, 152, 0)
    rh, rw, ph, pw = oracle.resize_with_pad_geometry(480, 640, 608, 608)
    assert rw in (607, 608) and rh in (455, 456) and pw == 0 and ph in (75, 76)          # the float32 quirk of TF may give 607
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(608, 608, 3), dtype=np.uint8)
    out = oracle.preprocess_image(img)
    assert out.dtype == np.float32 and np.array_equal(out, img.astype(np.float32) / np.float32(127.5) - np.float32(1))
    wide = rng.integers(0, 256, size=(100, 400, 3), dtype=np.uint8)
    o2 = oracle.preprocess_image(wide)
    assert o2.shape == (608, 608, 3) and (o2[:100] == -1).all() and (o2[-100:] == -1).all() and o2.min() >= -1 and o2.max() <= 1
    flat = np.full((50, 70, 3), 200, np.uint8)
    o3 = oracle.preprocess_image(flat, (64, 64))
    inside = o3[(o3 != -1).any(axis=-1)]
    assert np.allclose(inside, 200 / 127.5 - 1, atol=1e-6)


# ------------------------------------------------------------------------------------------------
# round 2: the second restatement and the bf16-faithful mode
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg_kw", [dict(), dict(use_mish=False, input_shape=(136, 68, 3), encoder_num_heads=3, encoder_key_dim=24,
                                                 mlp_head_dense_mish_block_repeats=2),
                                    dict(input_shape=(64, 64, 3), patch_size=16, embedding_dim=64, encoder_num_heads=4, encoder_key_dim=64),
                                    dict(input_shape=(70, 100, 3))])
def test_second_restatement_agrees_with_the_first(cfg_kw):
    """oracle/vitdet_oracle_torchnn.py is written on other primitives (F.unfold, nn.MultiheadAttention with its packed
    in-projection, nn.LayerNorm, F.mish / F.gelu, nn.Linear): a misreading of SAME padding, of the patch vector order, of the
    MHA kernel layouts, of the q scaling or of the head's flat Reshape in ONE of the two would show here."""
    import vitdet_oracle_torchnn as second
    from _util import tiny_config, images
    import vision_transformer_detector_b200 as vd
    cfg = tiny_config(**cfg_kw)
    w = vd.random_weights(cfg, seed=11, spread=True)
    x = images(cfg, 3)
    a = oracle.forward(w, cfg, x, np.float64)
    b = second.forward(w, cfg, x)
    assert np.abs(a - b).max() <= 1e-9 * np.abs(a).max()


def test_bf16_round_is_round_to_nearest_even():
    import torch
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(size=4000) * 3.3, [1.00390625, 1.01171875, -1.00390625, 0.0, 3.0e38, 1e-40, -2.5e-7]]).astype(np.float32)
    assert np.array_equal(oracle.bf16_round(x), torch.from_numpy(x).to(torch.bfloat16).to(torch.float64).numpy())
    assert oracle.bf16_round(np.array([[1.0, 2.0]])).shape == (1, 2)


def test_bf16_faithful_mode_is_consistent():
    """forward_bf16 = forward with the product's operand roundings: equal to the float64 forward when nothing needs
    rounding is impossible to arrange for a whole model, so check the two ends: it stays within the bf16 tolerance of the
    float64 result, and its first tap is exactly bf16(patches) @ bf16(W) + b + pos."""
    from _util import tiny_config, images, rel_err
    import vision_transformer_detector_b200 as vd
    cfg = tiny_config()
    w = vd.random_weights(cfg, seed=11, spread=True)
    x = images(cfg, 2)
    ref = oracle.forward(w, cfg, x, np.float64)
    got, inter = oracle.forward_bf16(w, cfg, x, return_intermediates=True)
    assert 1e-4 < rel_err(got, ref) < 2e-2
    R = oracle.bf16_round
    emb = R(oracle.extract_patches(x, cfg.patch_size)) @ R(w["linear_projection/kernel"]) + w["linear_projection/bias"].astype(np.float64) \
        + w["position_encoding/position_embedding/embeddings"].astype(np.float64)[None]
    assert np.array_equal(inter["embedded_patches"], emb)
