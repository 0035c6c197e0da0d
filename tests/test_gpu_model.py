"""GPU parity tests, whole path: create -> set_weights -> predict -> decode, through the reference-facing
Python API (which calls the C ABI), against the float64 oracle on identical weights and images."""
import json
import os

import numpy as np
import pytest

from _util import TOL_BF16, TOL_FP32, build_model, images, oracle, rel_err, tiny_config
import vision_transformer_detector_b200 as vd

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "threshold_vectors.json")


def test_library_weight_table_matches_host_table():
    for cfg in (vd.DetectorConfig(), tiny_config(mlp_head_dense_mish_block_repeats=2)):
        m = vd.VisionTransformerDetector(cfg, seed=None)
        assert m._specs == vd.weight_specs(cfg)
        assert m.count_params() == sum(int(np.prod(s)) for _, s in vd.weight_specs(cfg))
        m.close()


def test_create_defaults_and_api_surface():
    m = vd.create_vision_transformer_detector()
    assert m.name == "vision_transformer_detector"
    assert m.input_shape == (None, 608, 608, 3) and m.output_shape == (None, 17, 6)
    assert len(m.weights) == 245 and m.weights[0].name == "linear_projection/kernel:0"
    assert m.count_params() == 131_476_891
    w = m.get_weights()
    assert len(w) == 245 and w[0].shape == (867, 28)
    m2 = vd.VisionTransformerDetector.from_config(m.get_config(), seed=None)
    m2.set_weights(w)                                            # det.py:2123-2125, :2155-2157 usage
    assert all(np.array_equal(a, b) for a, b in zip(m2.get_weights(), w))
    with pytest.raises(ValueError):
        m2.set_weights(w[:-1])
    with pytest.raises(ValueError):
        m.predict(np.zeros((1, 600, 608, 3), np.float32))
    x = images(m.config, 1)
    assert m.predict(x).shape == (1, 17, 6)
    m.close(); m2.close()


def test_unset_weights_fail_loudly():
    m = vd.VisionTransformerDetector(tiny_config(), seed=None)
    with pytest.raises(vd._capi.VitdetError) as ei:
        m.predict(images(m.config, 1))
    assert ei.value.code == vd._capi.E_UNSET
    m.close()


@pytest.mark.parametrize("use_mish", [True, False])
@pytest.mark.parametrize("mode,tol", [("fp32", TOL_FP32), ("bf16", TOL_BF16)])
def test_tiny_model_logits_match_oracle(mode, tol, use_mish):
    cfg = tiny_config(use_mish=use_mish)
    w = vd.random_weights(cfg, seed=11, spread=True)
    x = images(cfg, 5)
    ref = oracle.forward(w, cfg, x, np.float64)
    m = build_model(cfg, w, mode)
    got = m.predict(x)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < tol
    m.close()


@pytest.mark.parametrize("cfg_over", [
    dict(input_shape=(64, 64, 3), patch_size=16, embedding_dim=64, encoder_num_heads=4, encoder_key_dim=64,
         encoder_mlp_quantities=3, encoder_repeat_times=2),                       # ViT-B-like proportions, no padding
    dict(input_shape=(136, 68, 3), encoder_num_heads=3, encoder_key_dim=24, mlp_head_dense_mish_block_repeats=2),
    dict(encoder_mlp_quantities=1, encoder_repeat_times=1, mlp_head_dense_layers_quantity=1),
    dict(encoder_num_heads=3, encoder_key_dim=96, encoder_repeat_times=2),       # heads wider than 64: two boxes per head tile
    dict(input_shape=(136, 68, 3), encoder_num_heads=2, encoder_key_dim=128, encoder_repeat_times=1),
    dict(encoder_num_heads=2, encoder_key_dim=72, encoder_repeat_times=1),
])
def test_configuration_knobs(cfg_over):
    cfg = tiny_config(**cfg_over)
    w = vd.random_weights(cfg, seed=4, spread=True)
    x = images(cfg, 3)
    ref = oracle.forward(w, cfg, x, np.float64)
    for mode, tol in (("fp32", TOL_FP32), ("bf16", TOL_BF16)):
        m = build_model(cfg, w, mode)
        assert rel_err(m.predict(x), ref) < tol
        m.close()


def _exact_decode_check(logits_gpu, rec, ref_logits, image_size):
    """north_star: in the fp32 mode, decoded detection indices and classes must be bit-exact wherever
    the reference's scores are not within tolerance of a threshold (bands 0.25*|dl| on
    sigma_0 and 39.5*|dl| on the class confidence, plus the rounding boundary of the class id)."""
    ref = oracle.decode(ref_logits, image_size=image_size)
    dl = float(np.abs(np.asarray(logits_gpu, np.float64) - ref_logits).max())
    obj, cc = ref["decoded"][..., 0], ref["class_conf"]
    near = (np.abs(obj - 0.5) <= 0.25 * dl + 1e-6) | (np.abs(cc - 0.5) <= 39.5 * dl + 1e-5) | (cc <= 39.5 * dl + 1e-5)
    keep = np.asarray(rec.keep).astype(bool)
    assert np.array_equal(keep[~near], ref["keep"][~near])
    assert np.array_equal(np.asarray(rec.class_id)[~near], ref["class_id"][~near])
    return int((~near).sum()), int(ref["keep"].sum())


def test_default_model_fp32_logits_and_exact_detections():
    """configs[0]-style case on the GPU: default config, Keras-default init AND the spread set."""
    cfg = vd.DetectorConfig()
    x = images(cfg, 2)
    for seed, spread in ((0, False), (1, True)):
        w = vd.random_weights(cfg, seed=seed, spread=spread)
        ref = oracle.forward(w, cfg, x, np.float64)
        m = build_model(cfg, w, "fp32")
        rec = m.detect(x, image_size=(608, 608))
        assert rel_err(rec.logits, ref) < TOL_FP32
        checked, kept = _exact_decode_check(rec.logits, rec, ref, (608, 608))
        assert checked >= 17                                    # the exactness check must not be vacuous
        # transform_predictions itself, on the GPU's own logits: float32 arithmetic vs float64
        dec_ref = oracle.transform_predictions(np.asarray(rec.logits, np.float64))
        assert np.abs(np.asarray(rec.decoded) - dec_ref).max() < 608 * 1e-6
        m.close()


def test_default_model_bf16_logits():
    cfg = vd.DetectorConfig()
    x = images(cfg, 3)
    w = vd.random_weights(cfg, seed=1, spread=True)
    ref = oracle.forward(w, cfg, x, np.float64)
    m = build_model(cfg, w, "bf16")
    got = m.predict(x)
    assert rel_err(got, ref) < TOL_BF16
    # images are independent: chunked execution gives the same rows
    m.set_chunk(2)
    assert np.array_equal(m.predict(x), got)
    m.close()


@pytest.mark.parametrize("name,cfg_kw", [
    ("hires", dict(input_shape=(1024, 1024, 3), patch_size=16)),                                   # BASELINE configs[3]: 4096 tokens
    ("vitb", dict(input_shape=(640, 640, 3), patch_size=16, embedding_dim=768, encoder_num_heads=12,
                  encoder_key_dim=64, encoder_repeat_times=12, encoder_mlp_quantities=3)),          # BASELINE configs[4]
])
def test_variant_configs_bf16_match_the_float32_oracle(name, cfg_kw):
    """The two variant configurations of BASELINE.json at their real shapes (one image).  The checker is the
    float32 torch restatement (the float64 numpy one needs minutes at 4096 tokens); its own distance to
    float64 is ~1e-6, far below the bf16 tolerance."""
    cfg = vd.DetectorConfig(**cfg_kw)
    w = vd.random_weights(cfg, seed=5, spread=True)
    x = images(cfg, 1)
    ref = oracle.forward_torch_f32(w, cfg, x)
    m = build_model(cfg, w, "bf16")
    got = m.predict(x)
    assert got.shape == (1, 17, 6)
    assert rel_err(got, ref) < TOL_BF16
    rec = m.detect(x)                       # decode scales by the model's own input size
    dref = oracle.decode(np.asarray(rec.logits, np.float64), image_size=cfg.input_shape[:2])
    assert np.abs(rec.decoded - dref["decoded"]).max() < max(cfg.input_shape[:2]) * 2e-6
    m.close()


def test_device_tensor_path_equals_host_path():
    import torch
    cfg = tiny_config()
    w = vd.random_weights(cfg, seed=2, spread=True)
    x = images(cfg, 4)
    m = build_model(cfg, w, "bf16")
    host = m.predict(x)
    dev = m(torch.from_numpy(x).cuda(), training=False)
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), host)
    rec_h = m.detect(x)
    rec_d = m.detect(torch.from_numpy(x).cuda())
    for a, b in zip((rec_h.decoded, rec_h.class_id, rec_h.class_conf, rec_h.keep, rec_h.corners),
                    (rec_d.decoded, rec_d.class_id, rec_d.class_conf, rec_d.keep, rec_d.corners)):
        assert np.array_equal(a, b.cpu().numpy())
    m.close()


def test_decode_reproduces_reference_golden_vectors():
    g = json.load(open(GOLDEN))
    slots = np.array([r["slot"] for r in g["reference"]], np.float32)
    rec = vd.decode_predictions(slots, strict=True, use_transform_predictions=False)
    assert rec.keep.astype(bool).tolist() == [r["keep_strict"] for r in g["reference"]]
    assert rec.class_id.tolist() == [r["class_id"] for r in g["reference"]]
    assert np.array_equal(rec.decoded, slots)
    d = np.array([r["slot"] for r in g["derived"]], np.float32)
    rs = vd.decode_predictions(d, strict=True, use_transform_predictions=False)
    rv = vd.decode_predictions(d, strict=False, use_transform_predictions=False)
    assert rs.keep.astype(bool).tolist() == [r["keep_strict"] for r in g["derived"]]
    assert rv.keep.astype(bool).tolist() == [r["keep_visualise"] for r in g["derived"]]
    assert rs.class_id.tolist() == [r["class_id"] for r in g["derived"]]


def test_transform_predictions_and_corners_match_oracle():
    rng = np.random.default_rng(9)
    logits = (rng.normal(size=(64, 17, 6)) * 3).astype(np.float32)
    logits[0, 0] = [np.nan, 0, 0, 0, 0, 0]
    logits[0, 1] = [100, -100, 100, -100, 100, -100]
    for size in ((608, 608), (1024, 1024), (480, 640)):
        ref = oracle.decode(logits.astype(np.float64), image_size=size)
        rec = vd.decode_predictions(logits, image_size=size)
        dec = vd.transform_predictions(logits, image_size=size)
        ok = ~np.isnan(ref["decoded"])
        assert np.abs(rec.decoded[ok] - ref["decoded"][ok]).max() < max(size) * 2e-6
        assert np.array_equal(dec, rec.decoded, equal_nan=True)
        assert np.isnan(rec.decoded[0, 0, 0]) and not rec.keep[0, 0]
        # corners / ids / keep computed by the oracle from the GPU's own float32 decode must agree exactly
        cid, cc, keep = oracle.threshold(rec.decoded.astype(np.float32))
        fin = ~np.isnan(rec.decoded[..., 0])
        assert np.array_equal(rec.class_id[fin], cid[fin]) and np.array_equal(rec.keep.astype(bool)[fin], keep[fin])
        assert np.array_equal(rec.corners[fin], oracle.corners(rec.decoded.astype(np.float32), size)[fin])
    e = vd.decode_predictions(np.zeros((0, 17, 6), np.float32))
    assert e.decoded.shape == (0, 17, 6) and e.keep.shape == (0, 17)


def test_iou_calculator_is_bit_exact_and_reproduces_known_answers():
    """N2 (first part): the reference's iou_calculator on the GPU, against its own test values and bit-exactly
    against the float32 oracle on random boxes (overlapping, disjoint, touching, degenerate)."""
    import torch
    g = json.load(open(GOLDEN))
    for row in g["iou"]:
        a = np.array([[[0.0, 79.0, *row["a"]]]], np.float32)
        b = np.array([[[0.0, 79.0, *row["b"]]]], np.float32)
        assert abs(float(vd.iou_calculator(a, b)[0, 0]) - row["iou"]) < 1e-6
    rng = np.random.default_rng(3)
    a = rng.uniform(0, 608, size=(64, 17, 6)).astype(np.float32)
    b = a + rng.normal(0, 40, size=a.shape).astype(np.float32)
    b[..., 4:] = np.abs(b[..., 4:])
    b[0] = a[0]                                  # identical boxes
    b[1, :, 2] = a[1, :, 2] + a[1, :, 5]         # touching / shifted by a full width
    a[2, :, 4:] = 0                              # zero-area labels
    ref = oracle.iou_calculator(a, b)
    got = vd.iou_calculator(a, b)
    assert got.shape == (64, 17) and got.dtype == np.float32
    assert np.array_equal(got, ref)
    dev = vd.iou_calculator(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    assert np.array_equal(dev.cpu().numpy(), ref)
    assert vd.iou_calculator(np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32)).shape == (0,)


def _logit(p):
    return float(np.log(p / (1 - p)))


def test_visualize_predictions_draws_the_kept_slots():
    """N3: thresholds, class ids and corner boxes come from the device decode (visualise rule); the host only draws."""
    B, H, W = 2, 608, 608
    imgs = np.zeros((B, H, W, 3), np.float32) - 1.0                  # black images
    logits = np.full((B, 17, 6), -20.0, np.float32)                  # every slot rejected ...
    # ... except: image 0 slot 3 = class 16 ("dog") at (cx 300.4, cy 200.4, h 100, w 60); image 1 slot 0 exactly AT the objectness threshold
    # (centres end in .4 so that float rounding cannot move an int() truncation across an integer)
    logits[0, 3] = [_logit(0.9), _logit(16.1 / 79), _logit(300.4 / 608), _logit(200.4 / 608), _logit(100 / 608), _logit(60 / 608)]
    logits[1, 0] = [0.0, _logit(2.0 / 79), _logit(100.4 / 608), _logit(100.4 / 608), _logit(40 / 608), _logit(40 / 608)]
    out = vd.visualize_predictions(imgs, predictions=logits)
    assert len(out) == 2 and out[0].shape == (H, W, 3) and out[0].dtype == np.uint8
    green = lambda im: (im[..., 1] == 255) & (im[..., 0] == 0) & (im[..., 2] == 0)
    g0 = green(out[0])
    assert g0[150, 270:330].all() and g0[250, 270:330].all() and g0[150:250, 270].all() and g0[150:250, 330].all()   # the rectangle
    assert not g0[200, 300]                                                                                         # hollow
    assert green(out[1])[80, 80:120].all()                 # objectness == 0.5 is kept by the visualise rule (skip only if '<')
    rec = vd.decode_predictions(logits, strict=False)
    assert rec.keep.sum() == 2 and rec.class_id[0, 3] == 16 and tuple(rec.corners[0, 3]) == (270, 150, 330, 250)
    assert vd.decode_predictions(logits, strict=True).keep.sum() == 1      # the metric rule drops the slot at the threshold
    # enlarged_image_scale: boxes and clip bounds scale, int() truncation after scaling
    big = vd.visualize_predictions(imgs[:1], predictions=logits[:1], enlarged_image_scale=1.5)
    assert big[0].shape == (912, 912, 3) and green(big[0])[225, 405:495].all()
    for scale in (1.0, 1.5, 0.7):
        r = vd.decode_predictions(logits, strict=False, corner_scale=scale)
        assert np.array_equal(r.corners, oracle.corners(r.decoded.astype(np.float32), (608, 608), scale))
    # label batches (already decoded rows, no confidence text): det.py:2420-2436
    labels = np.full((1, 17, 6), -8.0, np.float32); labels[..., 0] = 0
    labels[0, 1] = [1, 79, 10.2, 10.2, 10, 10]
    lab = vd.visualize_predictions([(imgs[:1], labels)])
    assert len(lab) == 1 and green(lab[0])[5, 5:15].all()


@pytest.mark.parametrize("shape", [(480, 640), (608, 608), (1024, 768), (123, 457), (700, 300), (1, 1), (2000, 3000)])
def test_preprocess_image_is_bit_exact(shape):
    """N4: resize_with_pad + clip + /127.5 - 1 from uint8 on the GPU, bit-exact against the float32 oracle."""
    import torch
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    img = rng.integers(0, 256, size=(*shape, 3), dtype=np.uint8)
    ref = oracle.preprocess_image(img)
    got = vd.preprocess_image(img)
    assert got.shape == (608, 608, 3) and got.dtype == np.float32
    assert np.array_equal(got, ref)
    assert vd.resize_with_pad_geometry(*shape, 608, 608) == oracle.resize_with_pad_geometry(*shape, 608, 608)
    dev = vd.preprocess_image(torch.from_numpy(img).cuda(), target_size=(320, 416))
    assert np.array_equal(dev.cpu().numpy(), oracle.preprocess_image(img, (320, 416)))


def test_image_file_to_detections(tmp_path):
    """File -> _get_image_tensor_coco -> predict -> visualize, the notebook's prediction path end to end."""
    from PIL import Image
    from vision_transformer_detector_b200 import vision_transformer_utilities as vu
    rng = np.random.default_rng(1)
    arr = rng.integers(0, 256, size=(360, 500, 3), dtype=np.uint8)
    path = str(tmp_path / "img.png")
    Image.fromarray(arr).save(path)
    tensor, size = vu._get_image_tensor_coco(path)
    assert size == (360, 500) and tensor.shape == (608, 608, 3)
    assert np.array_equal(tensor, oracle.preprocess_image(arr))
    m = vd.create_vision_transformer_detector(seed=1)
    logits = m.predict(tensor[None])
    out = vd.visualize_predictions(tensor[None], predictions=logits)
    assert logits.shape == (1, 17, 6) and out[0].shape == (608, 608, 3)
    m.close()


def test_full_size_batch_properties():
    """BASELINE configs[1] size (default config, batch 64, bf16), checked through size-independent properties: images
    are independent, so the result must be bit-identical under a permutation of the batch, under any split into
    sub-batches / encoder chunks, across repeated runs, and between the host-buffer path (staged H2D, lead chunks of
    4 / 12 images) and the device-tensor path (one chunk of 64).  A 3-image slice is also compared with the oracle."""
    import torch
    cfg = vd.DetectorConfig()
    w = vd.random_weights(cfg, seed=1, spread=True)
    m = build_model(cfg, w, "bf16")
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    x = torch.rand((64, *cfg.input_shape), generator=g, device="cuda") * 2 - 1
    rec = m.detect(x)
    base = rec.logits.cpu().numpy()
    assert np.isfinite(base).all()
    assert np.array_equal(m(x).cpu().numpy(), base)                                   # deterministic
    perm = torch.randperm(64, generator=torch.Generator().manual_seed(3))
    assert np.array_equal(m(x[perm.cuda()]).cpu().numpy(), base[perm.numpy()])        # permutation-equivariant
    assert np.array_equal(m(x[10:37]).cpu().numpy(), base[10:37])                     # any sub-batch
    host = m.detect(x.cpu().numpy())                                                  # predict_host: staged copy + lead chunks
    assert np.array_equal(host.logits, base)
    assert np.array_equal(host.keep, rec.keep.cpu().numpy()) and np.array_equal(host.corners, rec.corners.cpu().numpy())
    m.set_chunk(24)                                                                   # 24 + 24 + 16
    assert np.array_equal(m(x).cpu().numpy(), base)
    ref = oracle.forward_torch_f32(w, cfg, x[:3].cpu().numpy())
    assert rel_err(base[:3], ref) < TOL_BF16
    # decode of the full batch: ids / keep / corners recomputed by the oracle from the GPU's own decoded floats
    dec = rec.decoded.cpu().numpy()
    cid, cc, keep = oracle.threshold(dec)
    assert np.array_equal(rec.class_id.cpu().numpy(), cid) and np.array_equal(rec.keep.cpu().numpy().astype(bool), keep)
    assert np.array_equal(rec.corners.cpu().numpy(), oracle.corners(dec, cfg.input_shape[:2]))
    m.close()


def test_uint8_pixels_are_normalised_on_the_device_bit_exactly():
    """uint8 input ("next" row N4 fused into the patch kernel): with normalize_uint8=True x / 127.5 - 1 (utilities.py:446-447)
    happens inside patchify, so the result must equal predict(normalised float32) bit for bit, on the host and the device
    path; WITHOUT the flag a uint8 array is only cast to float32, as keras Model.predict does."""
    import torch
    cfg = tiny_config()
    rng = np.random.default_rng(7)
    u8 = rng.integers(0, 256, size=(5, *cfg.input_shape), dtype=np.uint8)
    u8[0, :3] = 0
    u8[1, -3:] = 255
    f32 = u8.astype(np.float32) / np.float32(127.5) - np.float32(1)
    for mode in ("fp32", "bf16"):
        model = build_model(cfg, vd.random_weights(cfg, seed=3, spread=True), mode)
        ref = model.predict(f32)
        assert np.array_equal(model.predict(u8, normalize_uint8=True), ref)
        assert np.array_equal(model(torch.from_numpy(u8).cuda(), normalize_uint8=True).cpu().numpy(), ref)
        a, b = model.detect(u8, normalize_uint8=True), model.detect(f32)
        assert np.array_equal(a.keep, b.keep) and np.array_equal(a.class_id, b.class_id) and np.array_equal(a.decoded, b.decoded)
        d = model.detect(torch.from_numpy(u8).cuda(), normalize_uint8=True)
        assert np.array_equal(d.decoded.cpu().numpy(), b.decoded) and np.array_equal(d.corners.cpu().numpy(), b.corners)
        # the Keras semantics: a plain dtype cast
        cast = model.predict(u8.astype(np.float32))
        assert np.array_equal(model.predict(u8), cast)
        assert np.array_equal(model(torch.from_numpy(u8).cuda()).cpu().numpy(), cast)
