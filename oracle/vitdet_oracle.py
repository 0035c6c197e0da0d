"""CPU ORACLE — test infrastructure, not product code.

A plain restatement of the reference's algorithm for the ViT-detector forward pass + head decode
(westlake-moonlight/vision_transformer_detector, vision_transformer_detector.py = "det.py").  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product (vision_transformer_detector_b200/) never does.

PARITY PINNING.  The arithmetic of the path lives in third-party packages that are NOT vendored in
the reference and NOT installable here (tensorflow 2.9.1, keras 2.9, tensorflow_addons 0.17-0.18; the
image has Python 3.12 and no network), and the reference ships no weights, activations or forward-pass
tests.  Therefore:
  * the threshold / class-id rule IS pinned: tests/golden/threshold_vectors.json transcribes the
    known-answer vectors of the reference's own tests (testcases_vision_transformer_detector.py:284-303,
    342-370, 405-426, 507, 566-585, 622-641, 688) and this oracle reproduces every one of them;
  * the forward pass (patches -> logits) and transform_predictions are "PARITY UNPINNED": they follow
    the reference source line by line and the published Keras/TF/TFA layer semantics (listed at each
    function), cross-checked here in float64 vs float32 and numpy vs torch, but could not be compared
    against a TensorFlow run.

Two arithmetic modes: numpy with dtype float64 (ground truth) or float32, and torch float32 on all host
cores (stands in for "TF on CPU" in the timed CPU baseline).
"""
from __future__ import annotations

import math

import numpy as np

CLASSES = 80                       # Constants.CLASSES                       det.py:20
MODEL_IMAGE_SIZE = (608, 608)      # Constants.MODEL_IMAGE_SIZE (h, w)       det.py:22
SLOTS = 17                         # Constants.MAX_DETECT_OBJECTS_QUANTITY   det.py:28
OBJECTNESS_THRESHOLD = 0.5         # det.py:41
CLASSIFICATION_CONFIDENCE_THRESHOLD = 0.5   # det.py:43
LN_EPSILON = 1e-3                  # keras.layers.LayerNormalization default epsilon


# ------------------------------------------------------------------------------------------------
# configuration helpers (duck-typed: any object with the reference's keyword names as attributes)
# ------------------------------------------------------------------------------------------------
def _cfg(cfg, name, default=None):
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    return getattr(cfg, name, default)


def default_config() -> dict:
    """Defaults of create_vision_transformer_detector, det.py:498-506."""
    return dict(input_shape=(608, 608, 3), patch_size=17, embedding_dim=28, encoder_num_heads=8, encoder_key_dim=40,
                encoder_mlp_quantities=8, encoder_repeat_times=8, mlp_head_last_units=136,
                mlp_head_dense_layers_quantity=7, mlp_head_dense_mish_block_repeats=1, use_mish=True)


def _kname(base: str, i: int) -> str:
    # keras auto-naming after keras.backend.clear_session() (det.py:548)
    return base if i == 0 else f"{base}_{i}"


def weight_table(cfg) -> list[tuple[str, tuple[int, ...]]]:
    """(Keras variable name, shape) in model.weights order, derived from det.py:239-495."""
    H_img, W_img = _cfg(cfg, "input_shape")[:2]
    p = _cfg(cfg, "patch_size")
    D = _cfg(cfg, "embedding_dim")
    H, d = _cfg(cfg, "encoder_num_heads"), _cfg(cfg, "encoder_key_dim")
    q, L = _cfg(cfg, "encoder_mlp_quantities"), _cfg(cfg, "encoder_repeat_times")
    T = math.ceil(H_img / p) * math.ceil(W_img / p)
    t = [("linear_projection/kernel", (3 * p * p, D)), ("linear_projection/bias", (D,)),           # det.py:297
         ("position_encoding/position_embedding/embeddings", (T, 1))]                               # det.py:148, 292
    for i in range(L):
        ln1, ln2, mha = _kname("layer_normalization", 2 * i), _kname("layer_normalization", 2 * i + 1), _kname("multi_head_attention", i)
        t += [(ln1 + "/gamma", (D,)), (ln1 + "/beta", (D,))]                                        # det.py:353
        for s in ("query", "key", "value"):                                                         # det.py:364
            t += [(f"{mha}/{s}/kernel", (D, H, d)), (f"{mha}/{s}/bias", (H, d))]
        t += [(mha + "/attention_output/kernel", (H, d, D)), (mha + "/attention_output/bias", (D,))]
        t += [(ln2 + "/gamma", (D,)), (ln2 + "/beta", (D,))]                                        # det.py:375
        fan = D
        for j in range(q):                                                                          # det.py:385-394
            u = D * 2 ** (q - 1 - j)
            t += [(f"MLP_{i + 1}_{j + 1}/kernel", (fan, u)), (f"MLP_{i + 1}_{j + 1}/bias", (u,))]
            fan = u
    k = 0
    t += [(_kname("dense", k) + "/kernel", (D, SLOTS)), (_kname("dense", k) + "/bias", (SLOTS,))]   # det.py:454
    k += 1
    fan = T
    n, rep, last = _cfg(cfg, "mlp_head_dense_layers_quantity"), _cfg(cfg, "mlp_head_dense_mish_block_repeats"), _cfg(cfg, "mlp_head_last_units")
    for e in reversed(range(n)):                                                                    # det.py:465-476
        for _ in range(rep):
            u = last * 2 ** e
            t += [(_kname("dense", k) + "/kernel", (fan, u)), (_kname("dense", k) + "/bias", (u,))]
            fan = u
            k += 1
    t += [("MLP_Head_no_Sigmoid/kernel", (fan, 6)), ("MLP_Head_no_Sigmoid/bias", (6,))]             # det.py:489
    return t


# ------------------------------------------------------------------------------------------------
# numpy restatement (dtype = float64 ground truth, or float32)
# ------------------------------------------------------------------------------------------------
def extract_patches(images: np.ndarray, p: int) -> np.ndarray:
    """tf.image.extract_patches(sizes=strides=[1,p,p,1], rates=1, padding='SAME') (det.py:195-197)
    + Reshape((-1, 3p^2)) (det.py:279-280).
    SAME: out = ceil(size / p); pad_total = out*p - size; pad_before = pad_total // 2; zeros.
    Depth order of a patch is (row, col, channel); tokens are row-major over the patch grid."""
    B, H, W, Cc = images.shape
    gh, gw = -(-H // p), -(-W // p)
    pt, pl = (gh * p - H) // 2, (gw * p - W) // 2
    padded = np.zeros((B, gh * p, gw * p, Cc), images.dtype)
    padded[:, pt:pt + H, pl:pl + W, :] = images
    x = padded.reshape(B, gh, p, gw, p, Cc).transpose(0, 1, 3, 2, 4, 5)
    return x.reshape(B, gh * gw, p * p * Cc)


def softplus(x: np.ndarray) -> np.ndarray:
    return np.logaddexp(x, 0)


def mish(x: np.ndarray) -> np.ndarray:
    """tfa.activations.mish: x * tanh(softplus(x))  (det.py:129)."""
    return x * np.tanh(softplus(x))


def gelu_tanh(x: np.ndarray) -> np.ndarray:
    """tfa.layers.GELU() default approximate=True (det.py:402, :483)."""
    c = x.dtype.type(math.sqrt(2.0 / math.pi))
    return x.dtype.type(0.5) * x * (1 + np.tanh(c * (x + x.dtype.type(0.044715) * x * x * x)))


def layer_norm(x: np.ndarray, gamma: np.ndarray, beta: np.ndarray, eps: float = LN_EPSILON) -> np.ndarray:
    """keras.layers.LayerNormalization(axis=-1): biased variance, epsilon inside the sqrt (det.py:353, :375)."""
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + x.dtype.type(eps)) * gamma + beta


def multi_head_attention(x: np.ndarray, wq, bq, wk, bk, wv, bv, wo, bo) -> np.ndarray:
    """keras.layers.MultiHeadAttention(num_heads=H, key_dim=d)(query=x, value=x) (det.py:364-369):
    q/k/v = einsum('abc,cde->abde') + bias; q *= 1/sqrt(d) AFTER the bias; scores = einsum('aecd,abcd->acbe')
    (k, q); softmax over keys; out = einsum('acbe,aecd->abcd'); projection einsum('abcd,cde->abe') + bias."""
    d = wq.shape[-1]
    q = np.einsum("abc,cde->abde", x, wq) + bq
    k = np.einsum("abc,cde->abde", x, wk) + bk
    v = np.einsum("abc,cde->abde", x, wv) + bv
    q = q * x.dtype.type(1.0 / math.sqrt(d))
    s = np.einsum("aecd,abcd->acbe", k, q)                       # (B, H, Tq, Tk)
    s = s - s.max(axis=-1, keepdims=True)
    e = np.exp(s)
    pr = e / e.sum(axis=-1, keepdims=True)
    o = np.einsum("acbe,aecd->abcd", pr, v)                      # (B, T, H, d)
    return np.einsum("abcd,cde->abe", o, wo) + bo


def attention_core(q: np.ndarray, k: np.ndarray, v: np.ndarray) -> np.ndarray:
    """The part of MHA between the projections: q,k,v (B,T,H,d), q unscaled."""
    d = q.shape[-1]
    s = np.einsum("aecd,abcd->acbe", k, q * q.dtype.type(1.0 / math.sqrt(d)))
    s = s - s.max(axis=-1, keepdims=True)
    e = np.exp(s)
    pr = e / e.sum(axis=-1, keepdims=True)
    return np.einsum("acbe,aecd->abcd", pr, v)


def forward(weights: dict, cfg, images: np.ndarray, dtype=np.float64, return_intermediates: bool = False):
    """model(images) -> raw logits (B, 17, 6): det.py:555-581 wiring of det.py:239-495."""
    w = {k: np.asarray(v, dtype=dtype) for k, v in weights.items()}
    act = mish if _cfg(cfg, "use_mish", True) else gelu_tanh
    p = _cfg(cfg, "patch_size")
    L, qn = _cfg(cfg, "encoder_repeat_times"), _cfg(cfg, "encoder_mlp_quantities")
    inter = {}
    x = extract_patches(np.asarray(images, dtype=dtype), p)                                # det.py:271-280
    x = x @ w["linear_projection/kernel"] + w["linear_projection/bias"]                    # det.py:297
    x = x + w["position_encoding/position_embedding/embeddings"][None, :, :]               # det.py:291-307 (1,T,1) broadcast
    inter["embedded_patches"] = x
    for i in range(L):
        ln1, ln2, mha = _kname("layer_normalization", 2 * i), _kname("layer_normalization", 2 * i + 1), _kname("multi_head_attention", i)
        side = x
        y = layer_norm(x, w[ln1 + "/gamma"], w[ln1 + "/beta"])
        y = multi_head_attention(y, w[mha + "/query/kernel"], w[mha + "/query/bias"], w[mha + "/key/kernel"], w[mha + "/key/bias"],
                                 w[mha + "/value/kernel"], w[mha + "/value/bias"], w[mha + "/attention_output/kernel"],
                                 w[mha + "/attention_output/bias"])
        x = y + side                                                                       # det.py:371
        side = x
        y = layer_norm(x, w[ln2 + "/gamma"], w[ln2 + "/beta"])
        for j in range(qn):                                                                # det.py:388-402
            y = act(y @ w[f"MLP_{i + 1}_{j + 1}/kernel"] + w[f"MLP_{i + 1}_{j + 1}/bias"])
        x = y + side                                                                       # det.py:408-412
        inter[f"block_{i + 1}"] = x
    inter["encoded_images"] = x
    B = x.shape[0]
    k = 0
    y = x @ w[_kname("dense", k) + "/kernel"] + w[_kname("dense", k) + "/bias"]            # det.py:454
    k += 1
    y = y.reshape(B, SLOTS, -1)                                                            # det.py:461 (flat reinterpretation)
    n, rep = _cfg(cfg, "mlp_head_dense_layers_quantity"), _cfg(cfg, "mlp_head_dense_mish_block_repeats")
    for _ in range(n * rep):                                                               # det.py:468-483
        y = act(y @ w[_kname("dense", k) + "/kernel"] + w[_kname("dense", k) + "/bias"])
        k += 1
    inter["head_last"] = y
    logits = y @ w["MLP_Head_no_Sigmoid/kernel"] + w["MLP_Head_no_Sigmoid/bias"]           # det.py:489
    if return_intermediates:
        return logits, inter
    return logits


def sigmoid(x: np.ndarray) -> np.ndarray:
    return 1 / (1 + np.exp(-x))


def transform_predictions(logits: np.ndarray, image_size=MODEL_IMAGE_SIZE, classes: int = CLASSES) -> np.ndarray:
    """det.py:586-647.  image_size = (height, width)."""
    s = sigmoid(np.asarray(logits))
    s = np.concatenate([s[..., :-4], np.clip(s[..., -4:], 0, 1)], axis=-1)                 # det.py:623-625
    ih, iw = image_size
    scale = np.array([1, classes - 1, iw, ih, ih, iw], dtype=s.dtype)                      # det.py:628-640
    return s * scale


def threshold(decoded: np.ndarray, objectness_threshold: float = OBJECTNESS_THRESHOLD,
              classification_threshold: float = CLASSIFICATION_CONFIDENCE_THRESHOLD, strict: bool = True):
    """Class id / class confidence / keep mask of already-decoded slots.
    strict=True : metric rule  det.py:1366-1384  keep iff objectness > thr and class_conf > thr
    strict=False: visualise    det.py:2264-2282  skip iff objectness < thr or class_conf < thr
    np.round / tf.round are round-half-to-even."""
    dec = np.asarray(decoded)
    obj, cls = dec[..., 0], dec[..., 1]
    cid = np.round(cls)
    err = np.abs(cls - cid)
    half = dec.dtype.type(0.5)
    cc = (half - err) / half
    if strict:
        keep = (obj > objectness_threshold) & (cc > classification_threshold)
    else:
        keep = ~(obj < objectness_threshold) & ~(cc < classification_threshold)
    return cid.astype(np.int32), cc, keep


def corners(decoded: np.ndarray, image_size=MODEL_IMAGE_SIZE, scale: float = 1.0) -> np.ndarray:
    """det.py:2294-2325: boxes times enlarged_image_scale, int() truncation, then clip to the (enlarged) image, whose
    size is round(size * scale) (det.py:2237-2252)."""
    dec = np.asarray(decoded)
    s = dec.dtype.type(scale)
    ih, iw = (int(round(float(dec.dtype.type(v) * s))) for v in image_size)
    cx, cy, bh, bw = dec[..., 2] * s, dec[..., 3] * s, dec[..., 4] * s, dec[..., 5] * s
    x0 = np.clip(np.trunc(cx - bw / 2).astype(np.int64), 0, int(iw))
    y0 = np.clip(np.trunc(cy - bh / 2).astype(np.int64), 0, int(ih))
    x1 = np.clip(np.trunc(cx + bw / 2).astype(np.int64), 0, int(iw))
    y1 = np.clip(np.trunc(cy + bh / 2).astype(np.int64), 0, int(ih))
    return np.stack([x0, y0, x1, y1], axis=-1).astype(np.int32)


def decode(logits: np.ndarray, image_size=MODEL_IMAGE_SIZE, objectness_threshold=OBJECTNESS_THRESHOLD,
           classification_threshold=CLASSIFICATION_CONFIDENCE_THRESHOLD, strict: bool = True) -> dict:
    dec = transform_predictions(logits, image_size)
    cid, cc, keep = threshold(dec, objectness_threshold, classification_threshold, strict)
    return {"decoded": dec, "class_id": cid, "class_conf": cc, "keep": keep, "corners": corners(dec, image_size)}


def resize_with_pad_geometry(h: int, w: int, th: int, tw: int):
    """Size arithmetic of tf.image.resize_with_pad (TF 2.9 image_ops_impl._resize_with_pad_common), in float32 as TF
    computes it: ratio = max(w/tw, h/th); resized = floor(size / ratio); pad = max(0, floor((target - size/ratio) / 2))."""
    f = np.float32
    ratio = max(f(w) / f(tw), f(h) / f(th))
    rhf, rwf = f(h) / ratio, f(w) / ratio
    rh, rw = int(np.floor(rhf)), int(np.floor(rwf))
    ph = max(0, int(np.floor((f(th) - rhf) / f(2))))
    pw = max(0, int(np.floor((f(tw) - rwf) / f(2))))
    return rh, rw, ph, pw


def preprocess_image(image_u8: np.ndarray, target=MODEL_IMAGE_SIZE) -> np.ndarray:
    """vision_transformer_utilities.py:435-447 after the file decode: tf.image.resize_with_pad (bilinear,
    half_pixel_centers=True, antialias=False; TF resize_bilinear CPU kernel: in = (i + 0.5) * scale - 0.5, lower =
    max(floor(in), 0), upper = min(ceil(in), size - 1), lerp = in - floor(in); top/bottom lerp in x then y, float32),
    zero padding, clip to [0, 255], / 127.5, - 1.  PARITY UNPINNED (TF semantics restated, not executed)."""
    f = np.float32
    img = np.asarray(image_u8)
    h, w = img.shape[:2]
    th, tw = target
    rh, rw, ph, pw = resize_with_pad_geometry(h, w, th, tw)
    hs, ws = f(h) / f(rh), f(w) / f(rw)
    iy = (np.arange(rh, dtype=f) + f(0.5)) * hs - f(0.5)
    ix = (np.arange(rw, dtype=f) + f(0.5)) * ws - f(0.5)
    fy, fx = np.floor(iy), np.floor(ix)
    y0 = np.maximum(fy.astype(np.int64), 0); y1 = np.minimum(np.ceil(iy).astype(np.int64), h - 1)
    x0 = np.maximum(fx.astype(np.int64), 0); x1 = np.minimum(np.ceil(ix).astype(np.int64), w - 1)
    ly, lx = (iy - fy)[:, None, None], (ix - fx)[None, :, None]
    src = img.astype(f)
    tl, tr = src[y0][:, x0], src[y0][:, x1]
    bl, br = src[y1][:, x0], src[y1][:, x1]
    top = tl + (tr - tl) * lx
    bot = bl + (br - bl) * lx
    res = top + (bot - top) * ly
    out = np.zeros((th, tw, 3), f)
    out[ph:ph + rh, pw:pw + rw] = res
    out = np.clip(out, 0, 255)
    return out / f(127.5) - f(1)


EPSILON = 1e-8                     # Constants.EPSILON det.py:24


def iou_calculator(label_bbox: np.ndarray, prediction_bbox: np.ndarray) -> np.ndarray:
    """det.py:761-875, statement by statement: edges, strict overlap test, zero the edges of non-overlapping pairs,
    sort the four edges, intersection = sorted[-2] - sorted[-3] per axis, IoU = I / (U + EPSILON).  Boxes are the last
    four entries (center_x, center_y, height, width) of the last axis.  Pinned by the reference's own known answers
    (tests.py:170-195: 0.64; tests.py:223-248: 0.49)."""
    lb, pb = np.asarray(label_bbox), np.asarray(prediction_bbox)
    half = lb.dtype.type(2)
    l_left, l_right = lb[..., -4] - lb[..., -1] / half, lb[..., -4] + lb[..., -1] / half            # det.py:789-790
    p_left, p_right = pb[..., -4] - pb[..., -1] / half, pb[..., -4] + pb[..., -1] / half            # det.py:791-794
    l_top, l_bottom = lb[..., -3] - lb[..., -2] / half, lb[..., -3] + lb[..., -2] / half            # det.py:796-797
    p_top, p_bottom = pb[..., -3] - pb[..., -2] / half, pb[..., -3] + pb[..., -2] / half            # det.py:798-801
    hit = (l_left < p_right) & (l_right > p_left) & (l_top < p_bottom) & (l_bottom > p_top)         # det.py:806-817
    hor = np.stack([l_top, l_bottom, p_top, p_bottom], axis=-1)                                     # det.py:826-828
    ver = np.stack([l_left, l_right, p_left, p_right], axis=-1)
    hor = np.sort(np.where(hit[..., None], hor, 0), axis=-1)                                        # det.py:837-845
    ver = np.sort(np.where(hit[..., None], ver, 0), axis=-1)
    inter = (hor[..., -2] - hor[..., -3]) * (ver[..., -2] - ver[..., -3])                           # det.py:849-853
    union = pb[..., -1] * pb[..., -2] + lb[..., -1] * lb[..., -2] - inter                           # det.py:856-866
    return inter / (union + lb.dtype.type(EPSILON))                                                 # det.py:873


# ------------------------------------------------------------------------------------------------
# torch float32 restatement on the host cores (the timed CPU baseline: "TF on CPU" stand-in)
# ------------------------------------------------------------------------------------------------
def forward_torch_f32(weights: dict, cfg, images, num_threads: int | None = None):
    """Same graph as forward(), float32, torch CPU ops (MKL/oneDNN on all host threads)."""
    import torch
    import torch.nn.functional as F
    if num_threads:
        torch.set_num_threads(int(num_threads))
    with torch.no_grad():
        w = {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))) for k, v in weights.items()}
        use_mish = _cfg(cfg, "use_mish", True)
        act = (lambda t: t * torch.tanh(F.softplus(t))) if use_mish else (lambda t: F.gelu(t, approximate="tanh"))
        p = _cfg(cfg, "patch_size")
        L, qn = _cfg(cfg, "encoder_repeat_times"), _cfg(cfg, "encoder_mlp_quantities")
        img = images if isinstance(images, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(images, dtype=np.float32))
        B, H, W, Cc = img.shape
        gh, gw = -(-H // p), -(-W // p)
        pt, pl = (gh * p - H) // 2, (gw * p - W) // 2
        padded = torch.zeros((B, gh * p, gw * p, Cc), dtype=torch.float32)
        padded[:, pt:pt + H, pl:pl + W, :] = img
        x = padded.reshape(B, gh, p, gw, p, Cc).permute(0, 1, 3, 2, 4, 5).reshape(B, gh * gw, p * p * Cc)
        x = x @ w["linear_projection/kernel"] + w["linear_projection/bias"]
        x = x + w["position_encoding/position_embedding/embeddings"][None]
        D = x.shape[-1]
        for i in range(L):
            ln1, ln2, mha = _kname("layer_normalization", 2 * i), _kname("layer_normalization", 2 * i + 1), _kname("multi_head_attention", i)
            side = x
            y = F.layer_norm(x, (D,), w[ln1 + "/gamma"], w[ln1 + "/beta"], LN_EPSILON)
            wq = w[mha + "/query/kernel"]
            d = wq.shape[-1]
            q = torch.einsum("abc,cde->abde", y, wq) + w[mha + "/query/bias"]
            k = torch.einsum("abc,cde->abde", y, w[mha + "/key/kernel"]) + w[mha + "/key/bias"]
            v = torch.einsum("abc,cde->abde", y, w[mha + "/value/kernel"]) + w[mha + "/value/bias"]
            q = q * (1.0 / math.sqrt(d))
            s = torch.einsum("aecd,abcd->acbe", k, q)
            pr = torch.softmax(s, dim=-1)
            o = torch.einsum("acbe,aecd->abcd", pr, v)
            y = torch.einsum("abcd,cde->abe", o, w[mha + "/attention_output/kernel"]) + w[mha + "/attention_output/bias"]
            x = y + side
            side = x
            y = F.layer_norm(x, (D,), w[ln2 + "/gamma"], w[ln2 + "/beta"], LN_EPSILON)
            for j in range(qn):
                y = act(y @ w[f"MLP_{i + 1}_{j + 1}/kernel"] + w[f"MLP_{i + 1}_{j + 1}/bias"])
            x = y + side
        k_ = 0
        y = x @ w[_kname("dense", k_) + "/kernel"] + w[_kname("dense", k_) + "/bias"]
        k_ += 1
        y = y.reshape(B, SLOTS, -1)
        n, rep = _cfg(cfg, "mlp_head_dense_layers_quantity"), _cfg(cfg, "mlp_head_dense_mish_block_repeats")
        for _ in range(n * rep):
            y = act(y @ w[_kname("dense", k_) + "/kernel"] + w[_kname("dense", k_) + "/bias"])
            k_ += 1
        logits = y @ w["MLP_Head_no_Sigmoid/kernel"] + w["MLP_Head_no_Sigmoid/bias"]
        return logits.numpy()


def weights_to_torch(weights: dict) -> dict:
    """Pre-converts a weight dict once so that repeated timed forward_torch_f32 calls do not pay for it."""
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)) for k, v in weights.items()}


# ------------------------------------------------------------------------------------------------
# bf16-faithful mode: the product's bf16 data path restated with float64 accumulation
# ------------------------------------------------------------------------------------------------
def bf16_round(x) -> np.ndarray:
    """Round to the nearest bfloat16 (ties to even), returned as float64.  Matches cvt.rn.bf16.f32 /
    __float2bfloat16_rn on the float32 value of x (NaN/Inf pass through)."""
    f = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    u = f.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    out = (r & 0xFFFFFFFF).astype(np.uint32).view(np.float32)
    out = np.where(np.isfinite(f), out, f)
    return out.astype(np.float64).reshape(np.shape(x))


def attention_core_bf16(q: np.ndarray, k: np.ndarray, v: np.ndarray) -> np.ndarray:
    """attention_core with the operand roundings of the tensor-core kernel: q, k, v are bf16 values; the scores and the
    softmax statistics are float (float64 here); the probabilities fed to the P.V product are rounded to bf16 while the
    row sum uses the unrounded ones; the result is divided by that sum (csrc/attention_tc8.cu)."""
    d = q.shape[-1]
    s = np.einsum("aecd,abcd->acbe", k, q) * (1.0 / math.sqrt(d))
    s = s - s.max(axis=-1, keepdims=True)
    e = np.exp(s)
    o = np.einsum("acbe,aecd->abcd", bf16_round(e), v)
    return o / np.moveaxis(e.sum(axis=-1), 1, 2)[..., None]


def forward_bf16(weights: dict, cfg, images: np.ndarray, return_intermediates: bool = False):
    """The forward pass with every operand rounding of the product's bf16 mode and float64 accumulation:
    bf16 patches and Dense weights, float32-like (here float64) residual stream, bf16 LayerNorm outputs, bf16 q/k/v and
    attention context, bf16 activations between the MLP / head layers, float weights for the slot projection and the
    final Dense(6).  A tensor-core kernel with a dropped k-tail, a wrong swizzle column or a mis-fused LayerNorm differs
    from this by far more than the accumulation-order noise (~1e-3) that separates a correct one from it."""
    w64 = {k: np.asarray(v, dtype=np.float64) for k, v in weights.items()}
    wb = {k: bf16_round(v) for k, v in weights.items()}
    act = mish if _cfg(cfg, "use_mish", True) else gelu_tanh
    p = _cfg(cfg, "patch_size")
    L, qn = _cfg(cfg, "encoder_repeat_times"), _cfg(cfg, "encoder_mlp_quantities")
    inter = {}
    x = bf16_round(extract_patches(np.asarray(images, dtype=np.float32), p))
    x = x @ wb["linear_projection/kernel"] + w64["linear_projection/bias"]
    x = x + w64["position_encoding/position_embedding/embeddings"][None, :, :]
    inter["embedded_patches"] = x
    for i in range(L):
        ln1, ln2, mha = _kname("layer_normalization", 2 * i), _kname("layer_normalization", 2 * i + 1), _kname("multi_head_attention", i)
        y = bf16_round(layer_norm(x, w64[ln1 + "/gamma"], w64[ln1 + "/beta"]))
        q = bf16_round(np.einsum("abc,cde->abde", y, wb[mha + "/query/kernel"]) + w64[mha + "/query/bias"])
        k = bf16_round(np.einsum("abc,cde->abde", y, wb[mha + "/key/kernel"]) + w64[mha + "/key/bias"])
        v = bf16_round(np.einsum("abc,cde->abde", y, wb[mha + "/value/kernel"]) + w64[mha + "/value/bias"])
        o = bf16_round(attention_core_bf16(q, k, v))
        x = np.einsum("abcd,cde->abe", o, wb[mha + "/attention_output/kernel"]) + w64[mha + "/attention_output/bias"] + x
        y = bf16_round(layer_norm(x, w64[ln2 + "/gamma"], w64[ln2 + "/beta"]))
        for j in range(qn):
            y = act(y @ wb[f"MLP_{i + 1}_{j + 1}/kernel"] + w64[f"MLP_{i + 1}_{j + 1}/bias"])
            if j + 1 < qn:
                y = bf16_round(y)
        x = y + x
        inter[f"block_{i + 1}"] = x
    inter["encoded_images"] = x
    B = x.shape[0]
    kk = 0
    y = bf16_round(x @ w64[_kname("dense", kk) + "/kernel"] + w64[_kname("dense", kk) + "/bias"])      # float weights, bf16 store
    kk += 1
    y = y.reshape(B, SLOTS, -1)
    n, rep = _cfg(cfg, "mlp_head_dense_layers_quantity"), _cfg(cfg, "mlp_head_dense_mish_block_repeats")
    for _ in range(n * rep):
        y = bf16_round(act(y @ wb[_kname("dense", kk) + "/kernel"] + w64[_kname("dense", kk) + "/bias"]))
        kk += 1
    inter["head_last"] = y
    logits = y @ w64["MLP_Head_no_Sigmoid/kernel"] + w64["MLP_Head_no_Sigmoid/bias"]
    if return_intermediates:
        return logits, inter
    return logits
