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


def map_case(seed, batch, slots, classes_used=(3, 17, 79), exact_class=0.5, max_labels=6, max_extra=6, image=608.0):
    """Seeded (y_true, y_pred) pair of already-decoded (batch, slots, 6) rows for the evaluation metric: per image a few
    labelled boxes of a few classes, predictions that jitter the labels (so IoUs land on both sides of 0.5 and of the
    ten thresholds), spurious predictions, sub-threshold ones, and class values either exactly on an integer
    (equal confidences -> the tie rule matters) or off by up to 0.3 (confidence below / above 0.5)."""
    rng = np.random.default_rng(seed)
    y_true = np.full((batch, slots, 6), -8.0, np.float32)
    y_true[..., 0] = 0
    y_pred = np.zeros((batch, slots, 6), np.float32)
    y_pred[..., 1] = rng.uniform(0, 79, (batch, slots))
    y_pred[..., 2:] = rng.uniform(0, image, (batch, slots, 4))
    y_pred[..., 0] = rng.uniform(0, 0.5, (batch, slots))          # not positive unless overwritten below
    for b in range(batch):
        nl = int(rng.integers(0, min(max_labels, slots) + 1))
        label_slots = rng.choice(slots, nl, replace=False)
        free = [s for s in rng.permutation(slots)]
        for s in label_slots:
            c = float(rng.choice(classes_used))
            box = [rng.uniform(50, image - 50), rng.uniform(50, image - 50), rng.uniform(20, 200), rng.uniform(20, 200)]
            if rng.random() < 0.2:
                box[2:] = [64.0, 64.0]                             # equal areas: the stable-sort rule decides
            y_true[b, s] = [1.0, c, *box]
            if rng.random() < 0.8 and free:                        # a prediction near this label
                ps = free.pop()
                jitter = rng.choice([0.0, 0.02, 0.1, 0.25, 0.6])
                pbox = [box[0] + rng.normal() * jitter * box[3], box[1] + rng.normal() * jitter * box[2],
                        box[2] * (1 + rng.normal() * jitter * 0.5), box[3] * (1 + rng.normal() * jitter * 0.5)]
                cls = c if rng.random() < exact_class else c + rng.uniform(-0.3, 0.3)
                y_pred[b, ps] = [rng.uniform(0.4, 1.0), np.clip(cls, 0, 79), *np.clip(pbox, 0, image)]
        for _ in range(int(rng.integers(0, max_extra + 1))):       # predictions with no label behind them
            if not free:
                break
            ps = free.pop()
            c = float(rng.choice(classes_used))
            cls = c if rng.random() < exact_class else c + rng.uniform(-0.3, 0.3)
            y_pred[b, ps] = [rng.uniform(0.4, 1.0), np.clip(cls, 0, 79), rng.uniform(0, image), rng.uniform(0, image),
                             rng.uniform(10, 200), rng.uniform(10, 200)]
    return y_true, y_pred


TF_GOLDEN = os.path.join(ROOT, "tests", "golden", "tf_forward.npz")


def load_tf_golden(path=TF_GOLDEN) -> dict:
    """Reads a file written by tools/make_tf_golden.py (outputs of the unmodified reference under TensorFlow 2.9)."""
    import json
    out = {}
    with np.load(path, allow_pickle=False) as z:
        for name in json.loads(str(z["cases"])):
            names = json.loads(str(z[f"{name}/weight_names"]))
            kwargs = json.loads(str(z[f"{name}/kwargs"]))
            if "input_shape" in kwargs:
                kwargs["input_shape"] = tuple(kwargs["input_shape"])
            taps = {k.split("/tap/")[1]: z[k] for k in z.files if k.startswith(f"{name}/tap/")}
            out[name] = dict(kwargs=kwargs, weight_names=names, weights=[z[f"{name}/w/{i}"] for i in range(len(names))],
                             images=z[f"{name}/images"], logits=z[f"{name}/logits"], decoded=z[f"{name}/decoded"], taps=taps)
    return out
