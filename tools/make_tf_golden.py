#!/usr/bin/env python
"""Pin the forward pass to the REAL reference: run this on a machine with the reference's own stack
(Python 3.10, tensorflow 2.9.x, tensorflow_addons 0.17/0.18, opencv, pillow — what the notebook's first cell
installs), then commit the file it writes:

    python tools/make_tf_golden.py --reference /path/to/vision_transformer_detector --out tests/golden/tf_forward.npz

It imports the UNMODIFIED reference module, builds models with create_vision_transformer_detector
(vision_transformer_detector.py:498-583), gives them a seeded "spread" weight set (Keras-default kernels x 3, biases
U(-0.5, 0.5): with the pure default initialisation every logit is ~0 and the check would be weak), runs model.predict on
seeded U(-1, 1) images and transform_predictions (vision_transformer_detector.py:586-647) on the logits, and stores per case

    <case>/kwargs          JSON of the create_... keyword arguments
    <case>/weight_names    the Keras variable names, in model.weights order (pins the positional order of get_weights())
    <case>/w/<index>       every array of model.get_weights()
    <case>/images          float32 (B, H, W, 3)
    <case>/logits          float32 (B, 17, 6)   model.predict(images)
    <case>/decoded         float32 (B, 17, 6)   transform_predictions(logits)
    <case>/tap/<name>      float32 outputs of the keras.layers.Add layers (embedded patches, every residual sum)

tests/test_tf_golden.py consumes the file when it exists: the CPU oracle must reproduce logits / decoded / taps, the weight
name table must equal the one the C library enumerates, and (on a B200) the CUDA path must match within the stated
tolerances.  While the file is absent the forward pass stays "PARITY UNPINNED" and the test says so.

This container cannot run it (no TensorFlow for Python 3.12, no network) — see DESIGN.md §2.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys

import numpy as np

CASES = {
    # the reference defaults (det.py:498-506), one 608x608 image pair
    "default": (dict(), 2),
    # small configurations that exercise SAME padding on both axes, a ragged token count, GELU, repeated head blocks
    "tiny": (dict(input_shape=(60, 130, 3), encoder_num_heads=2, encoder_mlp_quantities=3, encoder_repeat_times=2,
                  mlp_head_last_units=8, mlp_head_dense_layers_quantity=2), 5),
    "tiny_gelu": (dict(input_shape=(136, 68, 3), encoder_num_heads=3, encoder_key_dim=24, encoder_mlp_quantities=3,
                       encoder_repeat_times=2, mlp_head_last_units=8, mlp_head_dense_layers_quantity=2,
                       mlp_head_dense_mish_block_repeats=2, use_mish=False), 3),
}


def spread(weights, names, rng):
    out = []
    for w, n in zip(weights, names):
        if n.endswith("bias:0") or n.endswith("beta:0"):
            out.append(rng.uniform(-0.5, 0.5, size=w.shape).astype(np.float32))
        elif n.endswith("gamma:0"):
            out.append((w + rng.uniform(-0.3, 0.3, size=w.shape)).astype(np.float32))
        else:
            out.append((w * 3).astype(np.float32))
    return out


def dump(path: str, cases: dict) -> None:
    """cases: {name: dict(kwargs, weight_names, weights(list), images, logits, decoded, taps(dict))}.  Also used by the
    test-suite to write a stand-in file and exercise the consumer without TensorFlow."""
    flat = {"cases": np.array(json.dumps(sorted(cases)))}
    for name, c in cases.items():
        flat[f"{name}/kwargs"] = np.array(json.dumps(c["kwargs"]))
        flat[f"{name}/weight_names"] = np.array(json.dumps(list(c["weight_names"])))
        for i, w in enumerate(c["weights"]):
            flat[f"{name}/w/{i}"] = np.asarray(w, np.float32)
        for key in ("images", "logits", "decoded"):
            flat[f"{name}/{key}"] = np.asarray(c[key], np.float32)
        for t, v in c.get("taps", {}).items():
            flat[f"{name}/tap/{t}"] = np.asarray(v, np.float32)
    np.savez_compressed(path, **flat)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="directory holding the unmodified vision_transformer_detector.py")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tf_forward.npz"))
    ap.add_argument("--cases", default=",".join(CASES))
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    import tensorflow as tf
    from tensorflow import keras
    det = importlib.import_module("vision_transformer_detector")
    out = {}
    for name in args.cases.split(","):
        kwargs, batch = CASES[name]
        model = det.create_vision_transformer_detector(**kwargs)
        names = [v.name for v in model.weights]
        rng = np.random.default_rng(sum(map(ord, name)))
        weights = spread(model.get_weights(), names, rng)
        model.set_weights(weights)
        shape = model.input_shape[1:]
        images = rng.uniform(-1, 1, size=(batch, *shape)).astype(np.float32)
        logits = model.predict(images, batch_size=batch, verbose=0)
        # transform_predictions scales by Constants.MODEL_IMAGE_SIZE (608, 608) whatever the model's input size is
        decoded = det.transform_predictions(tf.constant(logits)).numpy()
        adds = [l for l in model.layers if isinstance(l, keras.layers.Add)]
        tap_model = keras.Model(model.inputs, [l.output for l in adds])
        tap_vals = tap_model.predict(images, batch_size=batch, verbose=0)
        taps = {"embedded_patches": tap_vals[0]}
        for i in range((len(adds) - 1) // 2):
            taps[f"block_{i + 1}"] = tap_vals[2 + 2 * i]       # the second Add of every encoder block (det.py:408-412)
        out[name] = dict(kwargs={k: (list(v) if isinstance(v, tuple) else v) for k, v in kwargs.items()}, weight_names=names,
                         weights=weights, images=images, logits=logits, decoded=decoded, taps=taps)
        print(f"{name}: {len(names)} weights, logits {logits.shape}, |logits|max {np.abs(logits).max():.4f}, tf {tf.__version__}")
    dump(args.out, out)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
