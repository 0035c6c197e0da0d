"""Pinning the forward pass to the real reference (Keras / TensorFlow 2.9).

tools/make_tf_golden.py, run where the reference's own stack exists, writes tests/golden/tf_forward.npz: weights (with
their Keras names in model.weights order), inputs, logits, transform_predictions output and the residual-stream taps of
the UNMODIFIED reference.  This container cannot produce that file (no TensorFlow for Python 3.12, no network), so:

  * when the file is present, the oracle (CPU) and the CUDA path (GPU) are checked against it;
  * while it is absent the forward pass is PARITY UNPINNED — the tests say so and skip;
  * the consumer itself is always exercised on a stand-in file of the same format, written by the tool's own `dump`
    from the second, independently written restatement (oracle/vitdet_oracle_torchnn.py).
"""
import importlib.util
import os

import numpy as np
import pytest

from _util import ROOT, TF_GOLDEN, TOL_BF16, TOL_FP32, load_tf_golden, oracle, rel_err, tiny_config, images
import vision_transformer_detector_b200 as vd

UNPINNED = ("PARITY UNPINNED: tests/golden/tf_forward.npz is absent — run tools/make_tf_golden.py on a TensorFlow 2.9 "
            "machine to pin the forward pass to the reference")


def _tool():
    spec = importlib.util.spec_from_file_location("make_tf_golden", os.path.join(ROOT, "tools", "make_tf_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def check_case_against_oracle(case: dict, tol: float = 1e-4) -> None:
    """What a reference dump pins: the Keras variable order and names, the forward pass, transform_predictions."""
    cfg = vd.DetectorConfig(**case["kwargs"])
    specs = vd.weight_specs(cfg)
    assert [n + ":0" for n, _ in specs] == list(case["weight_names"]), "Keras model.weights order / names differ from the weight table"
    assert [tuple(w.shape) for w in case["weights"]] == [s for _, s in specs]
    w = {n: a for (n, _), a in zip(specs, case["weights"])}
    logits, inter = oracle.forward(w, cfg, case["images"], np.float64, return_intermediates=True)
    assert rel_err(logits, case["logits"]) < tol
    for name, val in case["taps"].items():
        assert rel_err(inter[name], val) < tol, name
    # the reference scales by Constants.MODEL_IMAGE_SIZE whatever the model's input size is (det.py:637-640)
    dec = oracle.transform_predictions(np.asarray(case["logits"], np.float64))
    assert np.abs(dec - case["decoded"]).max() < 608 * 2e-6


def _standin_case(cfg_kw, batch, seed):
    import vitdet_oracle_torchnn as second
    cfg = tiny_config(**cfg_kw)
    w = vd.random_weights(cfg, seed=seed, spread=True)
    x = images(cfg, batch)
    logits = second.forward(w, cfg, x)
    _, inter = oracle.forward(w, cfg, x, np.float64, return_intermediates=True)
    kwargs = {k: getattr(cfg, k) for k in ("input_shape", "patch_size", "embedding_dim", "encoder_num_heads", "encoder_key_dim",
                                           "encoder_mlp_quantities", "encoder_repeat_times", "mlp_head_last_units",
                                           "mlp_head_dense_layers_quantity", "mlp_head_dense_mish_block_repeats", "use_mish")}
    specs = vd.weight_specs(cfg)
    taps = {k: v for k, v in inter.items() if k == "embedded_patches" or k.startswith("block_")}
    return dict(kwargs={k: (list(v) if isinstance(v, tuple) else v) for k, v in kwargs.items()}, weight_names=[n + ":0" for n, _ in specs],
                weights=[w[n] for n, _ in specs], images=x, logits=logits, decoded=oracle.transform_predictions(logits), taps=taps)


def test_consumer_on_a_standin_file(tmp_path):
    path = str(tmp_path / "standin.npz")
    _tool().dump(path, {"tiny": _standin_case({}, 3, 5), "gelu": _standin_case(dict(use_mish=False, input_shape=(136, 68, 3)), 2, 6)})
    cases = load_tf_golden(path)
    assert sorted(cases) == ["gelu", "tiny"]
    for c in cases.values():
        check_case_against_oracle(c)
    # a permuted weight order (what an unverified positional get_weights() assumption would look like) is caught
    bad = dict(cases["tiny"])
    bad["weight_names"] = list(reversed(bad["weight_names"]))
    with pytest.raises(AssertionError):
        check_case_against_oracle(bad)


def test_oracle_reproduces_the_reference_dump():
    if not os.path.exists(TF_GOLDEN):
        print(UNPINNED)
        pytest.skip(UNPINNED)
    for name, c in load_tf_golden().items():
        check_case_against_oracle(c)


@pytest.mark.gpu
def test_cuda_path_reproduces_the_reference_dump():
    if not os.path.exists(TF_GOLDEN):
        print(UNPINNED)
        pytest.skip(UNPINNED)
    for name, c in load_tf_golden().items():
        cfg = vd.DetectorConfig(**c["kwargs"])
        for mode, tol in (("fp32", TOL_FP32), ("bf16", TOL_BF16)):
            m = vd.VisionTransformerDetector(cfg, seed=None, compute_mode=mode)
            m.set_weights(c["weights"])                       # positional, exactly what the reference's get_weights() returned
            assert rel_err(m.predict(c["images"]), c["logits"]) < tol, (name, mode)
            m.close()
        assert np.abs(vd.transform_predictions(c["logits"]) - c["decoded"]).max() < 608 * 2e-6
