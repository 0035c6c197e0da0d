#!/usr/bin/env python
"""Layer-by-layer error growth of the CUDA path against the float64 oracle ("measure the error growth layer by
layer"): residual stream after the patch embedding and after every encoder block, the head's last activation and the
logits, for the fp32-accumulate mode (tensor-core split-bf16 and CUDA-core IEEE kernels) and the bf16 mode; the bf16 mode
also against the bf16-faithful oracle (same operand roundings, float64 accumulation).  Uses the library's debug taps.
Writes profiles/<tag>_error_growth.md.  The oracle is the checker here, as in the tests."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import vitdet_oracle as oracle
import vision_transformer_detector_b200 as vd


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rms(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(np.sqrt(np.mean(b ** 2)), 1e-30))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    cfg = vd.DetectorConfig()
    rng = np.random.default_rng(1234)
    x = rng.uniform(-1, 1, size=(1, *cfg.input_shape)).astype(np.float32)
    names = ["embedded_patches"] + [f"block_{i + 1}" for i in range(cfg.encoder_repeat_times)] + ["head_last"]
    out = {"config": "default (608x608, p17, D28, h8, d40, q8, L8, head 136/7/1, Mish), 1 image, U(-1,1)", "weight_sets": {}}
    for label, seed, spread in (("keras_default_init", 0, False), ("spread", 1, True)):
        w = vd.random_weights(cfg, seed=seed, spread=spread)
        t0 = time.time()
        ref, inter = oracle.forward(w, cfg, x, np.float64, return_intermediates=True)
        refb, interb = oracle.forward_bf16(w, cfg, x, return_intermediates=True)
        f32 = oracle.forward_torch_f32(w, cfg, x)
        rows = {"oracle_seconds": time.time() - t0, "float32_cpu_oracle_vs_float64": {"logits": rel(f32, ref)},
                "bf16_faithful_oracle_vs_float64": {n: rel(interb[n], inter[n]) for n in names} | {"logits": rel(refb, ref)}}
        for mode, opt in (("fp32_tensor_core", ("fp32", 1)), ("fp32_cuda_core", ("fp32", 0)), ("bf16", ("bf16", None))):
            m = vd.VisionTransformerDetector(cfg, seed=None, compute_mode=opt[0])
            m.set_weights(w)
            if opt[1] is not None:
                m.set_option("fp32_tc", opt[1])
            m.debug_taps(True)
            got = m.predict(x)
            taps = {n: m.debug_read(n, 1) for n in names}
            rows[mode + "_vs_float64"] = {n: {"max": rel(taps[n], inter[n]), "rms": rms(taps[n], inter[n])} for n in names}
            rows[mode + "_vs_float64"]["logits"] = {"max": rel(got, ref), "rms": rms(got, ref)}
            if mode == "bf16":
                rows["bf16_vs_bf16_faithful"] = {n: {"max": rel(taps[n], interb[n]), "rms": rms(taps[n], interb[n])} for n in names}
                rows["bf16_vs_bf16_faithful"]["logits"] = {"max": rel(got, refb), "rms": rms(got, refb)}
            m.close()
        out["weight_sets"][label] = rows
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{tag}_error_growth.json"), "w"), indent=1)
    lines = [f"# {tag}: error growth through the model (max|gpu - ref| / max|ref| per tensor; RMS ratio in brackets)", "",
             out["config"], "", "Reference = float64 numpy oracle (`oracle/vitdet_oracle.py`); `bf16 vs faithful` = against the oracle that rounds",
             "operands and stored activations to bf16 exactly where the kernels do and accumulates in float64.", ""]
    for label, rows in out["weight_sets"].items():
        lines += [f"## weights: {label}", "", "| tensor | fp32 mode, tensor cores (split bf16 x3) | fp32 mode, CUDA cores (IEEE) | bf16 mode | bf16 vs faithful | faithful oracle vs f64 |", "|---|---:|---:|---:|---:|---:|"]
        for n in names + ["logits"]:
            f = lambda d: f"{d[n]['max']:.2e} ({d[n]['rms']:.1e})"
            lines.append(f"| {n} | {f(rows['fp32_tensor_core_vs_float64'])} | {f(rows['fp32_cuda_core_vs_float64'])} | {f(rows['bf16_vs_float64'])} | "
                         f"{f(rows['bf16_vs_bf16_faithful'])} | {rows['bf16_faithful_oracle_vs_float64'][n]:.2e} |")
        lines += ["", f"float32 CPU oracle (torch) vs float64 on the logits: {rows['float32_cpu_oracle_vs_float64']['logits']:.2e}", ""]
    path = os.path.join(ROOT, "gpurun_out", f"{tag}_error_growth.md")
    open(path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
