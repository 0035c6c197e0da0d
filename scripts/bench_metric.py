"""Times the evaluation metric (update_state + result) on the GPU against the CPU oracle on the same seeded batches.
Not part of bench.py's contract line; prints one JSON object.  The reference's own figure for this step is 5-8 s per
8-image batch (eager TF loops)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import torch  # noqa: E402

from _util import map_case  # noqa: E402
import map_oracle  # noqa: E402
from vision_transformer_detector_b200 import MeanAveragePrecision  # noqa: E402


def main(batch=64, steps=20, cpu_steps=2):
    cases = [map_case(1000 + i, batch, 17, classes_used=tuple(range(80)), max_labels=8, max_extra=8) for i in range(steps)]
    dev_cases = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in cases]
    m = MeanAveragePrecision()
    for a, b in dev_cases[:3]:
        m.update_state(a, b, use_transform_predictions=False)
    m.result()
    m.reset_state()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for a, b in dev_cases:
        m.update_state(a, b, use_transform_predictions=False)
    e1.record()
    t0 = time.perf_counter()
    gpu_result = float(m.result())
    t_result = time.perf_counter() - t0
    e2.record()
    torch.cuda.synchronize()
    upd_ms = e0.elapsed_time(e1) / steps
    # host arrays through the C ABI (copies + synchronise inside)
    mh = MeanAveragePrecision()
    t0 = time.perf_counter()
    for a, b in cases:
        mh.update_state(a, b, use_transform_predictions=False)
    host_ms = (time.perf_counter() - t0) * 1e3 / steps
    ref = map_oracle.MeanAveragePrecision()
    t0 = time.perf_counter()
    for a, b in cases[:cpu_steps]:
        ref.update_state(a, b, use_transform_predictions=False)
    cpu_upd = (time.perf_counter() - t0) / cpu_steps
    t0 = time.perf_counter()
    ref.result()
    cpu_res = time.perf_counter() - t0
    print(json.dumps({
        "workload": f"MeanAveragePrecision, {batch} images x 17 slots per update_state, all 80 classes in play, L=3, K=14",
        "gpu_update_ms_device_tensors": upd_ms, "gpu_update_ms_host_arrays": host_ms, "gpu_result_ms_incl_d2h_sync": t_result * 1e3,
        "gpu_images_per_s_update": batch / (upd_ms * 1e-3),
        "cpu_oracle_update_s": cpu_upd, "cpu_oracle_result_s": cpu_res, "cpu_oracle_images_per_s_update": batch / cpu_upd,
        "reference_quoted": "5-8 s per 8-image evaluation step (forward + eager mAP), notebook cell 7",
        "map_after_all_updates": gpu_result, "kernel_launches": m.launch_count(),
    }))


if __name__ == "__main__":
    main()
