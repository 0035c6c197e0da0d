"""Host enqueue time vs device time of one forward+decode call at small and large batch (evidence for the launch
strategy: streams + programmatic dependent launch, no graph capture).  Prints one JSON object."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import vision_transformer_detector_b200 as vd  # noqa: E402


def main():
    torch.cuda.set_device(0)
    model = vd.create_vision_transformer_detector(seed=0)
    out = {}
    for B in (1, 2, 8, 64):
        x = torch.rand((B, 608, 608, 3), device="cuda") * 2 - 1
        for _ in range(5):
            model.detect(x)
        torch.cuda.synchronize()
        enq, dev = [], []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            t0 = time.perf_counter()
            model.detect(x)
            enq.append(time.perf_counter() - t0)
            e1.record()
            torch.cuda.synchronize()
            dev.append(e0.elapsed_time(e1))
        enq.sort(); dev.sort()
        out[f"B{B}"] = {"host_enqueue_ms_median": enq[len(enq) // 2] * 1e3, "device_ms_median": dev[len(dev) // 2],
                        "launches": model.launch_count()  // 25}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
