#!/usr/bin/env python
"""bench.py — images/sec of the ViT-detector forward + head decode on N B200s.

    python bench.py --gpus 1 --steps K --warmup W            (N = 1)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                (N > 1, one rank per GPU)
    python bench.py --impl reference ...                      (the CPU arm: the oracle restatement of the
                                                               reference on the host cores; TF cannot run here)

A "step" = one pass of the hot path (patches -> encoder -> head -> decode) over one batch of synthetic
images.  N = 1 runs BASELINE.json configs[1] (default config, batch 64, bf16); N > 1 runs configs[2]
(default config, global batch 1024 sharded 1024/N per GPU, detections all-gathered with NCCL).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images_per_sec_fwd_decode"
UNIT = "images/s"
FLOP_PER_IMAGE_DEFAULT = 199.747e9      # SURVEY Appendix A (algorithmic, unpadded)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=0, help="override the per-GPU batch (default: 64 at N=1, 1024/N at N>1)")
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--variant", default="default", choices=["default", "hires", "vitb"],
                    help="default = BASELINE configs[1]/[2]; hires = configs[3]; vitb = configs[4]")
    ap.add_argument("--chunk", type=int, default=0, help="encoder micro-batch (images); 0 = library default")
    ap.add_argument("--cpu-sample", type=int, default=4, help="images per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="also print a per-kernel-category time table to stderr")
    return ap.parse_args()


def variant_config(vd, name):
    if name == "default":
        return vd.DetectorConfig(), "default config (608x608, p17, D28, h8, d40, q8, L8, head 136/7/1, Mish)"
    if name == "hires":
        return (vd.DetectorConfig(input_shape=(1024, 1024, 3), patch_size=16),
                "hi-res variant (1024x1024, p16, 4096 tokens), other knobs default")
    return (vd.DetectorConfig(input_shape=(640, 640, 3), patch_size=16, embedding_dim=768, encoder_num_heads=12,
                              encoder_key_dim=64, encoder_repeat_times=12, encoder_mlp_quantities=3),
            "ViT-B/16-width variant (640x640, p16, D768, h12, d64, L12, q3)")


def flops_per_image(cfg) -> float:
    """Algorithmic FLOPs (2*M*N*K, unpadded) — formulas of SURVEY Appendix A."""
    T, P, D = cfg.tokens, cfg.patch_dim, cfg.embedding_dim
    H, d, L = cfg.encoder_num_heads, cfg.encoder_key_dim, cfg.encoder_repeat_times
    f = 2.0 * T * P * D
    mlp, fan = 0.0, D
    for u in cfg.encoder_mlp_units():
        mlp += 2.0 * T * fan * u
        fan = u
    f += L * (3 * 2.0 * T * D * H * d + 2 * 2.0 * H * T * T * d + 2.0 * T * H * d * D + mlp)
    S = 17
    f += 2.0 * T * D * S
    fan = T
    for u in cfg.head_units():
        f += 2.0 * S * fan * u
        fan = u
    f += 2.0 * S * fan * 6
    return f


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe), during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines: list[str] = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------------------------------------
# the reference arm / CPU baseline: oracle restatement on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_forward_decode_rate(vd, cfg, weights, sample: int, steps: int, warmup: int):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import torch
    import vitdet_oracle as oracle       # the ONLY use of oracle/ in bench.py: the timed CPU baseline
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wt = oracle.weights_to_torch(weights)
    rng = np.random.default_rng(99)
    x = rng.uniform(-1, 1, size=(sample, *cfg.input_shape)).astype(np.float32)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        logits = oracle.forward_torch_f32(wt, cfg, x)
        oracle.decode(logits, image_size=cfg.input_shape[:2])
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return sample * len(times) / total, total / len(times) * 1e3, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import vision_transformer_detector_b200 as vd
    cfg, desc = variant_config(vd, args.variant)
    weights = vd.random_weights(cfg, seed=1, spread=True)
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 1))
    # bounded: each step is `cpu_sample` images of the same workload
    value, ms, cores = cpu_forward_decode_rate(vd, cfg, weights, args.cpu_sample, steps, warmup)
    sample = f"{args.cpu_sample} images per step x {steps} steps of the same workload (f32, torch CPU, all host threads)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic U(-1,1) images, random-init weights (Keras initialisers, spread set)",
        "config": {"workload": desc + "; reference arm = float32 CPU restatement (oracle port; TensorFlow 2.9 is not installable in this image)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(torch, local_rank: int):
    """Pins this rank to the CPUs local to its GPU (sysfs local_cpulist of the GPU's PCI function) so that the pinned
    host buffers it allocates next are first-touched on the GPU's NUMA node: with 8 ranks each copying 568 MB per step
    the H2D path otherwise crosses the socket interconnect.  Best effort; returns the cpu list or None."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import vision_transformer_detector_b200 as vd
    from vision_transformer_detector_b200 import parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints a version banner on STDOUT when the communicator is created: send fd 1 to stderr until the
        # first collective has run, so that stdout carries exactly one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    cfg, desc = variant_config(vd, args.variant)
    if args.batch_per_gpu:
        B = args.batch_per_gpu
    elif args.variant == "default":
        B = 64 if world == 1 else 1024 // world
    else:
        B = 32
    S = 17
    model = vd.VisionTransformerDetector(cfg, seed=None, compute_mode=args.mode)
    model.set_weights(vd.random_weights(cfg, seed=1, spread=True))
    if args.chunk:
        model.set_chunk(args.chunk)

    # synthetic inputs: per-rank seed, resident in HBM for `value`, in pinned host memory for `e2e`
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    x_dev = torch.rand((B, *cfg.input_shape), generator=g, device=dev, dtype=torch.float32) * 2 - 1
    img_size = cfg.input_shape[:2]
    rec_bytes = S * (24 + 4 + 4 + 1 + 16)      # decoded, class_id, class_conf, keep, corners per image
    assert parallel.RECORD_WIDTH == 13

    def step():
        rec = model.detect(x_dev, image_size=img_size)
        if world > 1:
            # the path's only exchange: fixed-size detection records of every rank, all-gathered (NCCL)
            return parallel.all_gather_records(parallel.pack_records(rec))
        return rec

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    names = model.profile_categories()
    dominant = "gemm_mlp_2" if "gemm_mlp_2" in names else names[-1]
    model.launch_count(reset=True)
    model.profile_read(reset=True)
    model.profile_enable([dominant])
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    launches = model.launch_count(reset=True)      # kernels of this library only (NCCL / torch packing not counted)
    prof = model.profile_read(reset=True)
    model.profile_enable(None)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- e2e: the reference-facing call with HOST buffers (H2D + forward + decode + D2H inside) ----
    e2e = None
    e2e_u8 = None
    if not args.no_e2e:
        def time_host_path(x_np, bytes_per_elem):
            for _ in range(2):
                model.detect(x_np, image_size=img_size)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                rec = model.detect(x_np, image_size=img_size)     # vitdet_predict_host[_u8]: synchronous, returns numpy records
                if world > 1:
                    parallel.all_gather_records(torch.from_numpy(parallel.pack_records(rec)).to(dev))
            e1.record()
            barrier()
            wall = time.perf_counter() - t0
            tt = torch.tensor([max(e0.elapsed_time(e1) * 1e-3, wall)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return {"value": world * B * args.steps / float(tt.item()), "unit": UNIT,
                    "h2d_bytes_per_step": int(world * B * np.prod(cfg.input_shape) * bytes_per_elem),
                    "d2h_bytes_per_step": int(world * B * (S * 24 + rec_bytes)),
                    "host_affinity": (f"{len(numa_cpus)} cpus local to the GPU" if numa_cpus else "unbound")}

        x_host = torch.empty((B, *cfg.input_shape), dtype=torch.float32, pin_memory=True)
        x_host.copy_(x_dev)
        e2e = time_host_path(x_host.numpy(), 4)
        # the same call fed with the uint8 pixels the reference's input pipeline starts from
        # (vision_transformer_utilities.py:446-447: x / 127.5 - 1, here inside the patch kernel); extra, not the headline
        u_host = torch.empty((B, *cfg.input_shape), dtype=torch.uint8, pin_memory=True)
        u_host.copy_(((x_dev + 1) * 127.5).round().clamp(0, 255).to(torch.uint8))
        e2e_u8 = time_host_path(u_host.numpy(), 1)
        e2e_u8["input"] = "uint8 NHWC pixels, normalised on the device"
        del x_host, u_host

    # ---- roofline of the dominant kernel (the 3584 -> 1792 MLP GEMM in the default config) ----
    roofline = None
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak_tf, peak_src, pk_burst = 1590.0, "fallback (B200_PROFILING.md)", 1590.0
    if os.path.exists(peaks_path):
        pk = json.load(open(peaks_path))
        # the kernel is timed inside a long step -> sustained figure
        peak_tf, peak_src = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1590.0))), "measured sustained (MEASURED_PEAKS.json)"
        pk_burst = float(pk.get("bf16_tflops", peak_tf))
    if dominant in prof and args.mode == "bf16":
        ms_k, n_k = prof[dominant]
        units = cfg.encoder_mlp_units()
        K_, N_ = (units[0], units[1]) if len(units) > 1 else (cfg.embedding_dim, units[0])
        chunk = args.chunk or 64
        rows_total = B * cfg.tokens * args.steps * cfg.encoder_repeat_times     # rows pushed through this layer in the timed region
        flops_total = 2.0 * rows_total * K_ * N_
        achieved = flops_total / (ms_k * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "gemm_mlp_2_traffic.json")
        if os.path.exists(tpath) and args.variant == "default" and B == 64:
            traffic = json.load(open(tpath)).get("traffic_bytes_per_launch")      # from the committed ncu --set full capture
        roofline = {"bound": "tensor", "kernel": f"{'gemm_tc2_kernel (CTA pair)' if K_ >= 512 and N_ >= 128 else 'gemm_tc_kernel'}[{dominant}: K={K_} -> N={N_}, bias+Mish epilogue]",
                    "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                    "peak_source": peak_src, "peak_burst": pk_burst, "frac_of_burst": (achieved / pk_burst if pk_burst else None), "launches": n_k, "avg_launch_ms": ms_k / n_k,
                    "share_of_step": ms_k / (ms_step * args.steps), "traffic": traffic,
                    "algorithmic_bytes": 2.0 * (B * cfg.tokens * (K_ + N_) + K_ * N_)}

    breakdown = None
    if args.breakdown:
        model.profile_enable(names)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        breakdown = {k: {"ms_per_step": v[0] / 2, "launches_per_step": v[1] // 2} for k, v in model.profile_read().items()}
        model.profile_enable(None)

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            weights = {w.name[:-2]: w.numpy() for w in model.weights}
            v, ms, cores = cpu_forward_decode_rate(vd, cfg, weights, args.cpu_sample, 3, 1)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{args.cpu_sample} images per step x 3 steps of the same workload (float32 oracle port of the reference on torch CPU; TF 2.9 not installable)"}
        fpi = flops_per_image(cfg)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak" if (world == 1 or args.batch_per_gpu) else "strong",
            "vs_baseline": None, "dtype": args.mode, "data": "synthetic U(-1,1) images, random-init weights (Keras initialisers, spread set)",
            "config": {"workload": f"{desc}; batch {B}/GPU x {world} GPU(s); forward + fused decode"
                                   + ("; NCCL all-gather of detection records" if world > 1 else ""),
                       "global_batch": B * world, "tokens": cfg.tokens, "flop_per_image": fpi,
                       "l2_policy": "inputs and activations (>= 283 MB per step) exceed the 126 MB L2; no flush between steps"},
            "model_tflops": value * fpi / 1e12,
            "e2e": e2e, "e2e_uint8": e2e_u8, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }
        if breakdown:
            line["breakdown"] = breakdown
            print(json.dumps(breakdown, indent=1), file=sys.stderr)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
