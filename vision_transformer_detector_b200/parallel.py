"""Data-parallel plumbing of the path: images are independent, so the batch is sharded across
one process per GPU with no data-path collective; the only exchange is the all-gather of the fixed-size
per-image detection records.  torch.distributed carries it (NCCL on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np

RECORD_WIDTH = 13          # decoded[6] | class_id | class_conf | keep | corners[4], all as float32 (ints are exact < 2^24)


def shard_bounds(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of `total` images for `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_records(rec):
    """DetectionRecords (torch tensors or numpy arrays, shapes (B,17,...)) -> float32 [B*17, 13]."""
    if hasattr(rec.decoded, "is_cuda"):
        import torch
        B, S = rec.class_id.shape
        f = torch.float32
        return torch.cat([rec.decoded.reshape(B * S, 6), rec.class_id.reshape(B * S, 1).to(f),
                          rec.class_conf.reshape(B * S, 1), rec.keep.reshape(B * S, 1).to(f),
                          rec.corners.reshape(B * S, 4).to(f)], dim=1)
    B, S = rec.class_id.shape
    return np.concatenate([rec.decoded.reshape(B * S, 6), rec.class_id.reshape(B * S, 1).astype(np.float32),
                           rec.class_conf.reshape(B * S, 1), rec.keep.reshape(B * S, 1).astype(np.float32),
                           rec.corners.reshape(B * S, 4).astype(np.float32)], axis=1).astype(np.float32)


def unpack_records(packed, slots: int = 17) -> dict:
    """Inverse of pack_records on a gathered [N*17, 13] block -> dict of (N,17,...) arrays/tensors."""
    n = packed.shape[0] // slots
    if hasattr(packed, "is_cuda"):
        import torch
        return {"decoded": packed[:, 0:6].reshape(n, slots, 6), "class_id": packed[:, 6].to(torch.int32).reshape(n, slots),
                "class_conf": packed[:, 7].reshape(n, slots), "keep": packed[:, 8].to(torch.uint8).reshape(n, slots),
                "corners": packed[:, 9:13].to(torch.int32).reshape(n, slots, 4)}
    return {"decoded": packed[:, 0:6].reshape(n, slots, 6), "class_id": packed[:, 6].astype(np.int32).reshape(n, slots),
            "class_conf": packed[:, 7].reshape(n, slots), "keep": packed[:, 8].astype(np.uint8).reshape(n, slots),
            "corners": packed[:, 9:13].astype(np.int32).reshape(n, slots, 4)}


class RecordGather:
    """The C-ABI form of the exchange: vitdet_gather_detections (csrc/gather.cu) on a NCCL communicator that the library
    creates itself — rank 0 draws the unique id, torch.distributed (any backend) broadcasts its 128 bytes.  With the
    packed record written by the head-tail kernel (model.detect(..., packed=True)) the timed multi-GPU step contains no
    eager PyTorch kernel."""

    def __init__(self, device):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _capi
        self._lib = _capi.load()
        self.device = torch.device(device)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        buf = C.create_string_buffer(128)
        if self.rank == 0:
            _capi.check(self._lib.vitdet_nccl_unique_id(buf))
        t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=self.device if dist.get_backend() == "nccl" else "cpu")
        dist.broadcast(t, 0)
        uid = bytes(t.cpu().tolist())
        self._comm = C.c_void_p()
        with torch.cuda.device(self.device):
            _capi.check(self._lib.vitdet_nccl_comm_create(uid, self.rank, self.world, C.byref(self._comm)))

    def all_gather(self, packed):
        """packed: float32 CUDA tensor [rows, 13] of this rank -> [world * rows, 13] in rank order (asynchronous on
        torch's current stream)."""
        import ctypes as C
        import torch
        from . import _capi
        packed = packed.contiguous()
        out = torch.empty((self.world * packed.shape[0], packed.shape[1]), dtype=torch.float32, device=packed.device)
        with torch.cuda.device(packed.device):
            _capi.check(self._lib.vitdet_gather_detections(self._comm, C.c_void_p(packed.data_ptr()), int(packed.shape[0]),
                                                           C.c_void_p(out.data_ptr()),
                                                           C.c_void_p(torch.cuda.current_stream(packed.device).cuda_stream)))
        return out

    def close(self):
        if getattr(self, "_comm", None) is not None and self._comm:
            self._lib.vitdet_nccl_comm_destroy(self._comm)
            self._comm = None


def all_gather_records(packed):
    """All-gathers equally sized packed record blocks of every rank, in rank order (one collective)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    out = torch.empty((world * packed.shape[0], packed.shape[1]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed.contiguous())
    return out


def update_metric_sharded(metric, y_true_local, y_pred_local, use_transform_predictions: bool = True) -> None:
    """Evaluation with the batch sharded over ranks: the metric keeps the LATEST related images per class in dataset
    order (det.py:1427, 1852-1857), so the shards are all-gathered in rank order (= the contiguous shard_bounds order)
    and every rank applies the same update to its own replica of the state; `metric.result()` is then identical on
    all ranks and to a single-process run over the whole batch.  Shards must be equally sized (pad the last batch).
    y_*_local: (B_local, slots, 6) torch tensors on the rank's device (CUDA with NCCL, CPU with gloo)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        metric.update_state(y_true_local, y_pred_local, use_transform_predictions=use_transform_predictions)
        return
    world = dist.get_world_size()
    both = torch.stack([torch.as_tensor(y_true_local, dtype=torch.float32),
                        torch.as_tensor(y_pred_local, dtype=torch.float32).to(y_true_local.device)]).contiguous()   # (2, B_local, S, 6)
    flat = torch.empty((world * 2, *both.shape[1:]), dtype=both.dtype, device=both.device)
    dist.all_gather_into_tensor(flat, both)
    out = flat.view(world, *both.shape)
    y_true = out[:, 0].reshape(-1, *both.shape[2:])          # rank-major = global image order
    y_pred = out[:, 1].reshape(-1, *both.shape[2:])
    if not y_true.is_cuda:                                   # gloo / CPU tensors: hand numpy to the metric
        y_true, y_pred = y_true.numpy(), y_pred.numpy()
    metric.update_state(y_true, y_pred, use_transform_predictions=use_transform_predictions)
