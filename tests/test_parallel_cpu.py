"""world_size-2 gloo test (CPU) of the multi-GPU host logic: sharding, record packing and the all-gather
order.  The per-rank "decode" here is the oracle, standing in for the GPU records, so the test checks the
plumbing only."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _util import ROOT, oracle
from vision_transformer_detector_b200 import DetectionRecords, parallel


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _records_from_logits(logits: np.ndarray) -> DetectionRecords:
    d = oracle.decode(logits.astype(np.float32))
    return DetectionRecords(logits, d["decoded"].astype(np.float32), d["class_id"], d["class_conf"].astype(np.float32),
                            d["keep"].astype(np.uint8), d["corners"])


def _worker(rank: int, world: int, port: int, total: int, out_dir: str):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    logits = (rng.normal(size=(total, 17, 6)) * 2).astype(np.float32)       # same on every rank
    lo, hi = parallel.shard_bounds(total, world, rank)
    rec = _records_from_logits(logits[lo:hi])
    packed = torch.from_numpy(parallel.pack_records(rec))
    gathered = parallel.all_gather_records(packed)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), gathered.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_the_batch():
    for total, world in ((1024, 8), (1024, 2), (10, 4), (3, 8), (64, 1)):
        spans = [parallel.shard_bounds(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(1)
    rec = _records_from_logits((rng.normal(size=(5, 17, 6)) * 3).astype(np.float32))
    u = parallel.unpack_records(parallel.pack_records(rec))
    assert np.array_equal(u["decoded"], rec.decoded) and np.array_equal(u["class_id"], rec.class_id)
    assert np.array_equal(u["keep"], rec.keep) and np.array_equal(u["corners"], rec.corners)
    assert np.array_equal(u["class_conf"], rec.class_conf)


def test_two_rank_gather_equals_single_process(tmp_path):
    total, world = 8, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    gathered = np.load(os.path.join(str(tmp_path), "gathered.npy"))
    rng = np.random.default_rng(7)
    logits = (rng.normal(size=(total, 17, 6)) * 2).astype(np.float32)
    whole = parallel.pack_records(_records_from_logits(logits))
    assert gathered.shape == whole.shape == (total * 17, parallel.RECORD_WIDTH)
    assert np.array_equal(gathered, whole)            # rank order == batch order, bit-exact


def _metric_worker(rank: int, world: int, port: int, total: int, out_dir: str):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import map_oracle
    from _util import map_case
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    metric = map_oracle.MeanAveragePrecision()          # stands in for the GPU metric: same update_state signature
    for step in range(2):
        y_true, y_pred = map_case(60 + step, total, 17)                       # same on every rank
        lo, hi = parallel.shard_bounds(total, world, rank)
        parallel.update_metric_sharded(metric, torch.from_numpy(y_true[lo:hi]), torch.from_numpy(y_pred[lo:hi]),
                                       use_transform_predictions=False)
    np.save(os.path.join(out_dir, f"state_{rank}.npy"), metric.latest_positive_bboxes)
    np.save(os.path.join(out_dir, f"ap_{rank}.npy"), np.array([metric.result()]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_metric_equals_single_process(tmp_path):
    """The metric depends on the ORDER of the images (latest related images per class): sharded evaluation must gather
    the shards in rank order so that every rank ends with the single-process state."""
    import map_oracle
    from _util import map_case
    total, world = 12, 2
    port = _free_port()
    mp.spawn(_metric_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    ref = map_oracle.MeanAveragePrecision()
    for step in range(2):
        y_true, y_pred = map_case(60 + step, total, 17)
        ref.update_state(y_true, y_pred, use_transform_predictions=False)
    for rank in range(world):
        assert np.array_equal(np.load(os.path.join(str(tmp_path), f"state_{rank}.npy")), ref.latest_positive_bboxes)
        assert np.load(os.path.join(str(tmp_path), f"ap_{rank}.npy"))[0] == ref.result()
