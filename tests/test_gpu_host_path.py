"""GPU tests of the round-2 boundary additions: asynchronous host submissions, the packed detection record, per-handle
options, debug taps, and the IEEE (CUDA-core) form of the fp32 mode."""
import os

import numpy as np
import pytest

from _util import TOL_FP32, build_model, images, oracle, rel_err, tiny_config
import vision_transformer_detector_b200 as vd
from vision_transformer_detector_b200 import _capi, parallel

pytestmark = pytest.mark.gpu


def _same(a, b):
    for f in ("logits", "decoded", "class_id", "class_conf", "keep", "corners"):
        x, y = getattr(a, f), getattr(b, f)
        x = x.cpu().numpy() if hasattr(x, "cpu") else x
        y = y.cpu().numpy() if hasattr(y, "cpu") else y
        assert np.array_equal(x, y, equal_nan=True), f


def test_submit_collect_keeps_two_batches_in_flight():
    """vitdet_submit_host / vitdet_collect: results equal the synchronous call, tickets can be collected in any order,
    a third submission while two are in flight fails loudly, page-locked and pageable buffers give the same records."""
    import torch
    cfg = tiny_config()
    m = build_model(cfg, vd.random_weights(cfg, seed=2, spread=True), "bf16")
    xa, xb = images(cfg, 7, seed=1), images(cfg, 5, seed=2)
    ra, rb = m.detect(xa), m.detect(xb)
    ta = m.submit(xa)
    tb = m.submit(xb, packed=True)
    with pytest.raises(_capi.VitdetError):
        m.submit(xa)                                      # both slots busy
    got_b = m.collect(tb)                                 # out of order
    got_a = m.collect(ta)
    _same(got_a, ra); _same(got_b, rb)
    assert got_b.packed.shape == (5 * 17, parallel.RECORD_WIDTH) and got_a.packed is None
    pinned = torch.from_numpy(xa).pin_memory()
    _same(m.collect(m.submit(pinned.numpy())), ra)
    with pytest.raises(_capi.VitdetError):
        m.collect(ta)                                     # not in flight any more
    # a full-size batch: the staged (pageable) copy runs on the library's thread pool, sub-chunk by sub-chunk
    cfg2 = vd.DetectorConfig()
    m2 = build_model(cfg2, vd.random_weights(cfg2, seed=1, spread=True), "bf16")
    x2 = images(cfg2, 9)
    t1 = m2.submit(x2[:5]); t2 = m2.submit(x2[5:])
    r1, r2 = m2.collect(t1), m2.collect(t2)
    whole = m2.detect(x2)
    assert np.array_equal(np.concatenate([r1.logits, r2.logits]), whole.logits)
    m.close(); m2.close()


def test_packed_record_is_the_other_fields_in_one_row():
    import torch
    cfg = tiny_config()
    m = build_model(cfg, vd.random_weights(cfg, seed=3, spread=True), "bf16")
    x = images(cfg, 6)
    for inp in (x, torch.from_numpy(x).cuda()):
        rec = m.detect(inp, strict=False, packed=True)
        pk = rec.packed.cpu().numpy() if hasattr(rec.packed, "cpu") else rec.packed
        un = parallel.unpack_records(pk)
        get = lambda a: a.cpu().numpy() if hasattr(a, "cpu") else a
        assert np.array_equal(un["decoded"], get(rec.decoded)) and np.array_equal(un["class_id"], get(rec.class_id))
        assert np.array_equal(un["class_conf"], get(rec.class_conf)) and np.array_equal(un["keep"], get(rec.keep))
        assert np.array_equal(un["corners"], get(rec.corners))
        assert np.array_equal(pk, parallel.pack_records(rec).cpu().numpy() if hasattr(rec.decoded, "cpu") else parallel.pack_records(rec))
    # the stateless decode fills it too
    logits = m.predict(x)
    dec = vd.decode_predictions(logits)
    assert dec.packed is None
    m.close()


def test_options_and_taps_api():
    cfg = tiny_config()
    m = build_model(cfg, vd.random_weights(cfg, seed=4, spread=True), "bf16")
    assert m.get_option("fuse_ln") == 1 and m.get_option("fuse_tail") == 1 and m.get_option("gemm_pair") == 1 and m.get_option("fp32_tc") == 1
    with pytest.raises(_capi.VitdetError) as ei:
        m.set_option("no_such_switch", 1)
    assert ei.value.code == _capi.E_NOT_FOUND
    with pytest.raises(_capi.VitdetError):
        m.set_option("gemm_pair", 7)
    with pytest.raises(_capi.VitdetError) as ei:
        m.debug_read("block_1", 1)                        # taps are off
    assert ei.value.code == _capi.E_UNSET
    x = images(cfg, 3)
    base = m.predict(x)
    m.debug_taps(True)
    assert np.array_equal(m.predict(x), base)             # taps do not change the result
    t = m.debug_read("block_2", 3)
    assert t.shape == (3, cfg.tokens, cfg.embedding_dim) and np.isfinite(t).all()
    with pytest.raises(_capi.VitdetError):
        m.debug_read("block_9", 3)
    m.close()


@pytest.mark.parametrize("cfg_kw", [dict(), dict(use_mish=False, input_shape=(70, 100, 3))])
def test_fp32_mode_on_ieee_cuda_core_kernels(cfg_kw):
    """vitdet_set_option(h, "fp32_tc", 0): every product in IEEE float32 FMA (gemm_simt.cu, attention.cu) — the strict form
    of the fp32 mode; an order of magnitude closer to float64 than the tensor-core form, and both inside 1e-3."""
    cfg = tiny_config(**cfg_kw)
    w = vd.random_weights(cfg, seed=11, spread=True)
    x = images(cfg, 4)
    ref = oracle.forward(w, cfg, x, np.float64)
    m = build_model(cfg, w, "fp32")
    e_tc = rel_err(m.predict(x), ref)
    m.set_option("fp32_tc", 0)
    e_ieee = rel_err(m.predict(x), ref)
    assert e_ieee < 2e-5 and e_tc < TOL_FP32 / 5
    m.close()
    old = os.environ.get("VITDET_FP32")
    os.environ["VITDET_FP32"] = "simt"
    try:
        from vision_transformer_detector_b200 import ops
        import torch
        rng = np.random.default_rng(0)
        a = rng.normal(size=(300, 200)).astype(np.float32); wt = (rng.normal(size=(200, 70)) / 14).astype(np.float32)
        got = ops.dense(torch.from_numpy(a).cuda(), torch.from_numpy(wt).cuda(), None, mode="fp32").cpu().numpy()
        assert rel_err(got, a.astype(np.float64) @ wt.astype(np.float64)) < 2e-6
        q, k, v = (rng.normal(size=(2, 100, 2, 40)).astype(np.float32) for _ in range(3))
        got = ops.attention(*(torch.from_numpy(t).cuda() for t in (q, k, v)), mode="fp32").cpu().numpy()
        assert rel_err(got, oracle.attention_core(q.astype(np.float64), k.astype(np.float64), v.astype(np.float64))) < 2e-6
    finally:
        if old is None:
            del os.environ["VITDET_FP32"]
        else:
            os.environ["VITDET_FP32"] = old
