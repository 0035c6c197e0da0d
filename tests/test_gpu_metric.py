"""GPU parity tests of the evaluation metric (csrc/metric.cu behind MeanAveragePrecision) against the pinned oracle
(oracle/map_oracle.py) and the reference's own known answers.  Integer / index work and float32 arithmetic in the
reference's order: everything here is compared for exact equality."""
import json
import os

import numpy as np
import pytest

from _util import build_model, images, map_case, tiny_config
import map_oracle
from vision_transformer_detector_b200 import MeanAveragePrecision, decode_predictions, random_weights

pytestmark = pytest.mark.gpu

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "map_vectors.json")))


def assert_same_state(dev, ref):
    assert np.array_equal(dev.showed_up_classes, ref.showed_up_classes)
    assert np.array_equal(dev.labels_quantity_per_image, ref.labels_quantity_per_image)
    assert np.array_equal(dev.latest_positive_bboxes, ref.latest_positive_bboxes)
    assert np.array_equal(dev.average_precisions(), ref.average_precisions())
    assert np.array_equal(dev.average_precision_per_iou(), ref.average_precision_per_iou())
    assert dev.result() == ref.result()


@pytest.mark.parametrize("case", GOLDEN["cases"], ids=[c["name"] for c in GOLDEN["cases"]])
def test_reference_known_answers(case):
    """TestMeanAveragePrecision, tests.py:49-710: same inputs, same assertEqual on the float32 result."""
    m = MeanAveragePrecision()
    m.reset_state()
    m.update_state(np.array(case["y_true"], np.float32), np.array(case["y_pred"], np.float32), use_transform_predictions=False)
    assert m.result() == np.float32(case["expected"])
    assert m.launch_count() >= 4


def test_reset_metric():
    """tests.py:713-734."""
    m = MeanAveragePrecision()
    c = GOLDEN["cases"][10]
    m.update_state(np.array(c["y_true"], np.float32), np.array(c["y_pred"], np.float32), use_transform_predictions=False)
    assert m.result() > 0
    m.reset_state()
    assert np.all(np.isclose(m.latest_positive_bboxes, 0))
    assert np.all(np.isclose(m.labels_quantity_per_image, 0))
    assert not np.all(m.showed_up_classes)
    assert np.isclose(m.result(), 0)


def test_iou_thresholds_match_float32_linspace():
    assert np.array_equal(MeanAveragePrecision().iou_thresholds, map_oracle.iou_thresholds())


@pytest.mark.parametrize("seed,batch,slots,L,K,classes_used", [
    (1, 1, 10, 3, 14, (79,)),
    (2, 8, 17, 3, 14, (3, 17, 79)),
    (3, 16, 17, 3, 2, (5,)),                 # more matches / leftovers than K: the truncation and top-K paths
    (4, 16, 17, 2, 3, (0, 1)),               # class 0, ring overflow inside one batch
    (5, 64, 17, 3, 14, tuple(range(80))),    # every class
    (6, 5, 40, 4, 5, (7, 8)),                # more slots than a warp has lanes
    (7, 33, 17, 41, 15, (2, 30, 31, 60)),    # the reference's commented-out constants (det.py:32, 37)
    (8, 3, 1, 1, 1, (9,)),                   # smallest shapes
])
def test_random_batches_state_and_ap_are_bit_exact(seed, batch, slots, L, K, classes_used):
    ref = map_oracle.MeanAveragePrecision(latest_related_images=L, bboxes_per_image=K)
    dev = MeanAveragePrecision(latest_related_images=L, bboxes_per_image=K)
    for step in range(3):                     # state carries over between calls
        y_true, y_pred = map_case(100 * seed + step, batch, slots, classes_used=classes_used,
                                  max_labels=min(8, slots), max_extra=min(8, slots))
        ref.update_state(y_true, y_pred, use_transform_predictions=False)
        dev.update_state(y_true, y_pred, use_transform_predictions=False)
        assert_same_state(dev, ref)


def test_device_tensors_and_batch_splitting():
    """torch CUDA tensors stay on the device; one call with the whole batch == one call per image."""
    import torch
    y_true, y_pred = map_case(21, 24, 17)
    ref = map_oracle.MeanAveragePrecision()
    ref.update_state(y_true, y_pred, use_transform_predictions=False)
    whole, split = MeanAveragePrecision(), MeanAveragePrecision()
    t, p = torch.from_numpy(y_true).cuda(), torch.from_numpy(y_pred).cuda()
    whole.update_state(t, p, use_transform_predictions=False)
    for i in range(24):
        split.update_state(t[i:i + 1], p[i:i + 1], use_transform_predictions=False)
    assert_same_state(whole, ref)
    assert_same_state(split, ref)


def test_empty_and_degenerate_inputs():
    dev, ref = MeanAveragePrecision(), map_oracle.MeanAveragePrecision()
    empty = np.full((4, 17, 6), -8, np.float32)
    empty[..., 0] = 0
    dev.update_state(empty, empty, use_transform_predictions=False)          # scenario a everywhere
    assert dev.result() == 0 and not dev.showed_up_classes.any()
    dev.update_state(np.zeros((0, 17, 6), np.float32), np.zeros((0, 17, 6), np.float32), use_transform_predictions=False)
    # labels only (scenario b), predictions only (scenario c), zero-area boxes, identical boxes (isclose removes all ties)
    y_true, y_pred = map_case(31, 6, 17, classes_used=(11,))
    only_labels = y_pred.copy(); only_labels[..., 0] = 0
    no_labels = empty[:1].repeat(6, 0)
    twin = y_true.copy()
    twin[:, 1] = twin[:, 0]
    zero_area = y_true.copy(); zero_area[..., 4:] = np.where(zero_area[..., 4:] > 0, 0, zero_area[..., 4:])
    for yt, yp in [(y_true, only_labels), (no_labels, y_pred), (twin, twin), (zero_area, zero_area), (y_true, y_pred)]:
        dev.update_state(yt, yp, use_transform_predictions=False)
        ref.update_state(yt, yp, use_transform_predictions=False)
        assert_same_state(dev, ref)
    with pytest.raises(ValueError):
        dev.update_state(np.zeros((2, 17, 5), np.float32), np.zeros((2, 17, 5), np.float32))


def test_raw_head_outputs_go_through_transform_predictions():
    """use_transform_predictions=True (det.py:1341-1342): the metric decodes the raw logits itself, with the same device
    arithmetic as decode_predictions, so feeding the oracle the product's decoded rows must give the same state."""
    rng = np.random.default_rng(3)
    logits = rng.normal(0, 2.0, (16, 17, 6)).astype(np.float32)
    y_true, _ = map_case(41, 16, 17, classes_used=tuple(range(0, 80, 7)))
    decoded = decode_predictions(logits).decoded
    ref = map_oracle.MeanAveragePrecision()
    ref.update_state(y_true, decoded, use_transform_predictions=False)
    dev = MeanAveragePrecision()
    dev.update_state(y_true, logits)
    assert_same_state(dev, ref)
    assert ref.showed_up_classes.sum() > 5


def test_model_output_feeds_the_metric_on_device():
    """model(images) -> update_state without leaving the GPU; same state as the host round trip."""
    import torch
    cfg = tiny_config()
    model = build_model(cfg, random_weights(cfg, seed=5, spread=True), "fp32")
    x = images(cfg, 6)
    y_true, _ = map_case(51, 6, 17)
    logits_dev = model(torch.from_numpy(x).cuda())
    a = MeanAveragePrecision(image_size=cfg.input_shape[:2])
    a.update_state(torch.from_numpy(y_true).cuda(), logits_dev)
    b = MeanAveragePrecision(image_size=cfg.input_shape[:2])
    b.update_state(y_true, logits_dev.cpu().numpy())
    assert np.array_equal(a.latest_positive_bboxes, b.latest_positive_bboxes)
    assert np.array_equal(a.showed_up_classes, b.showed_up_classes)
    assert a.result() == b.result()
