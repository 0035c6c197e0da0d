"""CPU tests of the evaluation-metric oracle (oracle/map_oracle.py): pinned on the reference's own known answers
(tests/golden/map_vectors.json <- TestMeanAveragePrecision, tests.py:49-734), plus structural properties."""
import json
import os

import numpy as np
import pytest

from _util import map_case
import map_oracle

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "map_vectors.json")))


@pytest.mark.parametrize("case", GOLDEN["cases"], ids=[c["name"] for c in GOLDEN["cases"]])
def test_oracle_reproduces_the_reference_known_answers(case):
    m = map_oracle.MeanAveragePrecision()
    assert (m.classes, m.L, m.K) == (GOLDEN["classes"], GOLDEN["latest_related_images"], GOLDEN["bboxes_per_image"])
    m.update_state(np.array(case["y_true"], np.float32), np.array(case["y_pred"], np.float32), use_transform_predictions=False)
    # the reference asserts with assertEqual on the float32 tensor: exact equality with float32(expected)
    assert m.result() == np.float32(case["expected"])


def test_oracle_reset_state():
    """tests.py:713-734."""
    m = map_oracle.MeanAveragePrecision()
    c = GOLDEN["cases"][10]
    m.update_state(np.array(c["y_true"], np.float32), np.array(c["y_pred"], np.float32), use_transform_predictions=False)
    assert m.result() > 0
    m.reset_state()
    assert not m.latest_positive_bboxes.any() and not m.labels_quantity_per_image.any()
    assert not m.showed_up_classes.all()
    assert m.result() == 0


def test_iou_thresholds_are_float32_linspace():
    t = map_oracle.iou_thresholds()
    assert t.dtype == np.float32 and t.shape == (10,)
    assert t[0] == np.float32(0.5) and t[-1] == np.float32(0.95)
    assert np.all(np.diff(t) > 0) and np.allclose(t, np.linspace(0.5, 0.95, 10), atol=1e-6)


def test_batch_update_equals_image_by_image_updates():
    """update_state walks the batch in order and carries no cross-image state besides the rings."""
    y_true, y_pred = map_case(5, 12, 17)
    a, b = map_oracle.MeanAveragePrecision(), map_oracle.MeanAveragePrecision()
    a.update_state(y_true, y_pred, use_transform_predictions=False)
    for i in range(y_true.shape[0]):
        b.update_state(y_true[i:i + 1], y_pred[i:i + 1], use_transform_predictions=False)
    assert np.array_equal(a.latest_positive_bboxes, b.latest_positive_bboxes)
    assert np.array_equal(a.labels_quantity_per_image, b.labels_quantity_per_image)
    assert np.array_equal(a.showed_up_classes, b.showed_up_classes)
    assert a.result() == b.result()


def test_perfect_predictions_give_ap_one_and_no_predictions_give_zero():
    y_true, _ = map_case(9, 6, 17, exact_class=1.0)
    # distinct boxes per image so that every label is matched by its own copy
    m = map_oracle.MeanAveragePrecision()
    m.update_state(y_true, y_true, use_transform_predictions=False)
    assert m.result() == 1
    m.reset_state()
    empty = np.full_like(y_true, -8)
    empty[..., 0] = 0
    m.update_state(y_true, empty, use_transform_predictions=False)
    assert m.result() == 0 and m.showed_up_classes.any()


def test_ring_keeps_only_the_latest_related_images():
    y_true, y_pred = map_case(11, 9, 17, classes_used=(4,), max_labels=2)
    m = map_oracle.MeanAveragePrecision(latest_related_images=2, bboxes_per_image=3)
    m.update_state(y_true, y_pred, use_transform_predictions=False)
    tail = map_oracle.MeanAveragePrecision(latest_related_images=2, bboxes_per_image=3)
    related = [i for i in range(9) if (y_true[i, :, 1] == 4).any() or
               ((y_pred[i, :, 0] > 0.5) & (np.rint(y_pred[i, :, 1]) == 4) & (map_oracle.class_confidence(y_pred[i, :, 1]) > 0.5)).any()]
    assert len(related) > 2
    keep = related[-2:]
    tail.update_state(y_true[keep], y_pred[keep], use_transform_predictions=False)
    assert np.array_equal(m.latest_positive_bboxes[4], tail.latest_positive_bboxes[4])
    assert np.array_equal(m.labels_quantity_per_image[4], tail.labels_quantity_per_image[4])
