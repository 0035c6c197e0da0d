"""Writes tests/golden/map_vectors.json.

Each case transcribes one known-answer test of the reference's TestMeanAveragePrecision
(/root/reference/testcases_vision_transformer_detector.py, lines cited per case): the label tensor, the
prediction tensor (both (batch, 10, 6), already decoded — the tests call update_state with
use_transform_predictions=False) and the AP the test asserts with assertEqual on the float32 result.
TensorFlow cannot run in this image, so the expected values are the reference's own asserted constants,
not outputs regenerated here; this script only rebuilds the input tensors from the edits the tests make.

Tensor recipe shared by all cases (tests.py:59-62): every entry -8, then [..., 0] = 0.
Edit = [image, slot, first_field, values]; image = slot = -1 means "every slot of every image"
(the tests' `prediction[..., -4:] = ...`).
"""
import json
import os

import numpy as np

BOX79 = [1.0, 79.0, 10.2, 10.2, 10.0, 10.0]       # tests.py:66-73
BOX78 = [1.0, 78.0, 10.2, 10.2, 10.0, 10.0]       # tests.py:667-677

CASES = [
    dict(name="test_1_one_image_one_category", source="tests.py:49-89", batch=1, expected=1.0,
         label=[[0, 1, 0, BOX79]], pred=[]),
    dict(name="test_2_one_image_two_categories", source="tests.py:91-142", batch=1, expected=1.0,
         label=[[0, 1, 0, BOX79], [0, 2, 0, [1.0, 78.0, 9.5, 9.5, 5.0, 5.0]]], pred=[]),
    dict(name="test_3_one_image_low_iou", source="tests.py:144-195", batch=1, expected=0.3,
         label=[[0, 1, 0, BOX79]], pred=[[-1, -1, 2, [9.5, 9.5, 8.0, 8.0]]]),
    dict(name="test_4_one_image_zero_ap", source="tests.py:197-248", batch=1, expected=0.0,
         label=[[0, 1, 0, BOX79]], pred=[[-1, -1, 2, [9.5, 9.5, 7.0, 7.0]]]),
    dict(name="test_5_1_one_image_low_objectness", source="tests.py:250-303", batch=1, expected=0.0,
         label=[[0, 1, 0, BOX79]], pred=[[0, 1, 0, [0.49]]]),
    dict(name="test_5_2_one_image_two_predictions_one_low_objectness", source="tests.py:305-370", batch=1, expected=0.75,
         label=[[0, 1, 0, BOX79]], pred=[[0, 2, 0, [0.51, 79.0, 10.2, 10.2, 9.9, 9.9]]]),
    dict(name="test_6_one_image_low_classification_confidence", source="tests.py:372-426", batch=1, expected=0.0,
         label=[[0, 1, 0, BOX79]], pred=[[0, 1, 1, [79.255]]]),
    dict(name="test_7_two_images_one_category", source="tests.py:428-471", batch=2, expected=1.0,
         label=[[0, 1, 0, BOX79], [1, 5, 0, BOX79]], pred=[]),
    dict(name="test_8_two_images_one_zero_ap", source="tests.py:473-530", batch=2, expected=0.375,
         label=[[0, 1, 0, BOX79], [1, 0, 0, BOX79]], pred=[[1, 0, 1, [79.001]], [1, 0, 2, [9.5, 9.5, 7.0, 7.0]]]),
    dict(name="test_9_one_objectness_below_threshold", source="tests.py:532-585", batch=2, expected=0.5,
         label=[[0, 1, 0, BOX79], [1, 0, 0, BOX79]], pred=[[1, 0, 0, [0.49]]]),
    dict(name="test_10_classification_confidence_below_threshold", source="tests.py:587-641", batch=2, expected=0.5,
         label=[[0, 1, 0, BOX79], [1, 0, 0, BOX79]], pred=[[1, 0, 1, [79.3]]]),
    dict(name="test_11_two_categories_two_images", source="tests.py:643-710", batch=2, expected=0.6875,
         label=[[0, 1, 0, BOX79], [0, 2, 0, BOX78], [1, 1, 0, BOX79], [1, 2, 0, BOX78]],
         pred=[[0, 1, 1, [79.005]], [0, 1, 2, [9.5, 9.5, 7.0, 7.0]]]),
]

RESET = dict(name="test_12_reset_metric", source="tests.py:713-734",
             note="after reset_state(): latest_positive_bboxes and labels_quantity_per_image all zero, "
                  "showed_up_classes not all true, result() == 0")


def build(case, slots=10):
    """(label, prediction) float32 arrays of one case."""
    def apply(t, edits):
        for image, slot, first, values in edits:
            if image < 0:
                t[..., first:first + len(values)] = values
            else:
                t[image, slot, first:first + len(values)] = values
        return t
    base = np.full((case["batch"], slots, 6), -8.0, dtype=np.float32)
    base[..., 0] = 0
    label = apply(base.copy(), case["label"])
    pred = apply(label.copy(), case["pred"])          # the tests start every prediction from label.copy()
    return label, pred


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "map_vectors.json")
    out = {"slots": 10, "latest_related_images": 3, "bboxes_per_image": 14, "classes": 80, "cases": [], "reset": RESET}
    for c in CASES:
        label, pred = build(c)
        out["cases"].append({**c, "y_true": label.tolist(), "y_pred": pred.tolist()})
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(path)
