"""CPU ORACLE for the evaluation metric — test infrastructure, not product code.

Restates `MeanAveragePrecision` of the reference (vision_transformer_detector.py = "det.py", lines 1268-2060)
in numpy float32, statement by statement: the three state tensors are kept in the reference's own layout and
shifted physically exactly as the reference's `assign` calls do, so the state can be compared entry by entry
with the device implementation.  Only tests/, __graft_entry__.smoke() and bench.py may import this module.

PARITY PINNING: pinned.  tests/golden/map_vectors.json transcribes the 12 known-answer cases + the reset case
of the reference's own TestMeanAveragePrecision (testcases_vision_transformer_detector.py:49-734) — inputs and
the asserted AP — and tests/test_oracle.py checks this restatement reproduces every asserted value exactly
(the reference asserts with assertEqual on the float32 result).

TensorFlow semantics relied on (TF 2.9):
  * tf.round                      — round half to even                      -> np.rint
  * tf.experimental.numpy.isclose — |a-b| <= atol + rtol*|b|, rtol=1e-5, atol=1e-8, in the promoted dtype
  * tf.argsort / tf.sort          — built on top_k, which returns the LOWER index first among equal values
                                    (ascending sorts negate the values first, so the same tie rule holds)
  * tf.linspace(0.5, 0.95, 10)    — float32: first = start, last = stop, middle = start + delta*i with
                                    delta = (stop-start)/(num-1)
"""
from __future__ import annotations

import numpy as np

try:
    from .vitdet_oracle import (CLASSES, CLASSIFICATION_CONFIDENCE_THRESHOLD, MODEL_IMAGE_SIZE, OBJECTNESS_THRESHOLD,
                                iou_calculator, transform_predictions)
except ImportError:                      # tests put oracle/ itself on sys.path
    from vitdet_oracle import (CLASSES, CLASSIFICATION_CONFIDENCE_THRESHOLD, MODEL_IMAGE_SIZE, OBJECTNESS_THRESHOLD,
                               iou_calculator, transform_predictions)

LATEST_RELATED_IMAGES = 3          # Constants.LATEST_RELATED_IMAGES  det.py:32
BBOXES_PER_IMAGE = 14              # Constants.BBOXES_PER_IMAGE       det.py:37

F = np.float32


def iou_thresholds() -> np.ndarray:
    """tf.linspace(0.5, 0.95, num=10) in float32 (det.py:1876)."""
    start, stop = F(0.5), F(0.95)
    delta = F(F(stop - start) / F(9))
    mid = [F(start + F(delta * F(i))) for i in range(1, 9)]
    return np.array([start, *mid, stop], dtype=F)


def _isclose(a, b):
    a64, b64 = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a64 - b64) <= 1e-8 + 1e-5 * np.abs(b64)


def _isclose_f32(a, b):
    a, b = np.asarray(a, dtype=F), np.asarray(b, dtype=F)
    return np.abs(a - b) <= F(F(1e-8) + F(F(1e-5) * np.abs(b)))


def class_confidence(classification: np.ndarray) -> np.ndarray:
    """(0.5 - |c - round(c)|) / 0.5  (det.py:1366-1376)."""
    classification = np.asarray(classification, dtype=F)
    return ((F(0.5) - np.abs(classification - np.rint(classification))) / F(0.5)).astype(F)


def _argsort_stable(values: np.ndarray, descending: bool) -> np.ndarray:
    values = np.asarray(values)
    return np.argsort(-values if descending else values, kind="stable")


class MeanAveragePrecision:
    """det.py:1268-2060.  State layout = the reference's tf.Variables (det.py:1286-1305)."""

    def __init__(self, classes: int = CLASSES, latest_related_images: int = LATEST_RELATED_IMAGES,
                 bboxes_per_image: int = BBOXES_PER_IMAGE, image_size=MODEL_IMAGE_SIZE,
                 objectness_threshold: float = OBJECTNESS_THRESHOLD,
                 classification_threshold: float = CLASSIFICATION_CONFIDENCE_THRESHOLD):
        self.classes, self.L, self.K = int(classes), int(latest_related_images), int(bboxes_per_image)
        self.image_size = tuple(image_size)
        self.obj_thr, self.cls_thr = F(objectness_threshold), F(classification_threshold)
        self.reset_state()

    def reset_state(self):                                                              # det.py:2052-2060
        self.latest_positive_bboxes = np.zeros((self.classes, self.L, self.K, 2), dtype=F)
        self.labels_quantity_per_image = np.zeros((self.classes, self.L), dtype=F)
        self.showed_up_classes = np.zeros((self.classes,), dtype=bool)

    # ---------------------------------------------------------------------------------------------
    def update_state(self, y_true, y_pred, sample_weight=None, use_transform_predictions=True):
        y_true = np.asarray(y_true, dtype=F)
        y_pred = np.asarray(y_pred, dtype=F)
        if use_transform_predictions:                                                   # det.py:1341-1342
            y_pred = transform_predictions(y_pred, self.image_size, self.classes).astype(F)
        K = self.K

        # showed_up_classes (det.py:1346-1420)
        categories_label = y_true[..., 1]
        shown_label = categories_label[categories_label >= 0].astype(np.int32)
        conf_all = class_confidence(y_pred[..., 1])
        positive_all = (y_pred[..., 0] > self.obj_thr) & (conf_all > self.cls_thr)
        shown_pred = np.rint(y_pred[..., 1])[positive_all].astype(np.int32)
        for c in np.union1d(shown_pred, shown_label):
            if 0 <= c < self.classes:        # the reference would raise on an out-of-range index; ignored here
                self.showed_up_classes[c] = True

        for sample in range(y_true.shape[0]):                                           # det.py:1427
            one_label, one_pred = y_true[sample], y_pred[sample]
            categories_one_label = one_label[..., 1]
            categories_one_pred = np.rint(one_pred[..., 1])
            conf_one = class_confidence(one_pred[..., 1])
            positives_index = (one_pred[..., 0] > self.obj_thr) & (conf_one > self.cls_thr)       # det.py:1461-1464
            positives_one_pred = np.where(positives_index[:, None], one_pred, F(-8)).astype(F)   # det.py:1468-1470
            positives_category = np.where(positives_index, categories_one_pred, F(-8)).astype(F)  # det.py:1474-1476

            for category in range(self.classes):                                        # det.py:1480
                category_bool_label = _isclose(categories_one_label, category)
                category_bool_pred = _isclose(positives_category, category)
                any_label, any_pred = bool(category_bool_label.any()), bool(category_bool_pred.any())
                if not (any_label or any_pred):                                         # scenario a
                    continue

                # labels_quantity_per_image (det.py:1529-1543)
                self.labels_quantity_per_image[category, 1:] = self.labels_quantity_per_image[category, :-1].copy()
                self.labels_quantity_per_image[category, 0] = F(np.count_nonzero(category_bool_label))

                if any_label and not any_pred:                                          # scenario b, det.py:1551-1555
                    one_image = np.zeros((K, 2), dtype=F)

                elif any_pred and not any_label:                                        # scenario c, det.py:1559-1616
                    conf = class_confidence(positives_one_pred[category_bool_pred][:, 1])
                    if conf.shape[0] < K:
                        conf = np.concatenate([conf, np.zeros(K - conf.shape[0], dtype=F)])
                    else:
                        conf = conf[_argsort_stable(conf, descending=True)][:K]
                    one_image = np.stack([conf, np.zeros_like(conf)], axis=1)

                else:                                                                   # scenario d, det.py:1620-1839
                    bboxes_iou_pred = np.where(category_bool_pred[:, None], positives_one_pred[:, -4:], F(-8)).astype(F)
                    bboxes_category_label = one_label[:, -4:][category_bool_label]
                    area = (bboxes_category_label[:, -1] * bboxes_category_label[:, -2]).astype(F)
                    sorted_bboxes_label = bboxes_category_label[_argsort_stable(area, descending=False)]
                    one_image = np.zeros((K, 2), dtype=F)
                    new_bboxes_quantity = 0
                    for bbox_info in sorted_bboxes_label:                               # det.py:1661
                        bbox_iou_label = np.ones_like(bboxes_iou_pred) * bbox_info
                        ious = iou_calculator(bbox_iou_label, bboxes_iou_pred).astype(F)
                        max_iou = F(ious.max())
                        if max_iou > F(0.5):                                            # det.py:1686
                            new_bboxes_quantity += 1
                            position = _isclose_f32(ious, max_iou)
                            best = positives_one_pred[position]
                            new_bbox = np.array([[class_confidence(best[0, 1]), max_iou]], dtype=F)
                            one_image = np.concatenate([one_image, new_bbox], axis=0)[-K:]
                            bboxes_iou_pred = np.where(position[:, None], F(-8), bboxes_iou_pred).astype(F)
                        if new_bboxes_quantity == K:                                    # det.py:1754-1756
                            break
                    left_bool = np.all(bboxes_iou_pred >= 0, axis=-1)                   # det.py:1765-1766
                    left_pred = positives_one_pred[left_bool]
                    left_quantity = left_pred.shape[0]
                    if left_quantity > 0 and new_bboxes_quantity < K:                   # det.py:1783-1786
                        left_conf = class_confidence(left_pred[:, 1])
                        if new_bboxes_quantity + left_quantity > K:                     # det.py:1807-1824
                            left_conf = left_conf[_argsort_stable(left_conf, descending=True)][:K - new_bboxes_quantity]
                        left = np.stack([left_conf, np.zeros_like(left_conf)], axis=1)
                        one_image = np.concatenate([one_image, left], axis=0)[-K:]

                self.latest_positive_bboxes[category, 1:] = self.latest_positive_bboxes[category, :-1].copy()  # det.py:1852
                self.latest_positive_bboxes[category, 0] = one_image                                           # det.py:1857

    # ---------------------------------------------------------------------------------------------
    def average_precisions(self) -> np.ndarray:
        """(10, classes) AP per IoU threshold and class; classes that never showed up hold 0 (det.py:1876-2022)."""
        out = np.zeros((10, self.classes), dtype=F)
        for ti, iou_threshold in enumerate(iou_thresholds()):
            for category in range(self.classes):
                if not self.showed_up_classes[category]:
                    continue
                recall_precisions = [F(1)]
                true_positives, false_positives = F(0), F(0)
                boxes = self.latest_positive_bboxes[category].reshape(-1, 2)
                boxes = boxes[_argsort_stable(boxes[:, 0], descending=True)]            # det.py:1907-1915
                for conf, iou in boxes:                                                 # det.py:1920-1951
                    if conf > 0:
                        if iou > iou_threshold:
                            true_positives = F(true_positives + F(1))
                            precision = F(true_positives / F(true_positives + false_positives))
                            recall_precisions.append(precision)
                        else:
                            false_positives = F(false_positives + F(1))
                            precision = F(true_positives / F(true_positives + false_positives))
                            recall_precisions[-1] = precision
                labels_quantity = F(self.labels_quantity_per_image[category].sum(dtype=F))
                area = F(0)
                if labels_quantity > 0:                                                 # det.py:1962-1998
                    height = F(F(1) / labels_quantity)
                    recalls = len(recall_precisions) - 1
                    if recalls > 0:
                        edges = F(0)
                        for i in range(recalls):
                            edges = F(edges + F(recall_precisions[i] + recall_precisions[i + 1]))
                        area = F(F(edges * height) / F(2))
                out[ti, category] = area
        return out

    def average_precision_per_iou(self) -> np.ndarray:
        aps = self.average_precisions()
        shown = self.showed_up_classes
        per_iou = np.zeros((10,), dtype=F)
        n = int(shown.sum())
        if n:                                                                           # det.py:2031-2040
            for ti in range(10):
                s = F(0)
                for v in aps[ti][shown]:
                    s = F(s + v)
                per_iou[ti] = F(s / F(n))
        return per_iou

    def result(self) -> np.float32:                                                     # det.py:1865-2049
        per_iou = self.average_precision_per_iou()
        s = F(0)
        for v in per_iou:
            s = F(s + v)
        return F(s / F(10))
