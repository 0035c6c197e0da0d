"""visualize_predictions (reference det.py:2187-2456) without the GUI: the per-slot Python loop of the reference
(thresholds, class id, class confidence, corner boxes — det.py:2257-2325) runs as the device decode; this module
only draws what the records say.  The reference shows every image with cv.imshow and blocks on a key press; here
the annotated BGR images are returned (and optionally written to disk), so the path works headless.
"""
from __future__ import annotations

from collections.abc import Iterable

import numpy as np

from .vision_transformer_detector import Constants, decode_predictions

# The 80 COCO detection categories in the reference's `id_in_model` order (full_categories.csv).
COCO_CATEGORY_NAMES = (
    "person", "bicycle", "car", "motorcycle", "airplane", "bus", "train", "truck", "boat", "traffic light",
    "fire hydrant", "stop sign", "parking meter", "bench", "bird", "cat", "dog", "horse", "sheep", "cow", "elephant",
    "bear", "zebra", "giraffe", "backpack", "umbrella", "handbag", "tie", "suitcase", "frisbee", "skis", "snowboard",
    "sports ball", "kite", "baseball bat", "baseball glove", "skateboard", "surfboard", "tennis racket", "bottle",
    "wine glass", "cup", "fork", "knife", "spoon", "bowl", "banana", "apple", "sandwich", "orange", "broccoli", "carrot",
    "hot dog", "pizza", "donut", "cake", "chair", "couch", "potted plant", "bed", "dining table", "toilet", "tv",
    "laptop", "mouse", "remote", "keyboard", "cell phone", "microwave", "oven", "toaster", "sink", "refrigerator",
    "book", "clock", "vase", "scissors", "teddy bear", "hair drier", "toothbrush",
)

TEXT_HEIGHT = 20      # det.py:2220
TEXT_WIDTH = 60       # det.py:2221


def _category_name(categories_to_detect, class_id: int) -> str:
    if categories_to_detect is None:
        return COCO_CATEGORY_NAMES[class_id]
    if hasattr(categories_to_detect, "at"):                     # pandas DataFrame indexed by id_in_model (det.py:2273)
        return str(categories_to_detect.at[class_id, "name"])
    return str(categories_to_detect[class_id])


def _to_numpy(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


def _enlarged_bgr(image, scale: float) -> np.ndarray:
    """det.py:2226-2252: [-1, 1] float image -> uint8, resized by enlarged_image_scale, RGB -> BGR."""
    import cv2 as cv
    arr = (np.asarray(_to_numpy(image), np.float32) + 1.0) * 127.5          # det.py:2227-2228 (on a copy)
    arr = np.clip(np.rint(arr), 0, 255).astype(np.uint8)
    h, w = int(round(arr.shape[0] * scale)), int(round(arr.shape[1] * scale))
    if (h, w) != arr.shape[:2]:
        arr = cv.resize(arr, (w, h), interpolation=cv.INTER_CUBIC)
    return np.ascontiguousarray(arr[..., ::-1])


def _visualize_one_batch_prediction(images_batch, records, show_classification_confidence=True,
                                    categories_to_detect=None, enlarged_image_scale=1) -> list[np.ndarray]:
    """Draws the kept slots of every image (det.py:2224-2340).  `records` is a DetectionRecords whose `keep`,
    `class_id`, `class_conf` and `corners` were computed on the GPU with the visualise rule."""
    import cv2 as cv
    keep, cid, cc, cor = (_to_numpy(a) for a in (records.keep, records.class_id, records.class_conf, records.corners))
    out = []
    for b, image in enumerate(images_batch):
        image_bgr = _enlarged_bgr(image, enlarged_image_scale)
        image_width = image_bgr.shape[1]
        for s in np.nonzero(keep[b])[0]:
            name = _category_name(categories_to_detect, int(cid[b, s]))
            show_text = f"{name} {float(cc[b, s]):.0%}" if show_classification_confidence else name      # det.py:2284-2291
            x0, y0, x1, y1 = (int(v) for v in cor[b, s])
            cv.rectangle(img=image_bgr, pt1=(x0, y0), pt2=(x1, y1), color=(0, 255, 0), thickness=2)      # det.py:2327
            text_point = [x0, y0 - 6]                                                                    # det.py:2331-2337
            if y0 < TEXT_HEIGHT:
                text_point[1] = TEXT_HEIGHT
            if image_width - x0 < TEXT_WIDTH:
                text_point[0] = image_width - TEXT_WIDTH
            cv.putText(img=image_bgr, text=show_text, org=tuple(text_point), fontFace=cv.FONT_HERSHEY_TRIPLEX,
                       fontScale=0.5, color=(0, 255, 0))
        out.append(image_bgr)
    return out


def visualize_predictions(image_input, predictions=None, objectness_threshold=None, classification_threshold=None,
                          show_classification_confidence=True, categories_to_detect=None, enlarged_image_scale=1,
                          is_image=True, is_video=False, save_prefix: str | None = None) -> list[np.ndarray]:
    """Same arguments as the reference (det.py:2363-2369).  Returns the annotated BGR images instead of opening
    windows; `save_prefix` writes them as '<prefix>_<n>.png' (the reference saves on key press 's').

    image_input: an image batch (B, H, W, 3) in [-1, 1] together with `predictions` = raw model output (B, 17, 6); or,
    with predictions=None, an iterable of (images, labels) batches whose labels are already decoded rows, drawn without
    the confidence (det.py:2420-2436).  is_image / is_video are accepted for signature compatibility."""
    if objectness_threshold is None:
        objectness_threshold = Constants.OBJECTNESS_THRESHOLD.value
    if classification_threshold is None:
        classification_threshold = Constants.CLASSIFICATION_CONFIDENCE_THRESHOLD.value
    results: list[np.ndarray] = []
    if predictions is None:
        if isinstance(image_input, Iterable) and not isinstance(image_input, (str, bytes)):
            for element in image_input:
                images_batch, labels = element[0], element[1]
                size = tuple(int(v) for v in _to_numpy(images_batch).shape[1:3])
                rec = decode_predictions(labels, objectness_threshold, classification_threshold, strict=False,
                                         image_size=size, use_transform_predictions=False, corner_scale=enlarged_image_scale)
                results += _visualize_one_batch_prediction(images_batch, rec, False, categories_to_detect, enlarged_image_scale)
    else:
        # transform_predictions scales by Constants.MODEL_IMAGE_SIZE (det.py:637-640, :2447)
        rec = decode_predictions(predictions, objectness_threshold, classification_threshold, strict=False,
                                 image_size=Constants.MODEL_IMAGE_SIZE.value, corner_scale=enlarged_image_scale)
        results = _visualize_one_batch_prediction(image_input, rec, show_classification_confidence, categories_to_detect,
                                                  enlarged_image_scale)
    if save_prefix:
        import cv2 as cv
        for n, img in enumerate(results):
            cv.imwrite(f"{save_prefix}_{n}.png", img)
    return results
