"""vision_transformer_detector_b200 — B200-native (sm_100a) forward pass + anchor-free head decode of
the ViT detector of westlake-moonlight/vision_transformer_detector, behind the reference's own Python
entry points.  See DESIGN.md / INTEGRATION.md at the repository root."""
from .vision_transformer_detector import (  # noqa: F401
    Constants,
    DetectionRecords,
    DetectorConfig,
    MeanAveragePrecision,
    VisionTransformerDetector,
    create_vision_transformer_detector,
    decode_predictions,
    iou_calculator,
    mlp_head,
    random_weights,
    transform_predictions,
    transformer_encoder,
    transformer_preprocessor,
    weight_specs,
)

from .visualization import COCO_CATEGORY_NAMES, visualize_predictions  # noqa: F401,E402

from .vision_transformer_utilities import preprocess_image, resize_with_pad_geometry  # noqa: F401,E402

__all__ = [
    "COCO_CATEGORY_NAMES", "visualize_predictions", "preprocess_image", "resize_with_pad_geometry",
    "Constants", "DetectionRecords", "DetectorConfig", "MeanAveragePrecision", "VisionTransformerDetector",
    "create_vision_transformer_detector", "decode_predictions", "iou_calculator", "mlp_head", "random_weights",
    "transform_predictions", "transformer_encoder", "transformer_preprocessor", "weight_specs",
]
