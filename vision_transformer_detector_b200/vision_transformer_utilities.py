"""Host-side mirror of the part of the reference's `vision_transformer_utilities.py` that sits immediately before the
hot path: turning an image file (or a decoded uint8 array) into the model's input tensor.  The COCO dataset plumbing of
the reference module (annotation parsing at import time from hard-coded D:\\ paths, tf.data pipeline) is out of scope.

    _get_image_tensor_coco(one_image_path)        vision_transformer_utilities.py:418-449
    preprocess_image(image_uint8)                 the arithmetic of that function: resize_with_pad, clip, /127.5 - 1

The arithmetic runs on the GPU (vitdet_preprocess_image*); only the file decode is host work (PIL)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .vision_transformer_detector import Constants, _is_torch_cuda, _torch_stream_ptr

MODEL_IMAGE_HEIGHT, MODEL_IMAGE_WIDTH = Constants.MODEL_IMAGE_SIZE.value       # vision_transformer_utilities.py:23-24


def resize_with_pad_geometry(height: int, width: int, target_height: int, target_width: int) -> tuple[int, int, int, int]:
    """(resized_height, resized_width, pad_top, pad_left) of tf.image.resize_with_pad, in TF's float32 arithmetic."""
    v = [C.c_int() for _ in range(4)]
    _capi.check(_capi.load().vitdet_resize_with_pad_geometry(int(height), int(width), int(target_height), int(target_width),
                                                             *[C.byref(x) for x in v]))
    return tuple(x.value for x in v)


def preprocess_image(image, target_size=None):
    """uint8 (h, w, 3) -> float32 (H, W, 3) in [-1, 1]: resize_with_pad + clip + /127.5 - 1, on the GPU.
    numpy in -> numpy out; torch CUDA uint8 tensor in -> torch CUDA tensor out (asynchronous on torch's stream)."""
    th, tw = (MODEL_IMAGE_HEIGHT, MODEL_IMAGE_WIDTH) if target_size is None else (int(target_size[0]), int(target_size[1]))
    lib = _capi.load()
    if _is_torch_cuda(image):
        import torch
        x = image.contiguous()
        if x.dtype != torch.uint8 or x.dim() != 3 or x.shape[2] != 3:
            raise ValueError(f"image must be a uint8 tensor of shape (h, w, 3), got {x.dtype} {tuple(x.shape)}")
        out = torch.empty((th, tw, 3), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _capi.check(lib.vitdet_preprocess_image(C.c_void_p(x.data_ptr()), int(x.shape[0]), int(x.shape[1]),
                                                    C.c_void_p(out.data_ptr()), th, tw, _torch_stream_ptr(x.device)))
        return out
    x = np.ascontiguousarray(np.asarray(image))
    if x.dtype != np.uint8 or x.ndim != 3 or x.shape[2] != 3:
        raise ValueError(f"image must be a uint8 array of shape (h, w, 3), got {x.dtype} {x.shape}")
    out = np.empty((th, tw, 3), np.float32)
    _capi.check(lib.vitdet_preprocess_image_host(_capi.np_ptr(x), int(x.shape[0]), int(x.shape[1]), _capi.np_ptr(out), th, tw))
    return out


def _get_image_tensor_coco(one_image_path):
    """vision_transformer_utilities.py:418-449: (image_tensor (H, W, 3) float32 in [-1, 1], (original_height, original_width)).
    tf.io.read_file + tf.image.decode_image(channels=3) become a PIL decode; everything after it runs on the GPU."""
    from PIL import Image
    with Image.open(one_image_path) as im:
        arr = np.asarray(im.convert("RGB"), dtype=np.uint8)
    return preprocess_image(arr), (int(arr.shape[0]), int(arr.shape[1]))
