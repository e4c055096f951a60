"""Mirror of the reference's utils/image_utils.py for the accelerated path: same function name and return value,
the resize / ToTensor / Normalize run on the GPU (csrc/preprocess.cu), bit-identical to Pillow + torchvision.

JPEG/PNG decoding stays on the host (PIL), exactly as in the reference; nothing else does."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch


def _decode_rgb8(image_path) -> np.ndarray:
    from PIL import Image as PIL_Image
    pil_image = PIL_Image.open(image_path)
    if pil_image.mode != "RGB":
        # reference utils/image_utils.py:18-19: a non-RGB file is replaced by a blank RGB canvas of the same size
        pil_image = PIL_Image.new("RGB", pil_image.size)
    return np.asarray(pil_image, dtype=np.uint8)


def preprocess_image(image_path, img_size: int, engine) -> torch.Tensor:
    """reference utils/image_utils.py:5-23 -> (1, 3, S, S) float32, on the engine's device."""
    return engine.preprocess_rgb8([_decode_rgb8(image_path)], img_size)


def preprocess_images(image_paths: Sequence, img_size: int, engine) -> torch.Tensor:
    """Batch form used by demo.py's loop (demo.py:107-112): (B, 3, S, S) float32 on the device."""
    return engine.preprocess_rgb8([_decode_rgb8(p) for p in image_paths], img_size)
