"""Mirror of the reference's utils/image_utils.py for the accelerated path: same function name and return value,
the resize / ToTensor / Normalize run on the GPU (csrc/preprocess.cu), bit-identical to Pillow + torchvision.

``preprocess_image`` / ``preprocess_images`` decode on the host with PIL, exactly as the reference does (bit-identical
tensors); ``preprocess_images_nvjpeg`` decodes JPEG files on the GPU with nvJPEG as well (faster, pixels within a few grey
levels of libjpeg's)."""
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


_default_engine = None


def _engine_or_default(engine):
    """The preprocessing kernels need no model weights: without an explicit engine a bare handle on the current CUDA
    device is created once and reused (so the reference's two-argument call keeps working)."""
    global _default_engine
    if engine is not None:
        return engine
    if _default_engine is None or _default_engine.device.index != torch.cuda.current_device():
        from .config import swin_tiny_test
        from .engine import Engine
        _default_engine = Engine(swin_tiny_test(), torch.cuda.current_device())
    return _default_engine


def preprocess_image(image_path, img_size: int, engine=None) -> torch.Tensor:
    """reference utils/image_utils.py:5-23, same two positional arguments -> (1, 3, S, S) float32 on the CUDA device
    (the engine's, or the current device when no engine is given)."""
    return _engine_or_default(engine).preprocess_rgb8([_decode_rgb8(image_path)], img_size)


def preprocess_images_nvjpeg(image_paths: Sequence, img_size: int, engine=None) -> torch.Tensor:
    """The same for JPEG files with the DECODE on the GPU as well (nvJPEG): the files are read as bytes, nothing is decoded on
    the host.  Pixels differ from PIL's libjpeg decode by a few grey levels (different IDCT / chroma upsampling), so this is
    the fast path, not the bit-exact one; files nvJPEG cannot parse (PNG, ...) and hosts without libnvjpeg raise."""
    streams = []
    for p in image_paths:
        with open(p, "rb") as f:
            streams.append(f.read())
    return _engine_or_default(engine).preprocess_jpeg(streams, img_size)


def preprocess_images(image_paths: Sequence, img_size: int, engine=None) -> torch.Tensor:
    """Batch form of demo.py's loop (demo.py:107-112): (B, 3, S, S) float32 on the device, one launch pair for the
    whole list (xn_preprocess_rgb8_batch)."""
    return _engine_or_default(engine).preprocess_rgb8([_decode_rgb8(p) for p in image_paths], img_size)
