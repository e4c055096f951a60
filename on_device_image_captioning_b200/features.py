"""Bulk Swin feature extraction in the reference's storage layout (reference data_generator.py:96-160):
one dataset per image, key "<img_id>_features", value float32 (144, 1536) = SwinTransformer.forward_features of the
preprocessed image.  The reference walks the images one by one on the CPU/GPU with batch 1; here images are decoded on
the host, resized / normalised on the GPU (csrc/preprocess.cu, bit-identical to Pillow + torchvision) and pushed through
the Swin backbone in batches.  Data-parallel use: give every rank its slice of the image list (dist.shard_indices) and
its own output file; the reader (coco_dataloader.py:446,517) only needs the keys.

Container: HDF5 when h5py is importable (the reference's format), otherwise an .npz archive with the same keys."""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Sequence, Union

import numpy as np
import torch


def feature_key(img_id) -> str:
    return str(img_id) + "_features"          # data_generator.py:114-116


class FeatureWriter:
    """Append-only store of "<img_id>_features" -> (L, C) float32 arrays."""

    def __init__(self, path: str):
        self.path = path
        try:
            import h5py  # noqa: F401
            self._h5 = __import__("h5py").File(path, "w")
            self._mem = None
        except ImportError:
            self._h5 = None
            self._mem: Dict[str, np.ndarray] = {}

    @property
    def container(self) -> str:
        return "hdf5" if self._h5 is not None else "npz"

    def add(self, img_id, feats: np.ndarray):
        arr = np.ascontiguousarray(feats, dtype=np.float32)
        if self._h5 is not None:
            self._h5.create_dataset(feature_key(img_id), data=arr)
        else:
            self._mem[feature_key(img_id)] = arr

    def close(self):
        if self._h5 is not None:
            self._h5.close()
        else:
            with open(self.path, "wb") as f:       # exact path (np.savez would append ".npz" to a bare name)
                np.savez(f, **self._mem)


def read_features(path: str, img_id) -> np.ndarray:
    try:
        import h5py
        with h5py.File(path, "r") as f:
            return f[feature_key(img_id)][()]
    except ImportError:
        with np.load(path) as z:
            return z[feature_key(img_id)]


def extract_features(engine, images: Sequence[Union[str, np.ndarray, torch.Tensor]], img_ids: Sequence, out_path: str,
                     batch_size: int = 64, img_size: int = None) -> str:
    """images: file paths (decoded with PIL, non-RGB files become a blank canvas as in the reference) or decoded
    (H, W, 3) uint8 arrays.  Returns the container kind written ("hdf5" / "npz")."""
    from .image_utils import _decode_rgb8
    assert len(images) == len(img_ids)
    w = FeatureWriter(out_path)
    for b0 in range(0, len(images), batch_size):
        chunk = [(_decode_rgb8(im) if isinstance(im, (str, os.PathLike)) else im) for im in images[b0:b0 + batch_size]]
        x = engine.preprocess_rgb8(chunk, img_size)
        f = engine.forward_swin(x).cpu().numpy()                       # (B, 144, 1536)
        for i, img_id in enumerate(img_ids[b0:b0 + batch_size]):
            w.add(img_id, f[i])
    w.close()
    return w.container
