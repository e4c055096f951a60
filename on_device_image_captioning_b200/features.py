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


def shard_path(out_path: str, rank: int, world: int) -> str:
    """Per-rank container name: "<base>.rank<r>of<W><ext>" (the plain path for a single process)."""
    if world <= 1:
        return out_path
    base, ext = os.path.splitext(out_path)
    return f"{base}.rank{rank}of{world}{ext}"


def extract_features_sharded(engine, images: Sequence[Union[str, np.ndarray, torch.Tensor]], img_ids: Sequence, out_path: str,
                             rank: int = None, world: int = None, batch_size: int = 64, img_size: int = None) -> str:
    """One process per GPU: rank r extracts images r, r + W, r + 2W, ... (dist.shard_indices, the same split the
    captioning path uses) into its own container shard_path(out_path, r, W); no collective is involved, the ranks only
    meet at a barrier at the end when a process group is up.  The reference runs data_generator.py:96-160 in a single
    process.  Returns this rank's path; FeatureShards reads the set back by key."""
    import torch.distributed as td
    from .dist import shard_indices
    up = td.is_available() and td.is_initialized()
    rank = (td.get_rank() if up else 0) if rank is None else rank
    world = (td.get_world_size() if up else 1) if world is None else world
    assert len(images) == len(img_ids)
    idx = shard_indices(len(images), rank, world)
    path = shard_path(out_path, rank, world)
    extract_features(engine, [images[i] for i in idx], [img_ids[i] for i in idx], path, batch_size, img_size)
    if up:
        td.barrier()
    return path


class FeatureShards:
    """Read-side view over the per-rank containers of extract_features_sharded: features by image id, whichever rank
    wrote them (the data loader of the reference, coco_dataloader.py:446,517, looks features up by key only)."""

    def __init__(self, out_path: str, world: int):
        self.paths = [shard_path(out_path, r, world) for r in range(world)]
        self._where: Dict[str, str] = {}
        for p in self.paths:
            for k in self._keys(p):
                self._where[k] = p

    @staticmethod
    def _keys(path: str) -> List[str]:
        try:
            import h5py
            with h5py.File(path, "r") as f:
                return list(f.keys())
        except ImportError:
            with np.load(path) as z:
                return list(z.files)

    def __len__(self) -> int:
        return len(self._where)

    def __contains__(self, img_id) -> bool:
        return feature_key(img_id) in self._where

    def read(self, img_id) -> np.ndarray:
        return read_features(self._where[feature_key(img_id)], img_id)
