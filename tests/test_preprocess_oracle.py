"""CPU tests: the preprocessing oracle is pinned to Pillow + torchvision themselves (the third-party code the
reference's utils/image_utils.py:5-23 calls), bit for bit, and to the committed golden checksums."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import preprocess_oracle as P

CASES = [(50, 70, 96), (500, 333, 384), (384, 384, 384), (640, 480, 384), (100, 1000, 384), (37, 41, 48), (383, 385, 384), (1, 1, 16),
         (2, 900, 32)]


def synth_image(H, W, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    base = (127 + 90 * np.sin(xx / 7.0 + seed) * np.cos(yy / 11.0))[..., None] + rng.integers(-40, 41, (H, W, 3))
    return np.clip(base, 0, 255).astype(np.uint8)


def reference_preprocess(img, S):
    """The reference's transform stack, applied to an in-memory image."""
    import torchvision
    from PIL import Image
    t1 = torchvision.transforms.Resize((S, S))
    t2 = torchvision.transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    return t2(torchvision.transforms.ToTensor()(t1(Image.fromarray(img, "RGB")))).numpy()


@pytest.mark.parametrize("H,W,S", CASES)
def test_oracle_equals_pillow_torchvision(H, W, S):
    img = synth_image(H, W, H * 1000 + W)
    assert np.array_equal(P.preprocess_rgb8(img, S), reference_preprocess(img, S))


def test_oracle_matches_committed_checksums():
    want = json.load(open(os.path.join(GOLDEN_DIR, "preprocess_sha256.json")))
    for key, digest in want.items():
        H, W, S = map(int, key.split("x"))
        out = P.preprocess_rgb8(synth_image(H, W, H * 1000 + W), S)
        assert hashlib.sha256(out.tobytes()).hexdigest() == digest, key


def test_detokenisation_mirror():
    from on_device_image_captioning_b200 import language_utils as L
    vocab = ["<pad>", "a", "dog", "runs", "SOS", "EOS", "fast"]
    assert L.tokens2description([4, 1, 2, 3, 6, 5, 2], vocab, 4, 5) == "A dog runs fast."
    assert L.batch_tokens2description([[[4, 1, 2, 5, -1]], [[4, 2, 3, -1, -1]]], [[4], [3]], vocab, 4, 5) == ["A dog.", "Dog runs."]
    with pytest.raises(IndexError):
        L.tokens2description([4, 5], vocab, 4, 5)


def test_feature_store_layout(tmp_path):
    """"<img_id>_features" -> (144, 1536) float32, the keys coco_dataloader.py:446 reads."""
    from on_device_image_captioning_b200 import features as F
    path = str(tmp_path / "precalc_features.hdf5")
    w = F.FeatureWriter(path)
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((144, 1536)).astype(np.float32), rng.standard_normal((144, 1536)).astype(np.float32)
    w.add(391895, a)
    w.add("42", b)
    w.close()
    assert F.feature_key(391895) == "391895_features"
    assert np.array_equal(F.read_features(path, 391895), a) and np.array_equal(F.read_features(path, 42), b)
    assert F.read_features(path, 42).dtype == np.float32


def test_detokenisation_matches_reference_known_answers():
    """tests/golden/detok_cases.json: outputs of the reference's tokens2description (made by make_detok_golden.py)."""
    from on_device_image_captioning_b200 import language_utils as L
    g = json.load(open(os.path.join(GOLDEN_DIR, "detok_cases.json")))
    for c in g["cases"]:
        if c["caption"] is None:
            with pytest.raises(IndexError):
                L.tokens2description(c["tokens"], g["vocab"], g["sos"], g["eos"])
        else:
            assert L.tokens2description(c["tokens"], g["vocab"], g["sos"], g["eos"]) == c["caption"]
