"""Writes tests/golden/preprocess_sha256.json: SHA-256 of the float32 tensors the reference's own transform stack
(torchvision Resize/ToTensor/Normalize over Pillow, utils/image_utils.py:5-23) produces for the synthetic images of
tests/test_preprocess_oracle.py.  Run here (CPU): python tests/golden/make_preprocess_golden.py"""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_preprocess_oracle import CASES, synth_image, reference_preprocess

out = {f"{H}x{W}x{S}": hashlib.sha256(reference_preprocess(synth_image(H, W, H * 1000 + W), S).tobytes()).hexdigest()
       for H, W, S in CASES}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "preprocess_sha256.json"), "w"), indent=1)
print("wrote", len(out), "checksums")
