import os
import sys
import json

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(bytes(d.pop("meta_json")).decode())
    return d


def golden_setup(name):
    """(golden dict, cfg, state_dict, inputs, pads) rebuilt from the fixture's meta."""
    import torch
    from on_device_image_captioning_b200 import synth
    from on_device_image_captioning_b200.config import XNConfig
    g = load_golden(name)
    m = g["meta"]
    cfg = XNConfig(**m["cfg"])
    sd = synth.make_state_dict(cfg, seed=0, profile=m["profile"], eos_idx=m["eos"])
    fp = sum(float(sd[k].double().abs().sum()) for k in sorted(sd))
    assert abs(fp - m["weight_fingerprint"]) <= 1e-9 * abs(fp), "synthetic weights differ from the fixture's"
    if m["kind"] == "feat":
        x = synth.make_features(cfg, m["B"], seed=1)
    else:
        x = synth.make_images(cfg, m["B"], seed=1, kind=m["kind"])
    assert abs(float(x.double().abs().mean()) - m["input_absmean"]) < 1e-9
    return g, cfg, sd, x, list(m["enc_pads"])


def sub(t, n=4096):
    f = t.reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step].contiguous()


@pytest.fixture(scope="session")
def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))
