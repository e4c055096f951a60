"""Pin the CPU oracle (oracle/xnv2_oracle.py) to outputs of the unmodified reference.

The fixtures under tests/golden/ were produced by tests/golden/make_golden.py, which runs
the reference's own upstream classes (legacy_models/End_ExpansionNet_v2.py,
legacy_models/ExpansionNet_v2.py, legacy_models/captioning_model.py:111-241) on the same
synthetic checkpoint and inputs.  The reference ships no tests of its own.

Tolerance: the oracle issues the same torch CPU ops in the same order as the reference, so
on the same torch build it is bit-identical; 2e-6 absolute leaves room for a different
BLAS thread count on another host.
"""
import numpy as np
import pytest
import torch

from conftest import golden_setup, sub
from oracle import xnv2_oracle as O

TOL = 2e-6
CASES = ["tiny_e2e_peaky", "tiny_e2e_xavier", "feat_peaky_b5", "feat_xavier_b1", "full_e2e_xavier", "full_e2e_peaky",
         "full_p3_288_n2"]      # swin_patch_size=3 / img 288 with N_enc = N_dec = 2 (train.py:381-387, test.py:360-365)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fixture(name):
    torch.set_num_threads(max(1, min(8, torch.get_num_threads())))
    g, cfg, sd, x, pads = golden_setup(name)
    m = g["meta"]
    with torch.no_grad():
        taps = {}
        enc = O.forward_enc(sd, cfg, x, pads, taps)
        if cfg.has_swin:
            np.testing.assert_allclose(sub(taps["swin"]).numpy(), g["swin_sub"], rtol=0, atol=TOL * 10)
        np.testing.assert_allclose(sub(enc).numpy(), g["enc_sub"], rtol=0, atol=TOL)
        tok = torch.from_numpy(g["dec_tokens"])
        dp = g["dec_pads"].tolist()
        lp = O.forward_dec(sd, cfg, enc, pads, tok, dp, True)
        lg = O.forward_dec(sd, cfg, enc, pads, tok, dp, False)
        np.testing.assert_allclose(sub(lp).numpy(), g["dec_logprob_sub"], rtol=0, atol=TOL * 4)
        np.testing.assert_allclose(sub(lg).numpy(), g["dec_logits_sub"], rtol=0, atol=TOL * 4)
        tr = {}
        toks, lps = O.beam_search(sd, cfg, x, pads, m["sos"], m["eos"], m["beam"], m["how_many"], m["max_len"], trace=tr)
    for b in range(m["B"]):
        for j in range(m["how_many"]):
            ln = int(g["beam_len"][b, j])
            assert toks[b][j] == g["beam_tokens"][b, j, :ln].tolist()
    np.testing.assert_allclose(lps.numpy(), g["beam_logprobs"], rtol=0, atol=TOL * 4)


def test_mask_semantics():
    # utils/masking.py:22-47
    m = O.no_peak_and_pad_mask(2, 4, [0, 2])
    assert m[0].tolist() == [[1, 0, 0, 0], [1, 1, 0, 0], [1, 1, 1, 0], [1, 1, 1, 1]]
    assert m[1].tolist() == [[1, 0, 0, 0], [1, 1, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]]
    p = O.pad_mask(1, 3, 4, [1], [2])
    assert p[0].tolist() == [[1, 1, 0, 0], [1, 1, 0, 0], [0, 0, 0, 0]]


def test_geometry_tables():
    # relative_position_index / shift labels are pure geometry
    idx = O.relative_position_index(12)
    assert idx.shape == (144, 144) and int(idx.min()) == 0 and int(idx.max()) == 528
    assert int(idx[0, 0]) == 11 * 23 + 11 and int(idx[0, 143]) == 0 and int(idx[143, 0]) == 528
    lab = O.shift_region_labels(24, 12, 6)
    assert lab[0, 0] == 0 and lab[12, 0] == 3 and lab[18, 18] == 8 and lab[0, 17] == 1


def test_oracle_ensemble_equals_reference_class():
    """the oracle's ensemble search vs the fixture made by the reference's EsembleCaptioningModel."""
    import torch
    from conftest import load_golden
    from on_device_image_captioning_b200 import synth
    from on_device_image_captioning_b200.config import XNConfig
    from oracle import xnv2_oracle as O
    g = load_golden("ens_tiny_e2e")
    m = g["meta"]
    cfg = XNConfig(**m["cfg"])
    sds = [synth.make_state_dict(cfg, seed=s, profile=m["profile"], eos_idx=m["eos"]) for s in m["seeds"]]
    x = synth.make_images(cfg, m["B"], seed=1, kind=m["kind"])
    with torch.no_grad():
        tok, lp = O.beam_search(sds, cfg, x, [0] * m["B"], m["sos"], m["eos"], m["beam"], m["how_many"], m["max_len"])
    for b in range(m["B"]):
        for j in range(m["how_many"]):
            assert tok[b][j] == g["beam_tokens"][b, j, : g["beam_len"][b, j]].tolist()
    assert float((lp - torch.from_numpy(g["beam_logprobs"])).abs().max()) <= 1e-6
