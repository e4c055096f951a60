"""CPU-side checks: the C-ABI library loads and exports every symbol include/xnv2_b200.h declares,
the product path fails loudly without a GPU, the shim classes keep the reference's state_dict
layout and call signatures."""
import ctypes
import inspect
import os
import re
from argparse import Namespace

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "xnv2_b200.h")).read()
    return sorted(set(re.findall(r"\b(xn_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from on_device_image_captioning_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 18
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/xnv2_b200.h but not exported"
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS), set(declared) ^ set(_lib.EXPORTED_SYMBOLS)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from on_device_image_captioning_b200 import _lib, config
    from on_device_image_captioning_b200.engine import Engine, _cfg_struct
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(config.swin_tiny_test(), 0)
    lib = _lib.load()
    h = ctypes.c_void_p()
    cs = _cfg_struct(config.swin_tiny_test())
    rc = lib.xn_create(ctypes.byref(cs), 0, ctypes.byref(h))
    assert rc != 0 and b"CUDA" in lib.xn_last_error(None)


def _tiny_model():
    from on_device_image_captioning_b200 import End_ExpansionNet_v2
    words = [f"w{i}" for i in range(512)]
    return End_ExpansionNet_v2(
        swin_img_size=96, swin_patch_size=4, swin_in_chans=3, swin_embed_dim=64, swin_depths=[2, 2], swin_num_heads=[2, 4],
        swin_window_size=12, swin_mlp_ratio=4.0, swin_qkv_bias=True, swin_qk_scale=None, swin_drop_rate=0.0,
        swin_attn_drop_rate=0.0, swin_drop_path_rate=0.0, swin_norm_layer=torch.nn.LayerNorm, swin_ape=False,
        swin_patch_norm=True, swin_use_checkpoint=False, final_swin_dim=128, d_model=128, N_enc=2, N_dec=2, ff=256,
        num_heads=4, num_exp_enc_list=[8, 16, 24], num_exp_dec=4, output_word2idx={w: i for i, w in enumerate(words)},
        output_idx2word=words, max_seq_len=24, drop_args=Namespace(enc=0.0, dec=0.0, enc_input=0.0, dec_input=0.0, other=0.0),
        rank="cuda:0")


def test_shim_state_dict_layout_and_signatures():
    from on_device_image_captioning_b200 import config, synth, ExpansionNet_v2, E2E_ExpansionNet_Captioner
    m = _tiny_model()
    cfg = config.swin_tiny_test()
    sd = synth.make_state_dict(cfg, 0)
    assert list(m.state_dict().keys()) == list(sd.keys())
    # a reference checkpoint also carries geometry buffers: strict loading must accept them
    sd2 = dict(sd)
    sd2["swin_transf.layers.0.blocks.1.attn_mask"] = torch.zeros(4, 144, 144)
    sd2["swin_transf.layers.0.blocks.0.attn.relative_position_index"] = torch.zeros(144, 144, dtype=torch.long)
    m.load_state_dict(sd2)
    assert torch.equal(m.state_dict()["vocab_linear.weight"], sd["vocab_linear.weight"])
    # reference call signatures (legacy_models/captioning_model.py:24-26,111-112)
    fwd = inspect.signature(m.forward)
    assert list(fwd.parameters)[:6] == ["enc_x", "dec_x", "enc_x_num_pads", "dec_x_num_pads", "apply_log_softmax", "mode"]
    bs = inspect.signature(m.beam_search)
    assert list(bs.parameters) == ["enc_input", "enc_input_num_pads", "sos_idx", "eos_idx", "beam_size", "how_many_outputs",
                                   "max_seq_len", "sample_or_max"]
    with pytest.raises(AssertionError, match="requested output per sequence"):
        m.beam_search(torch.zeros(1, 3, 96, 96), [0], 1, 2, beam_size=2, how_many_outputs=3)
    cap = E2E_ExpansionNet_Captioner({"sos_idx": 1, "eos_idx": 2, "beam_size": 3}, model=m)
    assert list(inspect.signature(cap.__call__).parameters) == ["enc_x", "dec_x", "enc_x_num_pads", "dec_x_num_pads", "mode"]
    with pytest.raises(ValueError):
        E2E_ExpansionNet_Captioner({}, model=None)
    assert list(inspect.signature(ExpansionNet_v2.__init__).parameters)[1:14] == [
        "d_model", "N_enc", "N_dec", "ff", "num_heads", "num_exp_enc_list", "num_exp_dec", "output_word2idx",
        "output_idx2word", "max_seq_len", "drop_args", "img_feature_dim", "rank"]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_shim_refuses_cpu_parameters():
    m = _tiny_model()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(enc_x=torch.zeros(1, 3, 96, 96), enc_x_num_pads=[0], mode="beam_search", sos_idx=1, eos_idx=2, beam_size=2)
