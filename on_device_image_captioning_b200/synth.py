"""Deterministic synthetic checkpoints and inputs.

There is no network for real checkpoints, so parity and benchmarks run on synthetic
weights in the reference's ``model_state_dict`` layout (SURVEY.md Appendix B;
reference utils/saving_utils.py:66-71, demo.py:100-104).  Everything is drawn with
``Tensor.uniform_`` from one seeded CPU generator, in sorted-key order, so the same
bytes come out in the authoring container and on the GPU box.

Profiles
--------
``xavier``  the reference's random init (Q5: every dim>1 parameter is re-drawn with
            xavier-uniform, models/End_ExpansionNet_v2.py:112-114; Swin biases 0,
            LayerNorm 1/0, body Linear biases U(+-1/sqrt(fan_in))).  Near-uniform
            word distribution.
``peaky``   same shapes; non-trivial biases / LayerNorm affine, a sharpened vocabulary
            head and a raised EOS bias, so captions differ per image, top-k margins
            are wide and EOS is emitted (exercises the finished-beam rules,
            reference models/captioning_model.py:322-335).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .config import XNConfig


def state_dict_shapes(cfg: XNConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    """Parameter name -> shape, in the reference's checkpoint naming."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    if cfg.has_swin:
        p = "swin_transf."
        s[p + "patch_embed.proj.weight"] = (cfg.embed_dim, cfg.in_chans, cfg.patch_size, cfg.patch_size)
        s[p + "patch_embed.proj.bias"] = (cfg.embed_dim,)
        s[p + "patch_embed.norm.weight"] = (cfg.embed_dim,)
        s[p + "patch_embed.norm.bias"] = (cfg.embed_dim,)
        nst = len(cfg.depths)
        for si, (C, H, nh, depth) in enumerate(cfg.stage_dims()):
            hid = int(C * cfg.mlp_ratio)
            ws = min(cfg.window_size, H)
            for b in range(depth):
                q = f"{p}layers.{si}.blocks.{b}."
                s[q + "norm1.weight"] = (C,)
                s[q + "norm1.bias"] = (C,)
                s[q + "attn.relative_position_bias_table"] = ((2 * ws - 1) ** 2, nh)
                s[q + "attn.qkv.weight"] = (3 * C, C)
                s[q + "attn.qkv.bias"] = (3 * C,)
                s[q + "attn.proj.weight"] = (C, C)
                s[q + "attn.proj.bias"] = (C,)
                s[q + "norm2.weight"] = (C,)
                s[q + "norm2.bias"] = (C,)
                s[q + "mlp.fc1.weight"] = (hid, C)
                s[q + "mlp.fc1.bias"] = (hid,)
                s[q + "mlp.fc2.weight"] = (C, hid)
                s[q + "mlp.fc2.bias"] = (C,)
            if si < nst - 1:
                q = f"{p}layers.{si}.downsample."
                s[q + "reduction.weight"] = (2 * C, 4 * C)
                s[q + "norm.weight"] = (4 * C,)
                s[q + "norm.bias"] = (4 * C,)
        Cf = cfg.embed_dim * 2 ** (nst - 1)
        s[p + "norm.weight"] = (Cf,)
        s[p + "norm.bias"] = (Cf,)
    d, ff, V = cfg.d_model, cfg.ff, cfg.vocab
    ne = sum(cfg.num_exp_enc_list)

    def lin(name, o, i):
        s[name + ".weight"] = (o, i)
        s[name + ".bias"] = (o,)

    def ln(name):
        s[name + ".weight"] = (d,)
        s[name + ".bias"] = (d,)

    for i in range(cfg.n_enc):
        q = f"encoders.{i}."
        ln(q + "norm_1"); ln(q + "norm_2")
        s[q + "stc_exp.query_exp_vectors.weight"] = (ne, d)
        s[q + "stc_exp.bias_exp_vectors.weight"] = (ne, d)
        for n in ("key_embed", "class_a_embed", "class_b_embed", "selector_embed"):
            lin(q + "stc_exp." + n, d, d)
        lin(q + "ff.linear_1", ff, d); lin(q + "ff.linear_2", d, ff)
    for i in range(cfg.n_dec):
        q = f"decoders.{i}."
        ln(q + "norm_1"); ln(q + "norm_2"); ln(q + "norm_3")
        for n in ("Wq", "Wk", "Wv", "out_linear"):
            lin(q + "mha." + n, d, d)
        for n in ("cond_embed", "key_linear", "class_a_embed", "class_b_embed", "selector_embed"):
            lin(q + "dyn_exp." + n, d, d)
        s[q + "dyn_exp.query_exp_vectors.weight"] = (cfg.num_exp_dec, d)
        s[q + "dyn_exp.bias_exp_vectors.weight"] = (cfg.num_exp_dec, d)
        lin(q + "ff.linear_1", ff, d); lin(q + "ff.linear_2", d, ff)
    lin("input_linear", d, cfg.feat_dim)
    lin("vocab_linear", V, d)
    s["out_embedder.embed.weight"] = (V, d)
    s["pos_encoder.weight"] = (cfg.max_seq_len, d)
    lin("enc_reduce_group", d, d * cfg.n_enc); ln("enc_reduce_norm")
    lin("dec_reduce_group", d, d * cfg.n_dec); ln("dec_reduce_norm")
    return s


def _fans(shape) -> Tuple[int, int]:
    # torch.nn.init._calculate_fan_in_and_fan_out semantics
    rf = 1
    for x in shape[2:]:
        rf *= x
    return shape[1] * rf, shape[0] * rf


def make_state_dict(cfg: XNConfig, seed: int = 0, profile: str = "xavier",
                    eos_idx: int = 77) -> Dict[str, torch.Tensor]:
    assert profile in ("xavier", "peaky")
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * seed + 17)
    sd: Dict[str, torch.Tensor] = OrderedDict()
    peaky = profile == "peaky"
    shapes = state_dict_shapes(cfg)
    for name, shape in shapes.items():
        t = torch.empty(shape, dtype=torch.float32)
        is_norm = ".norm" in name or "norm_" in name or name.endswith("_norm.weight") or name.endswith("_norm.bias")
        if len(shape) > 1:
            fi, fo = _fans(shape)
            a = math.sqrt(6.0 / (fi + fo))
            if peaky and name == "vocab_linear.weight":
                a *= 6.0
            if peaky and name.endswith("relative_position_bias_table"):
                a = 0.5
            t.uniform_(-a, a, generator=g)
        elif is_norm:
            if name.endswith("weight"):
                t.uniform_(0.8, 1.2, generator=g) if peaky else t.fill_(1.0)
            else:
                t.uniform_(-0.1, 0.1, generator=g) if peaky else t.zero_()
        else:  # Linear / conv bias
            if name.startswith("swin_transf.") and not peaky:
                t.zero_()
            else:
                wshape = shapes[name[:-4] + "weight"]
                fi, _ = _fans(wshape)
                b = 1.0 / math.sqrt(fi)
                t.uniform_(-b, b, generator=g)
                if peaky and name == "vocab_linear.bias":
                    t.mul_(4.0)
                    t[eos_idx % shape[0]] += 7.0
        sd[name] = t
    return sd


def make_images(cfg: XNConfig, batch: int, seed: int = 1, kind: str = "mixed") -> torch.Tensor:
    """(B, in_chans, S, S) fp32 synthetic images.

    ``randn``  the distribution the reference's own latency benchmark feeds
               (benchmarking/benchmarking.py:86).
    ``mixed``  per-image low-frequency structure plus noise, in the value range
               ImageNet normalisation produces (utils/image_utils.py:10-16), so that
               different images give clearly different features.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(7919 * seed + 3)
    S = cfg.img_size
    if kind == "randn":
        u1 = torch.empty(batch, cfg.in_chans, S, S).uniform_(1e-7, 1.0, generator=g)
        u2 = torch.empty(batch, cfg.in_chans, S, S).uniform_(0.0, 1.0, generator=g)
        return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * math.pi * u2)
    yy = torch.linspace(0.0, 1.0, S).view(1, 1, S, 1)
    xx = torch.linspace(0.0, 1.0, S).view(1, 1, 1, S)
    fr = torch.empty(batch, cfg.in_chans, 1, 1).uniform_(0.5, 6.0, generator=g)
    ph = torch.empty(batch, cfg.in_chans, 1, 1).uniform_(0.0, 6.28, generator=g)
    am = torch.empty(batch, cfg.in_chans, 1, 1).uniform_(0.3, 1.5, generator=g)
    of = torch.empty(batch, cfg.in_chans, 1, 1).uniform_(-1.0, 1.0, generator=g)
    base = am * torch.sin(2 * math.pi * fr * yy + ph) * torch.cos(2 * math.pi * (fr * 0.7 + 0.3) * xx - ph) + of
    noise = torch.empty(batch, cfg.in_chans, S, S).uniform_(-0.35, 0.35, generator=g)
    return (base + noise).contiguous()


def make_features(cfg: XNConfig, batch: int, seed: int = 1) -> torch.Tensor:
    """(B, enc_len, feat_dim) fp32 features for the decoder-only model (the
    reference's data_generator.py:111-160 stores Swin outputs in this shape)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(104729 * seed + 5)
    t = torch.empty(batch, cfg.enc_len, cfg.feat_dim).uniform_(-1.7, 1.7, generator=g)
    sc = torch.empty(batch, 1, 1).uniform_(0.5, 1.5, generator=g)
    return (t * sc).contiguous()
