"""Model geometry for the ExpansionNet v2 captioning path.

The literals mirror the hyper-parameters the reference hard-codes at every call
site (reference demo.py:68-99, test.py:372-403, benchmarking/benchmarking.py:257-288).
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import List


@dataclass
class XNConfig:
    # Swin backbone (reference models/swin_transformer_mod.py:670-760)
    has_swin: bool = True
    img_size: int = 384
    patch_size: int = 4
    in_chans: int = 3
    embed_dim: int = 192
    depths: List[int] = field(default_factory=lambda: [2, 2, 18, 2])
    swin_heads: List[int] = field(default_factory=lambda: [6, 12, 24, 48])
    window_size: int = 12
    mlp_ratio: float = 4.0
    # captioning body (reference models/End_ExpansionNet_v2.py:11-110)
    feat_dim: int = 1536          # final_swin_dim / img_feature_dim
    d_model: int = 512
    n_enc: int = 3
    n_dec: int = 3
    ff: int = 2048
    num_heads: int = 8
    num_exp_enc_list: List[int] = field(default_factory=lambda: [32, 64, 128, 256, 512])
    num_exp_dec: int = 16
    vocab: int = 10000
    max_seq_len: int = 74
    enc_len: int = 144            # visual tokens entering the expansion encoder

    @property
    def grid0(self) -> int:
        return self.img_size // self.patch_size

    def stage_dims(self):
        """[(C, H, heads, depth)] per Swin stage."""
        out = []
        for s, d in enumerate(self.depths):
            out.append((self.embed_dim * 2 ** s, self.grid0 // 2 ** s, self.swin_heads[s], d))
        return out

    def to_dict(self):
        return asdict(self)


def swin_l_384(vocab: int = 10000, max_seq_len: int = 74, n_enc: int = 3, n_dec: int = 3) -> XNConfig:
    """BASELINE.json configs 1/2/4/5: End_ExpansionNet_v2 Swin-L/384, d=512."""
    return XNConfig(vocab=vocab, max_seq_len=max_seq_len, n_enc=n_enc, n_dec=n_dec)


def features_only(vocab: int = 10000, max_seq_len: int = 74, feat_dim: int = 1536) -> XNConfig:
    """BASELINE.json config 3: decoder-only ExpansionNet_v2 on (144 x feat_dim) features."""
    return XNConfig(has_swin=False, vocab=vocab, max_seq_len=max_seq_len, feat_dim=feat_dim)


def swin_tiny_test(vocab: int = 512, max_seq_len: int = 24) -> XNConfig:
    """A reduced geometry for quick CPU tests: 96x96 image -> 24x24 patches, window 12.

    Stage 0 has 2x2 windows (shifted-window mask exercised), stage 1 is a single
    12x12 window (shift disabled, reference swin_transformer_mod.py:334-337).
    """
    return XNConfig(img_size=96, embed_dim=64, depths=[2, 2], swin_heads=[2, 4],
                    feat_dim=128, d_model=128, n_enc=2, n_dec=2, ff=256, num_heads=4,
                    num_exp_enc_list=[8, 16, 24], num_exp_dec=4, vocab=vocab,
                    max_seq_len=max_seq_len, enc_len=144)
