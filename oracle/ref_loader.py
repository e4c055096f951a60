"""Import the UNMODIFIED reference  --  TEST / MEASUREMENT INFRASTRUCTURE ONLY.

Used (a) in the authoring container to validate oracle/xnv2_oracle.py and to generate the
committed golden fixtures (tests/golden/make_golden.py) and (b) by ``bench.py --impl reference``
to time the reference's own CPU implementation.  /root/reference does not exist on the GPU box:
there the byte-for-byte copy staged by oracle/stage_ref.py into the git-ignored ``oracle/_ref/``
(which travels with the gpurun snapshot) is imported instead.

The batch-correct, log-softmax-correct classes live in ``legacy_models/`` but import
each other as ``models.*``, so the loader stages a *temporary*
directory (never inside this repo) holding ``legacy_models`` under the name ``models``
next to the reference's ``utils`` and puts it on ``sys.path``.
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile
from argparse import Namespace

REFERENCE_ROOT = os.environ.get("XNV2_REFERENCE_ROOT", "/root/reference")
STAGED_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_staged = None


def _tree_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "legacy_models"))


def _staged_available() -> bool:
    return os.path.isfile(os.path.join(STAGED_DIR, "models", "End_ExpansionNet_v2.py")) and \
        os.path.isfile(os.path.join(STAGED_DIR, "utils", "masking.py"))


def available() -> bool:
    return _tree_available() or _staged_available()


def stage() -> str:
    global _staged
    if _staged is None:
        if not available():
            raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT} and nothing staged at {STAGED_DIR}")
        if _tree_available():
            tmp = tempfile.mkdtemp(prefix="xnv2_ref_")
            shutil.copytree(os.path.join(REFERENCE_ROOT, "legacy_models"), os.path.join(tmp, "models"))
            shutil.copytree(os.path.join(REFERENCE_ROOT, "utils"), os.path.join(tmp, "utils"))
        else:
            tmp = STAGED_DIR
        for m in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")]:
            del sys.modules[m]
        sys.path.insert(0, tmp)
        _staged = tmp
    return _staged


def drop_args():
    return Namespace(enc=0.0, dec=0.0, enc_input=0.0, dec_input=0.0, other=0.0)


def build_reference_model(cfg, state_dict, vocab_words=None, rank="cpu"):
    """Construct the reference (legacy/upstream) model for ``cfg`` on ``rank`` (CPU by default) and load
    ``state_dict`` into it.  Buffers the reference registers (relative_position_index,
    attn_mask) are geometry-only and are left as the reference computes them."""
    import torch
    stage()
    words = vocab_words or [f"w{i}" for i in range(cfg.vocab)]
    w2i = {w: i for i, w in enumerate(words)}
    if cfg.has_swin:
        from models.End_ExpansionNet_v2 import End_ExpansionNet_v2
        m = End_ExpansionNet_v2(
            swin_img_size=cfg.img_size, swin_patch_size=cfg.patch_size, swin_in_chans=cfg.in_chans,
            swin_embed_dim=cfg.embed_dim, swin_depths=list(cfg.depths), swin_num_heads=list(cfg.swin_heads),
            swin_window_size=cfg.window_size, swin_mlp_ratio=cfg.mlp_ratio, swin_qkv_bias=True, swin_qk_scale=None,
            swin_drop_rate=0.0, swin_attn_drop_rate=0.0, swin_drop_path_rate=0.0,
            swin_norm_layer=torch.nn.LayerNorm, swin_ape=False, swin_patch_norm=True, swin_use_checkpoint=False,
            final_swin_dim=cfg.feat_dim, d_model=cfg.d_model, N_enc=cfg.n_enc, N_dec=cfg.n_dec, ff=cfg.ff,
            num_heads=cfg.num_heads, num_exp_enc_list=list(cfg.num_exp_enc_list), num_exp_dec=cfg.num_exp_dec,
            output_word2idx=w2i, output_idx2word=words, max_seq_len=cfg.max_seq_len, drop_args=drop_args(), rank=rank)
    else:
        from models.ExpansionNet_v2 import ExpansionNet_v2
        m = ExpansionNet_v2(
            d_model=cfg.d_model, N_enc=cfg.n_enc, N_dec=cfg.n_dec, ff=cfg.ff, num_heads=cfg.num_heads,
            num_exp_enc_list=list(cfg.num_exp_enc_list), num_exp_dec=cfg.num_exp_dec,
            output_word2idx=w2i, output_idx2word=words, max_seq_len=cfg.max_seq_len, drop_args=drop_args(),
            img_feature_dim=cfg.feat_dim, rank=rank)
    missing, unexpected = m.load_state_dict(state_dict, strict=False)
    bad = [k for k in missing if not (k.endswith("relative_position_index") or k.endswith("attn_mask"))]
    assert not bad and not unexpected, (bad, unexpected)
    if str(rank) != "cpu":
        m = m.to(rank)
    return m.eval()
