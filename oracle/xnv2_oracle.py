"""CPU oracle for the ExpansionNet v2 captioning path  --  TEST INFRASTRUCTURE ONLY.

A functional, fp32, torch-CPU restatement of the reference's arithmetic for the hot
path named by BASELINE.json (Swin-L backbone -> static-expansion encoder ->
dynamic-expansion decoder -> beam search).  It works straight off a checkpoint
``state_dict`` in the reference's layout and keeps the
reference's algorithmic shape on purpose: whole-prefix re-decode every step, dense
masks, no caches.  Each function cites the reference lines it restates.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module, and only as the checker or the
timed CPU baseline.  The product path (the CUDA library behind
``include/xnv2_b200.h``) never calls it.

Parity pin: the reference has no tests or golden vectors of its own,
so this oracle is pinned against outputs of the *unmodified reference itself*, run in
the authoring container by ``tests/golden/make_golden.py`` (legacy classes imported
from /root/reference under the package name ``models``) and committed under
``tests/golden/``; ``tests/test_oracle_golden.py`` holds the comparison.

All citations are relative to the reference root.  "legacy" = legacy_models/ (the
upstream, batch-correct classes).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

SHIFT_MASK_VALUE = -100.0      # swin_transformer_mod.py:389-391 (legacy :297)
CROSS_MASK_FILL = -1e4         # layers.py:284 (legacy :249)
EXP_EPS = 1e-9                 # layers.py:106,208 (legacy :94,186)
LN_EPS = 1e-5                  # nn.LayerNorm default


def _ln(x: torch.Tensor, sd: SD, name: str) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], LN_EPS)


def _lin(x: torch.Tensor, sd: SD, name: str) -> torch.Tensor:
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


# --------------------------------------------------------------------------------------
# Swin backbone
# --------------------------------------------------------------------------------------

def relative_position_index(ws: int) -> torch.Tensor:
    """(ws*ws, ws*ws) int64: (yi-yj+ws-1)*(2ws-1) + (xi-xj+ws-1), tokens row-major in the
    window.  legacy swin_transformer_mod.py:163-172."""
    ys, xs = torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")
    y = ys.reshape(-1)
    x = xs.reshape(-1)
    return (y[:, None] - y[None, :] + ws - 1) * (2 * ws - 1) + (x[:, None] - x[None, :] + ws - 1)


def shift_region_labels(H: int, ws: int, shift: int) -> torch.Tensor:
    """(H, H) region label of every position of the *shifted* frame: 3 bands per axis,
    [0,H-ws), [H-ws,H-shift), [H-shift,H).  legacy swin_transformer_mod.py:281-292."""
    band = torch.zeros(H, dtype=torch.int64)
    band[H - ws:H - shift] = 1
    band[H - shift:] = 2
    return band[:, None] * 3 + band[None, :]


def _to_windows(x: torch.Tensor, ws: int) -> torch.Tensor:
    """(B,H,W,C) -> (B*nW, ws*ws, C); legacy swin_transformer_mod.py:102-115."""
    B, H, W, C = x.shape
    x = x.reshape(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, ws * ws, C)


def _from_windows(w: torch.Tensor, ws: int, B: int, H: int, W: int) -> torch.Tensor:
    """inverse of _to_windows; legacy swin_transformer_mod.py:118-132."""
    C = w.shape[-1]
    x = w.reshape(B, H // ws, W // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(B, H, W, C)


def window_attention(xw: torch.Tensor, sd: SD, pre: str, heads: int, ws: int,
                     mask: Optional[torch.Tensor]) -> torch.Tensor:
    """W-MSA on partitioned windows.  legacy swin_transformer_mod.py:183-214."""
    Bw, N, C = xw.shape
    hd = C // heads
    qkv = _lin(xw, sd, pre + "qkv").reshape(Bw, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (hd ** -0.5), qkv[1], qkv[2]
    att = q @ k.transpose(-2, -1)
    table = sd[pre + "relative_position_bias_table"]
    bias = table[relative_position_index(ws).reshape(-1)].reshape(N, N, heads).permute(2, 0, 1)
    att = att + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        att = att.reshape(Bw // nW, nW, heads, N, N) + mask[None, :, None]
        att = att.reshape(Bw, heads, N, N)
    att = torch.softmax(att, dim=-1)
    out = (att @ v).transpose(1, 2).reshape(Bw, N, C)
    return _lin(out, sd, pre + "proj")


def swin_block(x: torch.Tensor, sd: SD, pre: str, H: int, heads: int, ws_cfg: int, shifted: bool) -> torch.Tensor:
    """One SwinTransformerBlock.  legacy swin_transformer_mod.py:252-340."""
    B, L, C = x.shape
    ws, shift = ws_cfg, (ws_cfg // 2 if shifted else 0)
    if H <= ws_cfg:                      # single window: no shift (legacy :262-265)
        ws, shift = H, 0
    h = _ln(x, sd, pre + "norm1").reshape(B, H, H, C)
    mask = None
    if shift > 0:
        h = torch.roll(h, shifts=(-shift, -shift), dims=(1, 2))
        lab = _to_windows(shift_region_labels(H, ws, shift).reshape(1, H, H, 1).float(), ws).reshape(-1, ws * ws)
        diff = lab[:, None, :] - lab[:, :, None]
        mask = torch.where(diff != 0, torch.tensor(SHIFT_MASK_VALUE), torch.tensor(0.0))
    a = window_attention(_to_windows(h, ws), sd, pre + "attn.", heads, ws, mask)
    h = _from_windows(a, ws, B, H, H)
    if shift > 0:
        h = torch.roll(h, shifts=(shift, shift), dims=(1, 2))
    x = x + h.reshape(B, L, C)
    m = _lin(F.gelu(_lin(_ln(x, sd, pre + "norm2"), sd, pre + "mlp.fc1")), sd, pre + "mlp.fc2")
    return x + m


def patch_merging(x: torch.Tensor, sd: SD, pre: str, H: int) -> torch.Tensor:
    """2x2 neighbour concat (order (0,0),(1,0),(0,1),(1,1)) -> LN(4C) -> Linear 4C->2C, no
    bias.  legacy swin_transformer_mod.py:377-398."""
    B, L, C = x.shape
    g = x.reshape(B, H, H, C)
    cat = torch.cat([g[:, 0::2, 0::2], g[:, 1::2, 0::2], g[:, 0::2, 1::2], g[:, 1::2, 1::2]], dim=-1)
    cat = cat.reshape(B, (H // 2) * (H // 2), 4 * C)
    return F.linear(_ln(cat, sd, pre + "norm"), sd[pre + "reduction.weight"])


def swin_forward_features(sd: SD, cfg, img: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """(B,3,S,S) -> (B, L_last, C_last).  legacy swin_transformer_mod.py:630-642, PatchEmbed
    :511-519 (conv k=stride=patch, flatten, LN)."""
    p = "swin_transf."
    x = F.conv2d(img, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"], stride=cfg.patch_size)
    x = x.flatten(2).transpose(1, 2)
    x = _ln(x, sd, p + "patch_embed.norm")
    if taps is not None:
        taps["patch_embed"] = x
    nst = len(cfg.depths)
    for si, (C, H, heads, depth) in enumerate(cfg.stage_dims()):
        for b in range(depth):
            x = swin_block(x, sd, f"{p}layers.{si}.blocks.{b}.", H, heads, cfg.window_size, shifted=(b % 2 == 1))
            if taps is not None and b < 2:
                taps[f"stage{si}.block{b}"] = x
        if si < nst - 1:
            x = patch_merging(x, sd, f"{p}layers.{si}.downsample.", H)
        if taps is not None:
            taps[f"stage{si}.out"] = x
    return _ln(x, sd, p + "norm")


# --------------------------------------------------------------------------------------
# Masks (utils/masking.py:22-47) -- dense, as the reference builds them
# --------------------------------------------------------------------------------------

def pad_mask(bs: int, rows: int, cols: int, pad_row: Sequence[int], pad_col: Sequence[int]) -> torch.Tensor:
    m = torch.ones(bs, rows, cols)
    for b in range(bs):
        m[b, :, cols - int(pad_col[b]):] = 0
        m[b, rows - int(pad_row[b]):, :] = 0
    return m


def no_peak_and_pad_mask(bs: int, t: int, num_pads: Sequence[int]) -> torch.Tensor:
    m = torch.tril(torch.ones(t, t)).unsqueeze(0).repeat(bs, 1, 1)
    for b in range(bs):
        m[b, :, t - int(num_pads[b]):] = 0
        m[b, t - int(num_pads[b]):, :] = 0
    return m


# --------------------------------------------------------------------------------------
# Expansion encoder
# --------------------------------------------------------------------------------------

def static_expansion(x: torch.Tensor, sd: SD, pre: str, groups: Sequence[int], mask: torch.Tensor) -> torch.Tensor:
    """StaticExpansionBlock.forward with n_indexes = arange(sum(groups)).  legacy
    layers.py:45-90.  x is already layer-normed; mask is (B, E, N) ones/zeros."""
    d = x.shape[-1]
    qexp = sd[pre + "query_exp_vectors.weight"]
    bexp = sd[pre + "bias_exp_vectors.weight"]
    key = _lin(x, sd, pre + "key_embed")
    z = torch.matmul(qexp, key.transpose(-1, -2)) / (d ** 0.5)            # (B,E,N)
    a_fw = F.relu(z).masked_fill(mask == 0, 0.0)
    b_fw = F.relu(-z).masked_fill(mask == 0, 0.0)
    a_fw = a_fw / (a_fw.sum(dim=-1, keepdim=True) + EXP_EPS)
    b_fw = b_fw / (b_fw.sum(dim=-1, keepdim=True) + EXP_EPS)
    ca = torch.matmul(a_fw, _lin(x, sd, pre + "class_a_embed")) + bexp     # (B,E,d)
    cb = torch.matmul(b_fw, _lin(x, sd, pre + "class_b_embed")) + bexp
    zt = z.transpose(-2, -1)                                               # (B,N,E)
    a_bw, b_bw = F.relu(zt), F.relu(-zt)
    a_parts, b_parts, lo = [], [], 0
    for gsz in groups:                                                     # per-group normalise :70-80
        hi = lo + gsz
        a_parts.append(a_bw[:, :, lo:hi] / (a_bw[:, :, lo:hi].sum(dim=-1, keepdim=True) + EXP_EPS))
        b_parts.append(b_bw[:, :, lo:hi] / (b_bw[:, :, lo:hi].sum(dim=-1, keepdim=True) + EXP_EPS))
        lo = hi
    out_a = torch.matmul(torch.cat(a_parts, dim=-1), ca) / len(groups)
    out_b = torch.matmul(torch.cat(b_parts, dim=-1), cb) / len(groups)
    sel = torch.sigmoid(_lin(x, sd, pre + "selector_embed"))
    return sel * out_a + (1 - sel) * out_b


def _ff(x: torch.Tensor, sd: SD, pre: str) -> torch.Tensor:
    """FeedForward: Linear, ReLU, Linear.  legacy layers.py:267-270."""
    return _lin(F.relu(_lin(x, sd, pre + "linear_1")), sd, pre + "linear_2")


def encoder_body(sd: SD, cfg, feats: torch.Tensor, enc_pads: Sequence[int], taps: Optional[dict] = None) -> torch.Tensor:
    """input_linear -> N_enc pre-LN encoder layers -> concat -> reduce group + residual ->
    LN.  legacy End_ExpansionNet_v2.py:82-101 / ExpansionNet_v2.py:48-66; EncoderLayer
    legacy layers.py:104-109."""
    B, N, _ = feats.shape
    x = _lin(feats, sd, "input_linear")
    E = sum(cfg.num_exp_enc_list)
    mask = pad_mask(B, E, N, [0] * B, enc_pads)
    outs = []
    for i in range(cfg.n_enc):
        pre = f"encoders.{i}."
        x = x + static_expansion(_ln(x, sd, pre + "norm_1"), sd, pre + "stc_exp.", cfg.num_exp_enc_list, mask)
        x = x + _ff(_ln(x, sd, pre + "norm_2"), sd, pre + "ff.")
        outs.append(x)
        if taps is not None:
            taps[f"enc{i}"] = x
    x = x + _lin(torch.cat(outs, dim=-1), sd, "enc_reduce_group")
    return _ln(x, sd, "enc_reduce_norm")


def forward_enc(sd: SD, cfg, enc_input: torch.Tensor, enc_pads: Optional[Sequence[int]] = None,
                taps: Optional[dict] = None) -> torch.Tensor:
    B = enc_input.shape[0]
    if cfg.has_swin:
        feats = swin_forward_features(sd, cfg, enc_input, taps)
        if taps is not None:
            taps["swin"] = feats
        pads = [0] * B                      # "End to End case have no padding", legacy End_...:78,84
    else:
        feats = enc_input
        pads = list(enc_pads) if enc_pads is not None else [0] * B
    return encoder_body(sd, cfg, feats, pads, taps)


# --------------------------------------------------------------------------------------
# Decoder
# --------------------------------------------------------------------------------------

def dynamic_expansion(x: torch.Tensor, sd: SD, pre: str, n_exp: int, mask: torch.Tensor) -> torch.Tensor:
    """DynamicExpansionBlock.forward with n_indexes = arange(n_exp).  legacy layers.py:138-182.
    x (R,t,d) layer-normed; mask (R,t,t) causal-and-pad."""
    R, t, d = x.shape
    cond = _lin(x, sd, pre + "cond_embed").reshape(R, t, 1, d)
    qexp = (sd[pre + "query_exp_vectors.weight"].reshape(1, 1, n_exp, d) + cond).reshape(R, t * n_exp, d)
    bexp = (sd[pre + "bias_exp_vectors.weight"].reshape(1, 1, n_exp, d) + cond).reshape(R, t * n_exp, d)
    key = _lin(x, sd, pre + "key_linear")
    z = torch.matmul(qexp, key.transpose(-1, -2)) / (d ** 0.5)             # (R, t*n_exp, t)
    m1 = mask.unsqueeze(2).expand(R, t, n_exp, t).reshape(R, t * n_exp, t)
    a_fw = F.relu(z).masked_fill(m1 == 0, 0.0)
    b_fw = F.relu(-z).masked_fill(m1 == 0, 0.0)
    a_fw = a_fw / (a_fw.sum(dim=-1, keepdim=True) + EXP_EPS)
    b_fw = b_fw / (b_fw.sum(dim=-1, keepdim=True) + EXP_EPS)
    ca = torch.matmul(a_fw, _lin(x, sd, pre + "class_a_embed"))
    cb = torch.matmul(b_fw, _lin(x, sd, pre + "class_b_embed"))
    m2 = mask.unsqueeze(-1).expand(R, t, t, n_exp).reshape(R, t, t * n_exp)
    zt = z.transpose(-2, -1)
    a_bw = F.relu(zt).masked_fill(m2 == 0, 0.0)
    b_bw = F.relu(-zt).masked_fill(m2 == 0, 0.0)
    a_bw = a_bw / (a_bw.sum(dim=-1, keepdim=True) + EXP_EPS)
    b_bw = b_bw / (b_bw.sum(dim=-1, keepdim=True) + EXP_EPS)
    out_a = torch.matmul(a_bw, ca + bexp)
    out_b = torch.matmul(b_bw, cb + bexp)
    sel = torch.sigmoid(_lin(x, sd, pre + "selector_embed"))
    return sel * out_a + (1 - sel) * out_b


def cross_attention(q_in: torch.Tensor, kv: torch.Tensor, sd: SD, pre: str, heads: int, mask: torch.Tensor) -> torch.Tensor:
    """MultiHeadAttention.forward(q, k=v=encoder output).  legacy layers.py:231-257."""
    R, t, d = q_in.shape
    n = kv.shape[1]
    dk = d // heads
    k = _lin(kv, sd, pre + "Wk").reshape(R, n, heads, dk).transpose(1, 2)
    q = _lin(q_in, sd, pre + "Wq").reshape(R, t, heads, dk).transpose(1, 2)
    v = _lin(kv, sd, pre + "Wv").reshape(R, n, heads, dk).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) / (dk ** 0.5)
    s = s.masked_fill(mask.unsqueeze(1) == 0, CROSS_MASK_FILL)
    s = torch.softmax(s, dim=-1)
    o = torch.matmul(s, v).permute(0, 2, 1, 3).reshape(R, t, d)
    return _lin(o, sd, pre + "out_linear")


def forward_dec(sd: SD, cfg, cross: torch.Tensor, enc_pads: Sequence[int], tokens: torch.Tensor,
                dec_pads: Sequence[int], apply_log_softmax: bool = False,
                taps: Optional[dict] = None) -> torch.Tensor:
    """Whole-prefix decoder pass -> (R, t, V).  legacy End_ExpansionNet_v2.py:103-138 /
    ExpansionNet_v2.py:68-103; DecoderLayer legacy layers.py:200-212; EmbeddingLayer :16-17."""
    R, t = tokens.shape
    n = cross.shape[1]
    if cfg.has_swin:
        enc_pads = [0] * R
    causal = no_peak_and_pad_mask(R, t, dec_pads)
    xmask = pad_mask(R, t, n, dec_pads, enc_pads)
    y = sd["out_embedder.embed.weight"][tokens] * math.sqrt(float(cfg.d_model))
    y = y + sd["pos_encoder.weight"][torch.arange(t)].unsqueeze(0)
    outs = []
    for i in range(cfg.n_dec):
        pre = f"decoders.{i}."
        y = y + dynamic_expansion(_ln(y, sd, pre + "norm_1"), sd, pre + "dyn_exp.", cfg.num_exp_dec, causal)
        y = y + cross_attention(_ln(y, sd, pre + "norm_2"), cross, sd, pre + "mha.", cfg.num_heads, xmask)
        y = y + _ff(_ln(y, sd, pre + "norm_3"), sd, pre + "ff.")
        outs.append(y)
        if taps is not None:
            taps[f"dec{i}"] = y
    y = y + _lin(torch.cat(outs, dim=-1), sd, "dec_reduce_group")
    y = _ln(y, sd, "dec_reduce_norm")
    y = _lin(y, sd, "vocab_linear")
    return torch.log_softmax(y, dim=-1) if apply_log_softmax else y


# --------------------------------------------------------------------------------------
# Beam search (sample_or_max == 'max')
# --------------------------------------------------------------------------------------

def _min_gap(sorted_vals: torch.Tensor) -> float:
    if sorted_vals.shape[-1] < 2:
        return float("inf")
    return float((sorted_vals[..., :-1] - sorted_vals[..., 1:]).min())


def beam_search(sd: SD, cfg, enc_input: torch.Tensor, enc_pads: Sequence[int], sos_idx: int, eos_idx: int,
                beam_size: int = 3, how_many_outputs: int = 1, max_seq_len: int = 20,
                trace: Optional[dict] = None) -> Tuple[List[List[List[int]]], torch.Tensor]:
    """legacy captioning_model.py:111-241 (== models/captioning_model.py:220-427), 'max' branch.

    ``trace`` (optional dict) receives the decision margins: the smallest gap between
    the k-th and (k+1)-th value of every top-k of the search, per image, so that
    bit-exactness claims about another implementation can be qualified.
    """
    assert how_many_outputs <= beam_size, "requested output per sequence must be lower than beam width"
    B = enc_input.shape[0]
    k = beam_size
    enc_pads = list(enc_pads)
    # ``sd`` may be a list of state dicts: the ensemble of legacy_models/ensemble_captioning_model.py (test.py:334), whose
    # step distribution is log(mean_m softmax(logits_m)) (:55-84) fed to exactly the same search (:86-241)
    sds = list(sd) if isinstance(sd, (list, tuple)) else None
    if sds is not None:
        crosses = [forward_enc(s_, cfg, enc_input, enc_pads) for s_ in sds]
        cross = crosses[0]
        def _fd(_sd, _cfg, cross_in, pads_in, tok_in, dpads_in, _apply):
            rep = cross_in.shape[0] // crosses[0].shape[0]
            ps = []
            for s_, c_ in zip(sds, crosses):
                c_in = c_ if rep == 1 else c_.unsqueeze(1).expand(B, rep, c_.shape[1], c_.shape[2]).reshape(B * rep, c_.shape[1], c_.shape[2])
                ps.append(torch.softmax(forward_dec(s_, _cfg, c_in, pads_in, tok_in, dpads_in, False).unsqueeze(0), dim=-1))
            return torch.cat(ps, dim=0).mean(dim=0).log()
    else:
        cross = forward_enc(sd, cfg, enc_input, enc_pads)
        _fd = forward_dec
    n, d = cross.shape[1], cross.shape[2]
    vocab_margin = torch.full((B,), float("inf"))
    merge_margin = torch.full((B,), float("inf"))

    # step 0 (:120-140): one row per image holding [SOS]
    tok0 = torch.full((B, 1), sos_idx, dtype=torch.long)
    lp = _fd(sd, cfg, cross, enc_pads, tok0, [0] * B, True)          # (B,1,V)
    top_v, top_i = torch.topk(lp, k=k, sorted=True)
    if lp.shape[-1] > k:
        mv = torch.topk(lp, k=k + 1, sorted=True).values
        vocab_margin = torch.minimum(vocab_margin, (mv[:, 0, :-1] - mv[:, 0, 1:]).min(dim=-1).values)
    classes = torch.cat([torch.full((B, k, 1), sos_idx, dtype=torch.long), top_i.transpose(-2, -1)], dim=-1)
    logprobs = torch.cat([torch.zeros(B, k, 1), top_v.transpose(-2, -1)], dim=-1)

    cross_rep = cross.unsqueeze(1).expand(B, k, n, d).reshape(B * k, n, d)       # :142-146
    enc_pads_rep = [enc_pads[i] for i in range(B) for _ in range(k)]
    cumul = logprobs.sum(dim=-1, keepdim=True)
    num_elem = torch.full((B * k,), 2, dtype=torch.long)
    ar = torch.arange(B).unsqueeze(-1)

    for t in range(2, max_seq_len):                                          # :155
        flat = classes.reshape(B * k, t)
        lp = _fd(sd, cfg, cross_rep, enc_pads_rep, flat, (t - num_elem).tolist(), True)[:, t - 1, :]
        tv, ti = torch.topk(lp, k=k, sorted=True)                              # :163
        if lp.shape[-1] > k:
            mv = torch.topk(lp, k=k + 1, sorted=True).values
            gaps = (mv[:, :-1] - mv[:, 1:]).min(dim=-1).values
            has_eos_rows = (flat == eos_idx).any(dim=-1)                     # finished rows are overridden below
            gaps = torch.where(has_eos_rows, torch.full_like(gaps, float("inf")), gaps)
            vocab_margin = torch.minimum(vocab_margin, gaps.reshape(B, k).min(dim=-1).values)
        word_cls = ti[:, :k].reshape(B, k, k)
        word_lp = tv[:, :k].reshape(B, k, k).clone()
        has_eos = (classes == eos_idx).any(dim=-1, keepdim=True)             # :176-184
        word_lp[:, :, 0:1] = torch.where(has_eos, torch.zeros_like(word_lp[:, :, 0:1]), word_lp[:, :, 0:1])
        word_lp[:, :, 1:] = torch.where(has_eos, torch.full_like(word_lp[:, :, 1:], -999.0), word_lp[:, :, 1:])
        comp = (cumul + word_lp).reshape(B, k * k)                           # :186-191
        _, ci = torch.topk(comp, k=k, sorted=True)
        if k * k > k:
            mv = torch.topk(comp, k=k + 1, sorted=True).values
            merge_margin = torch.minimum(merge_margin, (mv[:, :-1] - mv[:, 1:]).min(dim=-1).values)
        parent, word = ci // k, ci % k
        classes = classes[ar, parent]                                        # :193-212
        logprobs = logprobs[ar, parent]
        new_cls = word_cls[ar, parent].gather(-1, word.unsqueeze(-1))
        new_lp = word_lp[ar, parent].gather(-1, word.unsqueeze(-1))
        classes = torch.cat([classes, new_cls], dim=-1)
        logprobs = torch.cat([logprobs, new_lp], dim=-1)
        cumul = logprobs.sum(dim=-1, keepdim=True)                           # :214
        num_elem = num_elem.reshape(B, k)[ar, parent].reshape(B * k)         # :217-220
        had_eos = (classes[:, :, :-1] == eos_idx).any(dim=-1).reshape(B * k)
        num_elem = num_elem + (~had_eos).long()
        if int((num_elem != t + 1).sum()) == B * k:                          # :222
            break

    score = cumul / num_elem.reshape(B, k, 1)                                # :226-227
    fv, fi = torch.topk(score.squeeze(-1), k=k)
    final_margin = (fv[:, :-1] - fv[:, 1:]).min(dim=-1).values if k > 1 else torch.full((B,), float("inf"))
    res_tok: List[List[List[int]]] = []
    res_lp: List[torch.Tensor] = []
    for i in range(B):
        res_tok.append([])
        for j in range(how_many_outputs):
            idx = int(fi[i, j])
            ln = int(num_elem[i * k + idx])
            res_tok[i].append(classes[i, idx, :ln].tolist())
            res_lp.append(logprobs[i, idx, :ln])
    lp_out = torch.nn.utils.rnn.pad_sequence(res_lp, batch_first=True).reshape(B, how_many_outputs, -1)
    if trace is not None:
        trace["vocab_margin"] = vocab_margin
        trace["merge_margin"] = merge_margin
        trace["final_margin"] = final_margin
        trace["enc_out"] = cross
    return res_tok, lp_out
