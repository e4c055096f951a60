"""Parity checks of the CUDA path (through the C ABI) against the CPU oracle and the golden
fixtures.  Each check returns a list of (label, measured, tolerance) triples; pytest asserts
measured <= tolerance (tests/test_gpu_parity.py) and tools/gpu_selftest.py prints them all.

Tolerances.  fp32 mode: BASELINE.json asks for logits/features within 1e-5 *relative*; we
measure max|a-b| / max|b| (the tensors are O(1)) and allow 2e-5 on the 24-block Swin output,
1e-5 elsewhere.  bf16 mode: 2e-3 relative is the stated target for features/logits; measured
as relative Frobenius error.  Captions: bit-exact token sequences in fp32 mode wherever the
oracle's decision margin (recorded in the fixture) exceeds 2e-5.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np
import torch

from conftest import golden_setup, sub
from oracle import xnv2_oracle as O
from on_device_image_captioning_b200.engine import Engine, unpack_beam_results
from on_device_image_captioning_b200.config import XNConfig

Triple = Tuple[str, float, float]
_engines = {}


def rel_max(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_fro(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def bare_engine() -> Engine:
    if "bare" not in _engines:
        from on_device_image_captioning_b200.config import swin_tiny_test
        _engines["bare"] = Engine(swin_tiny_test(), 0)
    return _engines["bare"]


def engine_for(name: str, precision: str):
    key = (name, precision)
    if key not in _engines:
        # keep at most one big model resident per precision
        for k in [k for k in _engines if k != "bare" and k[0] != name]:
            _engines.pop(k)[0].close()
        g, cfg, sd, x, pads = golden_setup(name)
        e = Engine(cfg, 0)
        e.load_state_dict(sd, precision)
        _engines[key] = (e, g, cfg, sd, x, pads)
    return _engines[key]


# ------------------------------------------------------------------ single kernels
def check_layernorm() -> List[Triple]:
    e = bare_engine()
    out = []
    g = torch.Generator().manual_seed(0)
    for rows, C in [(7, 192), (1000, 768), (33, 1536), (5, 3072), (64, 512)]:
        x = torch.randn(rows, C, generator=g) * 3 + 0.5
        w = torch.rand(C, generator=g) + 0.5
        b = torch.randn(C, generator=g)
        y = e.op_layernorm(x, w, b)
        ref = torch.nn.functional.layer_norm(x, (C,), w, b, 1e-5)
        out.append((f"layernorm[{rows}x{C}] max-abs", float((y.cpu() - ref).abs().max()), 2e-5))
    return out


def check_linear_fp32() -> List[Triple]:
    e = bare_engine()
    out = []
    g = torch.Generator().manual_seed(1)
    for (M, N, K, act, res) in [(200, 192, 192, 0, False), (1000, 576, 192, 0, True), (300, 768, 3072, 1, False),
                                (192, 10000, 512, 0, False), (77, 512, 2048, 2, True), (5, 64, 128, 0, False),
                                (4096, 2304, 768, 0, False)]:
        x = torch.randn(M, K, generator=g)
        w = torch.randn(N, K, generator=g) / math.sqrt(K)
        b = torch.randn(N, generator=g)
        r = torch.randn(M, N, generator=g) if res else None
        y = e.op_linear(x, w, b, r, act, "fp32")
        ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
        if act == 1:
            ref = torch.nn.functional.gelu(ref)
        elif act == 2:
            ref = torch.relu(ref)
        if res:
            ref = ref + r.double()
        out.append((f"linear_fp32[{M}x{N}x{K} act{act} res{int(res)}] rel-max", rel_max(y, ref), 1e-5))
    return out


def check_linear_bf16(precision: str = "bf16") -> List[Triple]:
    """tcgen05 GEMM vs an fp64 product of the same 16-bit-rounded operands (so only the fp32
    accumulation order differs)."""
    e = bare_engine()
    rnd = (lambda t: t.bfloat16()) if precision == "bf16" else (lambda t: t.half())
    out = []
    g = torch.Generator().manual_seed(2)
    for (M, N, K, act, res) in [(128, 128, 64, 0, False), (128, 256, 128, 0, False), (300, 192, 192, 0, True),
                                (1000, 576, 192, 0, False), (4096, 3072, 768, 1, False), (2500, 768, 3072, 0, True),
                                (777, 1536, 1536, 0, False), (20000, 384, 384, 0, False), (64, 10000, 512, 0, False)]:
        x = torch.randn(M, K, generator=g)
        w = torch.randn(N, K, generator=g) / math.sqrt(K)
        b = torch.randn(N, generator=g)
        r = torch.randn(M, N, generator=g) if res else None
        y = e.op_linear(x, w, b, r, act, precision)
        xb, wb = rnd(x).double(), rnd(w).double()
        ref = torch.nn.functional.linear(xb, wb, b.double())
        if act == 1:
            ref = torch.nn.functional.gelu(ref)
        if res:
            ref = ref + r.double()
        out.append((f"linear_{precision}_tcgen05[{M}x{N}x{K} act{act} res{int(res)}] rel-max", rel_max(y, ref), 2e-5))
    return out


def check_linear_skinny(precision: str = "fp16") -> List[Triple]:
    """Decoder-step GEMM (mma.sync, resident operands, cluster split-K, LayerNorm fused into the A load) vs an fp64
    product of the same 16-bit-rounded operands."""
    e = bare_engine()
    rnd = (lambda t: t.bfloat16()) if precision == "bf16" else (lambda t: t.half())
    out = []
    g = torch.Generator().manual_seed(12)
    # (M, N, K, act, res, ln, x16): the decoder's shapes at batch 64 x beam 3, ragged rows / columns, batch 1
    cases = [(192, 2560, 512, 0, False, True, False), (192, 512, 512, 0, False, False, True), (192, 512, 512, 0, True, False, True),
             (192, 2048, 512, 2, False, True, False), (192, 512, 2048, 0, True, False, True), (192, 512, 1536, 0, True, False, False),
             (3, 512, 512, 0, True, False, True), (5, 2048, 512, 2, False, True, False), (100, 130, 256, 1, True, False, True),
             (320, 10000, 512, 0, False, True, False), (65, 64, 512, 0, False, True, False), (40, 512, 1536, 0, True, False, False)]
    for (M, N, K, act, res, ln, x16) in cases:
        x = torch.randn(M, K, generator=g) * 1.5 + 0.3
        w = torch.randn(N, K, generator=g) / math.sqrt(K)
        b = torch.randn(N, generator=g)
        r = torch.randn(M, N, generator=g) if res else None
        ga = (1.0 + 0.2 * torch.randn(K, generator=g)) if ln else None
        be = (0.1 * torch.randn(K, generator=g)) if ln else None
        y = e.op_linear_skinny(x, w, b, r, act, precision, ga, be, x16)
        xa = torch.nn.functional.layer_norm(x, (K,), ga, be, 1e-5) if ln else x
        ref = torch.nn.functional.linear(rnd(xa).double(), rnd(w).double(), b.double())
        if act == 1:
            ref = torch.nn.functional.gelu(ref)
        elif act == 2:
            ref = torch.relu(ref)
        if res:
            ref = ref + r.double()
        # the fused LayerNorm rounds to 16 bits from a slightly different fp32 value than torch's: allow 1 operand ulp
        tol = 2e-5 if not ln else (3e-3 if precision == "fp16" else 2e-2)
        out.append((f"linear_{precision}_skinny[{M}x{N}x{K} act{act} res{int(res)} ln{int(ln)} x16={int(x16)}] rel-max", rel_max(y, ref), tol))
    return out


def check_linear_ln_on_load(precision: str = "fp16") -> List[Triple]:
    """tcgen05 GEMM with LayerNorm-on-load (the epilogue warps normalise the fp32 rows into the UMMA shared-memory layout)
    vs the two-kernel path it replaces (LayerNorm kernel -> 16-bit operand -> the same GEMM): bit-identical outputs, plus
    the fp64 reference of the rounded operands."""
    e = bare_engine()
    cast = (lambda t: t.bfloat16()) if precision == "bf16" else (lambda t: t.half())
    out = []
    g = torch.Generator().manual_seed(31)
    for (M, N, act, res, ln) in [(96, 2560, 0, False, True), (192, 2048, 2, False, True), (96, 10000, 0, False, True),
                                 (100, 512, 0, True, True), (1, 512, 0, False, True), (128, 640, 1, True, False), (77, 130, 0, False, True)]:
        K = 512
        x = (torch.randn(M, K, generator=g) * 1.7 + 0.4).cuda()
        w16 = cast((torch.randn(N, K, generator=g) / math.sqrt(K)).cuda())
        b = torch.randn(N, generator=g).cuda()
        r = torch.randn(M, N, generator=g).cuda() if res else None
        ga = (1.0 + 0.2 * torch.randn(K, generator=g)).cuda() if ln else None
        be = (0.1 * torch.randn(K, generator=g)).cuda() if ln else None
        y_fused = e.op_gemm_raw(3, x, w16, b, r, torch.empty(M, N, device="cuda"), act, precision, ga, be)
        a16 = cast(e.op_layernorm(x, ga, be)) if ln else cast(x)
        y_two = e.op_gemm_raw(0, a16, w16, b, r, torch.empty(M, N, device="cuda"), act, precision)
        out.append((f"ln_on_load_{precision}[{M}x{N}x{K} act{act} res{int(res)} ln{int(ln)}] max |fused - (LN kernel, GEMM)|",
                    float((y_fused - y_two).abs().max()), 0.0))
        ref = torch.nn.functional.linear(a16.double(), w16.double(), b.double())
        if act == 1:
            ref = torch.nn.functional.gelu(ref)
        elif act == 2:
            ref = torch.relu(ref)
        if res:
            ref = ref + r.double()
        out.append((f"ln_on_load_{precision}[{M}x{N}x{K}] vs fp64 product of the rounded operands rel-max", rel_max(y_fused, ref), 2e-5))
    return out


def check_window_attention(precision: str = "fp32", kernel: str = "default") -> List[Triple]:
    """kernel: "default" (the engine's choice: tcgen05 where supported in the 16-bit modes), "mma" (mma.sync kernel forced),
    "tc" (alias of default; labels the row).  The larger shapes give every persistent CTA of the tcgen05 kernel several
    items (both smem stages, barrier phase wrap-around) and cover masked and unmasked windows of a shifted block."""
    e = bare_engine()
    out = []
    g = torch.Generator().manual_seed(3)
    if precision != "fp32":
        e.set_option("attn_tc", 0 if kernel == "mma" else 1)
    shapes = [(2, 24, 2, 0), (2, 24, 2, 6), (1, 48, 3, 6), (3, 12, 4, 0)]
    if precision != "fp32":
        shapes += [(4, 48, 6, 6), (2, 24, 24, 6), (5, 12, 48, 0), (7, 48, 6, 0)]
    for (B, H, heads, shift) in shapes:
        C = heads * 32
        qkv = torch.randn(B * H * H, 3 * C, generator=g)
        table = torch.randn(529, heads, generator=g) * 0.5
        y = e.op_window_attention(qkv, table, B, H, C, heads, shift, precision)
        # oracle: same math through the reference-shaped partition/roll path
        q_in = qkv.bfloat16().float() if precision == "bf16" else (qkv.half().float() if precision == "fp16" else qkv)
        x = q_in.reshape(B, H, H, 3 * C)
        if shift:
            x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
        xw = O._to_windows(x, 12)
        Bw = xw.shape[0]
        qkv_w = xw.reshape(Bw, 144, 3, heads, 32).permute(2, 0, 3, 1, 4)
        q, k, v = qkv_w[0] * (32 ** -0.5), qkv_w[1], qkv_w[2]
        att = q @ k.transpose(-2, -1)
        att = att + table[O.relative_position_index(12).reshape(-1)].reshape(144, 144, heads).permute(2, 0, 1).unsqueeze(0)
        if shift:
            lab = O._to_windows(O.shift_region_labels(H, 12, shift).reshape(1, H, H, 1).float(), 12).reshape(-1, 144)
            diff = lab[:, None, :] - lab[:, :, None]
            mask = torch.where(diff != 0, torch.tensor(-100.0), torch.tensor(0.0))
            nW = mask.shape[0]
            att = (att.reshape(Bw // nW, nW, heads, 144, 144) + mask[None, :, None]).reshape(Bw, heads, 144, 144)
        o = (torch.softmax(att, -1) @ v).transpose(1, 2).reshape(Bw, 144, C)
        o = O._from_windows(o, 12, B, H, H)
        if shift:
            o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
        ref = o.reshape(B * H * H, C)
        tol = {"fp32": 1e-5, "bf16": 1e-2, "fp16": 2e-3}[precision]
        out.append((f"window_attention_{precision}/{kernel}[B{B} H{H} heads{heads} shift{shift}] rel-max", rel_max(y, ref), tol))
    if precision != "fp32":
        e.set_option("attn_tc", 1)
    return out


def check_logsoftmax_topk() -> List[Triple]:
    e = bare_engine()
    g = torch.Generator().manual_seed(4)
    out = []
    for rows, V, k in [(6, 10000, 3), (17, 10000, 5), (4, 512, 8), (3, 777, 1)]:
        x = torch.randn(rows, V, generator=g) * 2
        tv, ti, lp = e.op_logsoftmax_topk(x, k, want_logprob=True)
        ref = torch.log_softmax(x, -1)
        rv, ri = torch.topk(ref, k, sorted=True)
        out.append((f"logsoftmax[{rows}x{V}] max-abs", float((lp.cpu() - ref).abs().max()), 4e-6))
        out.append((f"topk[{rows}x{V} k{k}] index mismatches", float((ti.cpu().long() != ri).sum()), 0.0))
        out.append((f"topk[{rows}x{V} k{k}] value max-abs", float((tv.cpu() - rv).abs().max()), 4e-6))
    return out


# ------------------------------------------------------------------ model-level vs oracle + golden
def check_encoder(name: str, precision: str = "fp32") -> List[Triple]:
    e, g, cfg, sd, x, pads = engine_for(name, precision)
    out = []
    with torch.no_grad():
        taps = {}
        ref = O.forward_enc(sd, cfg, x, pads, taps)
    # bf16 operands cannot reach the 2e-3 target through 24 Swin blocks (measured 6e-3): it is held
    # to 1e-2; fp16 operands (same tensor-core rate, 3 more mantissa bits) meet 2e-3 and are the default 16-bit mode
    tol_feat = {"fp32": 2e-5, "fp16": 2e-3, "bf16": 1e-2}[precision]
    fro = precision != "fp32"
    m = rel_fro if fro else rel_max
    kind = "rel-fro" if fro else "rel-max"
    if cfg.has_swin:
        sw = e.forward_swin(x)
        out.append((f"{name}/{precision} swin features vs oracle {kind}", m(sw, taps["swin"]), tol_feat))
        if precision == "fp32":
            out.append((f"{name}/{precision} swin features vs golden(reference) max-abs",
                        float(np.abs(sub(sw.cpu()).numpy() - g["swin_sub"]).max()), 5e-5))
    enc = e.forward_enc(x, pads)
    out.append((f"{name}/{precision} encoder output vs oracle {kind}", m(enc, ref), tol_feat))
    if precision == "fp32":
        out.append((f"{name}/{precision} encoder output vs golden(reference) max-abs",
                    float(np.abs(sub(enc.cpu()).numpy() - g["enc_sub"]).max()), 2e-5))
    return out


def check_decoder(name: str, precision: str = "fp32") -> List[Triple]:
    e, g, cfg, sd, x, pads = engine_for(name, precision)
    out = []
    with torch.no_grad():
        enc = O.forward_enc(sd, cfg, x, pads)          # oracle encoder output as the common cross input
        tok = torch.from_numpy(g["dec_tokens"])
        dp = g["dec_pads"].tolist()
        ref_lp = O.forward_dec(sd, cfg, enc, pads, tok, dp, True)
        ref_lg = O.forward_dec(sd, cfg, enc, pads, tok, dp, False)
    lg = e.forward_dec(enc, pads, tok, dp, False)
    lp = e.forward_dec(enc, pads, tok, dp, True)
    # bf16 operands (8 significand bits) are an optional, coarser mode: 5e-2 on the sharpened "peaky" vocabulary head
    tol = {"fp32": 1e-5, "fp16": 2e-3, "bf16": 5e-2}[precision]
    out.append((f"{name}/{precision} teacher-forced logits vs oracle rel-max (incl. padded rows)", rel_max(lg, ref_lg), tol))
    out.append((f"{name}/{precision} teacher-forced log-probs vs oracle rel-max", rel_max(lp, ref_lp), tol))
    if precision == "fp32":
        out.append((f"{name}/{precision} logits vs golden(reference) max-abs",
                    float(np.abs(sub(lg.cpu()).numpy() - g["dec_logits_sub"]).max()), 2e-5))
        ti = torch.topk(lp.cpu(), 8, dim=-1).indices.numpy()
        out.append((f"{name}/{precision} top-8 index mismatches vs golden(reference)", float((ti != g["dec_top8_idx"]).sum()), 0.0))
    return out


def check_beam(name: str, precision: str = "fp32") -> List[Triple]:
    e, g, cfg, sd, x, pads = engine_for(name, precision)
    m = g["meta"]
    tok, ln, lp = e.beam_search(x, pads, m["sos"], m["eos"], m["beam"], m["how_many"], m["max_len"])
    toks, lps = unpack_beam_results(tok, ln, lp)
    out = []
    margin = np.minimum(np.minimum(g["vocab_margin"], g["merge_margin"]), g["final_margin"])
    bad_tok = 0
    checked = 0
    for b in range(m["B"]):
        if precision == "fp32" and margin[b] <= 2e-5:
            continue                                   # decision closer than the fp32 tolerance: not a parity claim
        checked += 1
        for j in range(m["how_many"]):
            n = int(g["beam_len"][b, j])
            if toks[b][j] != g["beam_tokens"][b, j, :n].tolist():
                bad_tok += 1
    if precision == "fp32":
        out.append((f"{name}/fp32 caption token sequences differing from the reference ({checked}/{m['B']} images with margin>2e-5)",
                    float(bad_tok), 0.0))
        ref_lp = torch.from_numpy(g["beam_logprobs"])
        if bad_tok == 0 and checked == m["B"] and tuple(lps.shape) == tuple(ref_lp.shape):
            out.append((f"{name}/fp32 caption log-probs vs reference rel-max", rel_max(lps, ref_lp), 1e-5))
    else:
        out.append((f"{name}/{precision} caption token sequences differing from the reference (informational)", float(bad_tok), float("inf")))
    return out


def check_preprocess() -> List[Triple]:
    """GPU preprocessing (Pillow-exact resize + ToTensor + Normalize) vs the CPU oracle: bit-exact float32 tensors,
    host and device inputs, down- and up-scaling, extreme aspect ratios, 1-pixel images."""
    from oracle import preprocess_oracle as P
    from test_preprocess_oracle import synth_image
    e = bare_engine()
    out = []
    for (H, W, S) in [(50, 70, 96), (500, 333, 384), (384, 384, 384), (640, 480, 384), (100, 1000, 384), (37, 41, 48),
                      (1200, 800, 384), (383, 385, 384), (1, 1, 16), (2, 900, 32), (2160, 3840, 384)]:
        img = synth_image(H, W, H * 1000 + W)
        ref = P.preprocess_rgb8(img, S)
        y_host = e.preprocess_rgb8([img], S)[0].cpu().numpy()
        y_dev = e.preprocess_rgb8([torch.from_numpy(img).cuda()], S)[0].cpu().numpy()
        out.append((f"preprocess[{H}x{W}->{S}] elements differing from the oracle (host input)", float((y_host != ref).sum()), 0.0))
        out.append((f"preprocess[{H}x{W}->{S}] elements differing from the oracle (device input)", float((y_dev != ref).sum()), 0.0))
    # the file-level mirror of utils/image_utils.py:preprocess_image: PNG files (lossless), including the reference's rule
    # that a non-RGB file is replaced by a blank RGB canvas of the same size (:18-19)
    import tempfile
    from PIL import Image
    from on_device_image_captioning_b200.image_utils import preprocess_image
    from test_preprocess_oracle import reference_preprocess
    with tempfile.TemporaryDirectory() as td:
        rgb = synth_image(211, 333, 5)
        Image.fromarray(rgb, "RGB").save(td + "/a.png")
        Image.fromarray(rgb[..., 0], "L").save(td + "/gray.png")
        y = preprocess_image(td + "/a.png", 96, e)[0].cpu().numpy()
        out.append(("preprocess_image(RGB png) elements differing from Pillow + torchvision", float((y != reference_preprocess(rgb, 96)).sum()), 0.0))
        yg = preprocess_image(td + "/gray.png", 96, e)[0].cpu().numpy()
        blank = reference_preprocess(np.zeros((211, 333, 3), dtype=np.uint8), 96)
        out.append(("preprocess_image(grayscale png -> blank canvas rule) elements differing", float((yg != blank).sum()), 0.0))
    imgs = [synth_image(h, w, 7 * h + w) for (h, w) in [(300, 400), (480, 640), (200, 200)]]
    yb = e.preprocess_rgb8(imgs, 96).cpu().numpy()
    bad = sum(int((yb[i] != P.preprocess_rgb8(im, 96)).sum()) for i, im in enumerate(imgs))
    out.append(("preprocess batch of mixed sizes: elements differing from the oracle", float(bad), 0.0))
    return out


def check_feature_extraction() -> List[Triple]:
    """the bulk extractor (decode -> GPU preprocess -> Swin, batched) writes, per image id, exactly the
    features a one-image call produces, in the reference's "<img_id>_features" layout."""
    import tempfile
    from on_device_image_captioning_b200 import features as F
    from test_preprocess_oracle import synth_image
    e, g, cfg, sd, x, pads = engine_for("tiny_e2e_peaky", "fp32")
    imgs = [synth_image(h, w, 3 * h + w) for (h, w) in [(120, 160), (96, 96), (300, 200), (64, 500), (97, 101)]]
    ids = [11, 22, 33, 44, 55]
    with tempfile.TemporaryDirectory() as td:
        path = td + "/precalc_features.hdf5"
        F.extract_features(e, imgs, ids, path, batch_size=2)
        bad = 0.0
        for im, i in zip(imgs, ids):
            one = e.forward_swin(e.preprocess_rgb8([im]))[0].cpu().numpy()
            got = F.read_features(path, i)
            bad += float(got.shape != (cfg.enc_len, cfg.feat_dim)) + float(np.abs(got - one).max() > 1e-6)
    return [("feature extractor: images whose stored features differ from the single-image call", bad, 0.0)]


def check_ensemble() -> List[Triple]:
    """ensemble beam search (two tiny end-to-end models) vs the fixture produced by the reference's
    EsembleCaptioningModel: bit-exact captions in fp32, through the raw engine call and the drop-in class."""
    from conftest import load_golden
    from on_device_image_captioning_b200 import synth
    from on_device_image_captioning_b200.models import End_ExpansionNet_v2, EsembleCaptioningModel
    g = load_golden("ens_tiny_e2e")
    m = g["meta"]
    cfg = XNConfig(**m["cfg"])
    sds = [synth.make_state_dict(cfg, seed=sd_, profile=m["profile"], eos_idx=m["eos"]) for sd_ in m["seeds"]]
    x = synth.make_images(cfg, m["B"], seed=1, kind=m["kind"])
    out = []
    margin = np.minimum(np.minimum(g["vocab_margin"], g["merge_margin"]), g["final_margin"])
    for precision in ("fp32", "fp16"):
        engs = []
        for sd_ in sds:
            e = Engine(cfg, 0)
            e.load_state_dict(sd_, precision)
            engs.append(e)
        tok, ln, lp = Engine.ensemble_beam_search(engs, x, None, m["sos"], m["eos"], m["beam"], m["how_many"], m["max_len"])
        toks, lps = unpack_beam_results(tok, ln, lp)
        bad = 0
        for b in range(m["B"]):
            if margin[b] <= 2e-5:
                continue
            for j in range(m["how_many"]):
                bad += toks[b][j] != g["beam_tokens"][b, j, : int(g["beam_len"][b, j])].tolist()
        if precision == "fp32":
            out.append(("ensemble/fp32 caption token sequences differing from the reference class", float(bad), 0.0))
            ref_lp = torch.from_numpy(g["beam_logprobs"])
            if bad == 0 and tuple(lps.shape) == tuple(ref_lp.shape):
                out.append(("ensemble/fp32 caption log-probs vs reference rel-max", rel_max(lps, ref_lp), 1e-5))
        else:
            out.append((f"ensemble/{precision} caption token sequences differing from the reference (informational)", float(bad), float("inf")))
        for e in engs:
            e.close()
    # drop-in class surface (test.py:334 style)
    import argparse
    words = [f"w{i}" for i in range(cfg.vocab)]
    da = argparse.Namespace(enc=0.0, dec=0.0, enc_input=0.0, dec_input=0.0, other=0.0)
    models = []
    for sd_ in sds:
        mm = End_ExpansionNet_v2(swin_img_size=cfg.img_size, swin_patch_size=cfg.patch_size, swin_in_chans=cfg.in_chans,
                                 swin_embed_dim=cfg.embed_dim, swin_depths=list(cfg.depths), swin_num_heads=list(cfg.swin_heads),
                                 swin_window_size=cfg.window_size, swin_mlp_ratio=cfg.mlp_ratio, swin_qkv_bias=True, swin_qk_scale=None,
                                 swin_drop_rate=0.0, swin_attn_drop_rate=0.0, swin_drop_path_rate=0.0, swin_norm_layer=torch.nn.LayerNorm,
                                 swin_ape=False, swin_patch_norm=True, swin_use_checkpoint=False, final_swin_dim=cfg.feat_dim,
                                 d_model=cfg.d_model, N_enc=cfg.n_enc, N_dec=cfg.n_dec, ff=cfg.ff, num_heads=cfg.num_heads,
                                 num_exp_enc_list=list(cfg.num_exp_enc_list), num_exp_dec=cfg.num_exp_dec,
                                 output_word2idx={w: i for i, w in enumerate(words)}, output_idx2word=words,
                                 max_seq_len=cfg.max_seq_len, drop_args=da, rank=0, precision="fp32")
        mm.load_state_dict(sd_)
        models.append(mm.to(0).eval())
    ens = EsembleCaptioningModel(models, 0)
    with torch.no_grad():
        pred, plp = ens(enc_x=x.cuda(), enc_x_num_pads=[0] * m["B"], mode="beam_search", beam_size=m["beam"],
                        how_many_outputs=m["how_many"], beam_max_seq_len=m["max_len"], sample_or_max="max",
                        sos_idx=m["sos"], eos_idx=m["eos"])
    bad = sum(pred[b][j] != g["beam_tokens"][b, j, : int(g["beam_len"][b, j])].tolist()
              for b in range(m["B"]) if margin[b] > 2e-5 for j in range(m["how_many"]))
    out.append(("ensemble drop-in class: caption token sequences differing from the reference class", float(bad), 0.0))
    return out


def check_caption_host() -> List[Triple]:
    """xn_caption_host (host buffers in, host tokens out -- the call bench.py's e2e times): pinned input takes the in-graph
    chunked-copy path (Swin chunks of 32 overlapping the copies), pageable input the plain copy; both must return exactly
    what the device-input beam search returns, on repeated calls (eager, capture, replay) and for a batch that is not a
    multiple of the copy chunk."""
    e, g, cfg, sd, x, pads = engine_for("full_e2e_peaky", "fp16")
    from on_device_image_captioning_b200 import synth
    m = g["meta"]
    out = []
    for B in (70, 3):
        xs = synth.make_images(cfg, B, seed=21, kind="mixed")
        t_ref, l_ref, _ = e.beam_search(xs, None, m["sos"], m["eos"], 3, 1, 20)
        ref = _tokens_list(t_ref, l_ref)
        for kind, host in (("pinned", xs.clone().pin_memory()), ("pageable", xs.clone())):
            bad = 0
            for _ in range(4):                      # first sight, capture, two replays
                tok, ln, lp = e.caption_host(host, m["sos"], m["eos"], 3, 1, 20)
                bad += sum(1 for i, t in enumerate(_tokens_list(tok, ln)) if t != ref[i])
            out.append((f"caption_host[{kind} input, B={B}] captions differing from the device-input call (4 calls)", float(bad), 0.0))
    return out


def check_engine_pair() -> List[Triple]:
    """EnginePair (two handles taking alternate pipelined host-buffer calls, up to four in flight): every call must return
    exactly what the single handle's device-input beam search returns for that batch, over eager / capture / replay of
    both handles and with tickets ended out of order."""
    from on_device_image_captioning_b200 import synth
    from on_device_image_captioning_b200.engine import EnginePair
    e, g, cfg, sd, x, pads = engine_for("full_e2e_peaky", "fp16")
    m = g["meta"]
    B, L = 5, 20
    hosts = [synth.make_images(cfg, B, seed=40 + i, kind="mixed").pin_memory() for i in range(3)]
    refs = []
    for hst in hosts:
        t_ref, l_ref, _ = e.beam_search(hst, None, m["sos"], m["eos"], 3, 1, L)
        refs.append(_tokens_list(t_ref, l_ref))
    pair = EnginePair(cfg, 0)
    pair.load_state_dict(sd, "fp16")
    outs = [(torch.empty(B, 1, L, dtype=torch.int32).pin_memory(), torch.empty(B, 1, dtype=torch.int32).pin_memory(),
             torch.empty(B, 1, L, dtype=torch.float32).pin_memory()) for _ in range(4)]
    bad, n_calls, q = 0, 14, []

    def end(idx):
        nonlocal bad
        i, t = q.pop(idx)
        pair.caption_host_end(t)
        tok, ln, _ = outs[i % 4]
        bad += sum(1 for b, tl in enumerate(_tokens_list(tok, ln)) if tl != refs[i % 3][b])

    for i in range(n_calls):
        q.append((i, pair.caption_host_begin(hosts[i % 3], m["sos"], m["eos"], 3, 1, L, outs[i % 4])))
        if len(q) == 4:
            end(1 if i % 5 == 0 else 0)          # now and then a younger ticket first (the other handle's)
            end(0)
    while q:
        end(0)
    err = 0.0
    try:
        pair.caption_host_end(12345)
        err = 1.0
    except RuntimeError:
        pass
    pair.close()
    return [(f"EnginePair: captions differing from the single handle over {n_calls} pipelined calls", float(bad), 0.0),
            ("EnginePair: unknown ticket accepted", err, 0.0)]


# ------------------------------------------------------------------ BASELINE.json configurations at full size
def _tokens_list(tok, ln):
    tok, ln = tok.cpu(), ln.cpu()
    return [tok[b, 0, : int(ln[b, 0])].tolist() for b in range(tok.shape[0])]


def check_config3_features_beam5(B: int = 256) -> List[Triple]:
    """BASELINE.json configs[2]: decoder-only ExpansionNet_v2 on precomputed (144 x 1536) features, batch 256, beam 5.
    fp32: the first images are decoded by the CPU oracle (seconds) and must match token for token where the oracle's
    own decision margins allow a parity claim; the rest of the batch is held to batch invariance (same captions whether
    an image is decoded inside the batch of 256 or in a batch of 8).  fp16: invariance only."""
    from on_device_image_captioning_b200 import synth
    from on_device_image_captioning_b200.config import features_only
    cfg = features_only(vocab=1000, max_seq_len=24)
    sd = synth.make_state_dict(cfg, seed=0, profile="peaky", eos_idx=7)
    x = synth.make_features(cfg, B, seed=3)
    out = []
    n_or = 3
    with torch.no_grad():
        tr = {}
        ref_tok, ref_lp = O.beam_search(sd, cfg, x[:n_or], [0] * n_or, 5, 7, 5, 1, 20, trace=tr)
        ref_m = torch.minimum(torch.minimum(tr["vocab_margin"], tr["merge_margin"]), tr["final_margin"]).tolist()
    for precision in ("fp32", "fp16"):
        e = Engine(cfg, 0)
        e.load_state_dict(sd, precision)
        tok, ln, lp = e.beam_search(x, [0] * B, 5, 7, 5, 1, 20)
        full = _tokens_list(tok, ln)
        diff = 0
        for b0 in range(0, B, 64):                      # a few sub-batches spread over the batch
            t8, l8, _ = e.beam_search(x[b0:b0 + 8].contiguous(), [0] * 8, 5, 7, 5, 1, 20)
            diff += sum(1 for i, t in enumerate(_tokens_list(t8, l8)) if t != full[b0 + i])
        out.append((f"config3/{precision} B={B} beam 5: captions that change with the batch they are decoded in", float(diff), 0.0))
        if precision == "fp32":
            bad = sum(1 for i in range(n_or) if ref_m[i] > 2e-5 and full[i] != ref_tok[i][0])
            out.append((f"config3/fp32 captions differing from the CPU oracle (first {n_or} images)", float(bad), 0.0))
        lens = ln.cpu().flatten().tolist()
        out.append((f"config3/{precision} caption lengths outside [2, 20]", float(sum(1 for v in lens if v < 2 or v > 20)), 0.0))
        e.close()
    return out


def check_config4_batch512_chunking() -> List[Triple]:
    """BASELINE.json configs[3] per-GPU shape: 512 images in one call (8 Swin chunks of 64, 1536 decoder rows).  The
    batch is 8 copies of 64 distinct images, so every chunk must reproduce the captions of a plain 64-image call."""
    e, g, cfg, sd, x, pads = engine_for("full_e2e_peaky", "fp16")
    from on_device_image_captioning_b200 import synth
    m = g["meta"]
    x64 = synth.make_images(cfg, 64, seed=11, kind="mixed")
    t64, l64, _ = e.beam_search(x64, None, m["sos"], m["eos"], 3, 1, 20)
    base = _tokens_list(t64, l64)
    x512 = x64.repeat(8, 1, 1, 1).contiguous()
    t, l, _ = e.beam_search(x512, None, m["sos"], m["eos"], 3, 1, 20)
    big = _tokens_list(t, l)
    diff = sum(1 for i in range(512) if big[i] != base[i % 64])
    distinct = len({tuple(c) for c in base})
    return [("config4/fp16 B=512: captions differing from the 64-image call they repeat", float(diff), 0.0),
            ("config4/fp16 distinct captions among the 64 images (>= 8 expected, the check is not vacuous)", float(-distinct), -8.0),
            ("config4/fp16 workspace GiB (of 180)", e.workspace_bytes / 2 ** 30, 60.0)]


# ------------------------------------------------------------------ round-2 parity rows
def check_config2_batch64(name: str, precision: str = "fp32") -> List[Triple]:
    """BASELINE.json configs[1] at its full batch (64 images, beam 3, max_len 20) against the fixture the UNMODIFIED
    reference produced in one batch-64 call (tests/golden/make_golden.py c2_b64_*; the oracle was asserted identical to
    it at generation time).  fp32: captions bit-exact on every image whose recorded decision margin exceeds 2e-5, log-probs
    and sub-sampled encoder output within the fp32 tolerance.  16-bit: the same comparisons, reported (rel-max)."""
    e, g, cfg, sd, x, pads = engine_for(name, precision)
    m = g["meta"]
    out = []
    enc = e.forward_enc(x, None)
    ref_enc = torch.from_numpy(g["enc_sub"])
    out.append((f"{name}/{precision} B=64 encoder output vs reference fixture rel-max", rel_max(sub(enc.cpu()), ref_enc),
                {"fp32": 2e-5, "fp16": 4e-3, "bf16": 3e-2}[precision]))
    tok, ln, lp = e.beam_search(x, None, m["sos"], m["eos"], m["beam"], m["how_many"], m["max_len"])
    toks, lps = unpack_beam_results(tok, ln, lp)
    margin = np.minimum(np.minimum(g["vocab_margin"], g["merge_margin"]), g["final_margin"])
    checked = [b for b in range(m["B"]) if margin[b] > 2e-5]
    bad = sum(1 for b in checked if toks[b][0] != g["beam_tokens"][b, 0, : int(g["beam_len"][b, 0])].tolist())
    if precision == "fp32":
        out.append((f"{name}/fp32 B=64 captions differing from the reference ({len(checked)}/{m['B']} images with margin>2e-5)", float(bad), 0.0))
        ref_lp = torch.from_numpy(g["beam_logprobs"])
        same = [b for b in checked if toks[b][0] == g["beam_tokens"][b, 0, : int(g["beam_len"][b, 0])].tolist()]
        if same and tuple(lps.shape) == tuple(ref_lp.shape):
            out.append((f"{name}/fp32 B=64 caption log-probs vs reference rel-max", rel_max(lps.cpu()[same], ref_lp[same]), 1e-5))
    else:
        out.append((f"{name}/{precision} B=64 captions differing from the reference fp32 run, of {len(checked)} (informational)",
                    float(bad), float("inf")))
    distinct = len({tuple(t[0]) for t in toks})
    out.append((f"{name}/{precision} distinct captions in the batch (informational)", float(distinct), float("inf")))
    return out


def check_image_to_logits_16bit(name: str, precision: str = "fp16") -> List[Triple]:
    """The 16-bit path END TO END: image -> Swin -> expansion encoder -> teacher-forced decoder, everything on the GPU in
    the 16-bit mode (the decoder consumes the GPU's own encoder output, not the oracle's), against the fp32 oracle.
    Measured as rel-max = max|a-b| / max|b| -- the same metric as the fp32 rows.  north_star states 2e-3 relative for the
    16-bit mode: fp16 is held to it on logits; the feature rows are held to the measured bound."""
    e, g, cfg, sd, x, pads = engine_for(name, precision)
    out = []
    with torch.no_grad():
        taps = {}
        ref_enc = O.forward_enc(sd, cfg, x, pads, taps)
        tok = torch.from_numpy(g["dec_tokens"])
        dp = g["dec_pads"].tolist()
        ref_lg = O.forward_dec(sd, cfg, ref_enc, pads, tok, dp, False)
        ref_lp = O.forward_dec(sd, cfg, ref_enc, pads, tok, dp, True)
    tol = {"fp16": (4e-3, 4e-3, 2e-3), "bf16": (3e-2, 3e-2, 3e-2)}[precision]
    if cfg.has_swin:
        sw = e.forward_swin(x)
        out.append((f"{name}/{precision} swin features vs oracle rel-max", rel_max(sw, taps["swin"]), tol[0]))
        out.append((f"{name}/{precision} swin features vs oracle rel-fro", rel_fro(sw, taps["swin"]), tol[0]))
    enc = e.forward_enc(x, pads)
    out.append((f"{name}/{precision} encoder output (GPU end to end) vs oracle rel-max", rel_max(enc, ref_enc), tol[1]))
    lg = e.forward_dec(enc, pads, tok, dp, False)
    lp = e.forward_dec(enc, pads, tok, dp, True)
    out.append((f"{name}/{precision} image->logits (GPU encoder output into GPU decoder) vs oracle rel-max", rel_max(lg, ref_lg), tol[2]))
    out.append((f"{name}/{precision} image->log-probs vs oracle rel-max", rel_max(lp, ref_lp), tol[2]))
    n = tok.shape[0] * tok.shape[1]
    top1 = float((lg.argmax(-1).cpu() != ref_lg.argmax(-1)).sum())
    out.append((f"{name}/{precision} top-1 word differs from the oracle at N of {n} positions (informational)", top1, float("inf")))
    out.append((f"{name}/{precision} overflow flag", float(e.overflow_flag()), 0.0))
    return out


def check_tensor_core_options(name: str = "full_e2e_xavier", precision: str = "fp16") -> List[Triple]:
    """Round-2 tensor-core variants against the kernels they replace, same weights and images: the static-expansion
    scores / class / out contractions on tcgen05 (option se_tc = 2 vs 0: mma.sync + (B,E,N)-layout kernels) and the
    TF32 tensor-core patch embedding (pe_tc = 1 vs 0: CUDA-core fp32).  Both variants must meet the 16-bit tolerance
    against the fp32 oracle; the difference between them is reported."""
    e, g, cfg, sd, x, pads = engine_for(name, precision)
    out = []
    with torch.no_grad():
        taps = {}
        ref = O.forward_enc(sd, cfg, x, pads, taps)
    res = {}
    for tag, se, pe in (("se_tc=0,pe_tc=0", 0, 0), ("se_tc=1,pe_tc=0", 1, 0), ("se_tc=2,pe_tc=1", 2, 1)):
        e.set_option("se_tc", se)
        e.set_option("pe_tc", pe)
        if cfg.has_swin:
            sw = e.forward_swin(x).clone()
            out.append((f"{name}/{precision} [{tag}] swin features vs oracle rel-max", rel_max(sw, taps["swin"]), 4e-3))
        enc = e.forward_enc(x, pads).clone()
        res[tag] = enc
        out.append((f"{name}/{precision} [{tag}] encoder output vs oracle rel-max", rel_max(enc, ref), 4e-3))
    out.append((f"{name}/{precision} encoder output, tensor-core options on vs off, rel-max", rel_max(res["se_tc=2,pe_tc=1"], res["se_tc=0,pe_tc=0"]), 4e-3))
    e.set_option("se_tc", 2)
    e.set_option("pe_tc", 1)
    return out


def check_decoder_splitk(name: str = "full_e2e_peaky", precision: str = "fp16") -> List[Triple]:
    """Long-K decoder-step linears (ff2, reduce group) as K slices on the batched tcgen05 GEMM, summed by the LayerNorm
    launch that follows (option dec_splitk), against the single-launch form: teacher-forced logits of both vs the fp32
    oracle, and the beam-search captions of both against each other."""
    e, g, cfg, sd, x, pads = engine_for(name, precision)
    out = []
    with torch.no_grad():
        enc = O.forward_enc(sd, cfg, x, pads)
        # enough rows for the tcgen05 path (the skinny kernel takes <= 64 rows): repeat the fixture's sequences
        reps = max(1, 96 // enc.shape[0] + 1)
        enc_r = enc.repeat(reps, 1, 1)
        pads_r = (list(pads) * reps) if pads is not None else None
        tok = torch.from_numpy(g["dec_tokens"]).repeat(reps, 1)
        dp = g["dec_pads"].tolist() * reps
        ref_lg = O.forward_dec(sd, cfg, enc, pads, torch.from_numpy(g["dec_tokens"]), g["dec_pads"].tolist(), False).repeat(reps, 1, 1)
    res = {}
    for v in (0, 1):
        e.set_option("dec_splitk", v)
        lg = e.forward_dec(enc_r, pads_r, tok, dp, False).clone()
        res[v] = lg
        out.append((f"{name}/{precision} [dec_splitk={v}] teacher-forced logits ({lg.shape[0]} rows) vs oracle rel-max", rel_max(lg, ref_lg), 2e-3))
    # not bit-equal: fp32 sums in a different order flip a few 16-bit roundings of the next layer's operand (measured 2.4e-4)
    out.append((f"{name}/{precision} logits split-K vs single launch rel-max", rel_max(res[1], res[0]), 1e-3))
    e.set_option("dec_splitk", 0)
    return out


def check_fp16_saturation() -> List[Triple]:
    """fp16 stores QKV / attention output / MLP hidden as halves (max 65504).  Scale the fc1 weights and bias of one
    stage-1 block (the hidden activations, hence fc2's input) by 300: fp16 must survive (finite, and as close to the
    fp32 oracle as without the scaling).  Scale by 3e5: the hidden overflows, the device-side check of the encoder output
    must raise the overflow flag, the drop-in class must refuse with a clear error, and bf16 (fp32 range) must be finite."""
    from on_device_image_captioning_b200 import synth
    from on_device_image_captioning_b200.config import swin_tiny_test
    cfg = swin_tiny_test()
    x = synth.make_images(cfg, 2, seed=1, kind="mixed")
    out = []
    for scale, expect_flag in ((300.0, False), (3e5, True)):
        sd = synth.make_state_dict(cfg, 0, "peaky", eos_idx=77)
        for k in ("swin_transf.layers.0.blocks.1.mlp.fc1.weight", "swin_transf.layers.0.blocks.1.mlp.fc1.bias"):
            sd[k] = sd[k] * scale
        with torch.no_grad():
            ref = O.forward_enc(sd, cfg, x, [0, 0])
        e = Engine(cfg, 0)
        e.load_state_dict(sd, "fp16")
        e.overflow_flag()
        enc = e.forward_enc(x, None)
        flag = e.overflow_flag()
        finite = bool(torch.isfinite(enc).all())
        out.append((f"saturation x{scale:g}: fp16 overflow flag == {expect_flag}", float(flag != expect_flag), 0.0))
        out.append((f"saturation x{scale:g}: fp16 encoder output finite == {not expect_flag}", float(finite == expect_flag), 0.0))
        if not expect_flag:
            out.append((f"saturation x{scale:g}: fp16 encoder output vs fp32 oracle rel-max", rel_max(enc, ref), 4e-3))
        else:
            e.load_state_dict(sd, "bf16")
            encb = e.forward_enc(x, None)
            out.append((f"saturation x{scale:g}: bf16 (fp32 exponent range) encoder output non-finite values",
                        float((~torch.isfinite(encb)).sum()), 0.0))
            out.append((f"saturation x{scale:g}: bf16 overflow flag", float(e.overflow_flag()), 0.0))
            out.append((f"saturation x{scale:g}: bf16 encoder output vs fp32 oracle rel-max", rel_max(encb, ref), 5e-2))
        e.close()
    return out


def check_demo_known_answers() -> List[Triple]:
    """BASELINE.json configs[0]: the reference's four demo images (demo.py:106-129), fp32, batch 1 each, beam 3,
    max_len 20, vocabulary of demo_coco_tokens.pickle -- against tests/golden/demo_c1.npz, made by the unmodified reference
    (tests/golden/make_demo_golden.py).  Pixels: the 384x384 RGB8 image after the reference's resize goes through
    xn_preprocess_rgb8 (ToTensor + Normalize; 384 -> 384 is the identity resample); for the two committed JPEG files the
    whole chain file -> PIL decode -> GPU resize -> caption is run as well.  Captions must be bit-exact where the decision
    margin exceeds 2e-5 (napoleon-like 1e-6 margins are reported, not asserted)."""
    import os
    from conftest import load_golden, GOLDEN_DIR
    from on_device_image_captioning_b200 import synth
    from on_device_image_captioning_b200.image_utils import preprocess_image
    from on_device_image_captioning_b200.language_utils import tokens2description
    g = load_golden("demo_c1")
    m = g["meta"]
    cfg = XNConfig(**m["cfg"])
    sd = synth.make_state_dict(cfg, seed=0, profile=m["profile"], eos_idx=m["eos"])
    fp = sum(float(sd[k].double().abs().sum()) for k in sorted(sd))
    assert abs(fp - m["weight_fingerprint"]) <= 1e-9 * abs(fp)
    for k in [k for k in _engines if k != "bare"]:
        _engines.pop(k)[0].close()
    e = Engine(cfg, 0)
    e.load_state_dict(sd, "fp32")
    i2w = {int(k): v for k, v in m["used_words"].items()}
    out = []
    for n, name in enumerate(m["images"]):
        st = g[f"stats_{n}"]
        margin = float(min(st[3], st[4], st[5]))
        x = e.preprocess_rgb8([g[f"u8_{n}"]], cfg.img_size)
        out.append((f"demo/{name} mean of the preprocessed tensor vs reference abs diff", abs(float(x.double().mean()) - float(st[0])), 1e-6))
        variants = [("resized pixels", x)]
        path = os.path.join(GOLDEN_DIR, "demo_material", name)
        if os.path.exists(path):
            xf = preprocess_image(path, cfg.img_size, e)
            out.append((f"demo/{name} file -> PIL decode -> GPU resize: elements differing from the reference tensor",
                        float((xf != x).sum()), 0.0))
            variants.append(("file", xf))
        for what, xin in variants:
            enc = e.forward_enc(xin, None)
            out.append((f"demo/{name} [{what}] mean |encoder output| vs reference rel", abs(float(enc.double().abs().mean()) / float(st[1]) - 1.0), 1e-5))
            tok, ln, lp = e.beam_search(xin, None, m["sos"], m["eos"], m["beam"], 1, m["max_len"])
            toks, lps = unpack_beam_results(tok, ln, lp)
            same = toks[0][0] == g[f"tokens_{n}"].tolist()
            if margin > 2e-5:
                out.append((f"demo/{name} [{what}] caption tokens differ from the reference (margin {margin:.1e})", float(not same), 0.0))
            else:
                out.append((f"demo/{name} [{what}] caption tokens differ from the reference (margin {margin:.1e} <= 2e-5: reported only)",
                            float(not same), float("inf")))
            if same:
                out.append((f"demo/{name} [{what}] sum of log-probs vs reference abs diff", abs(float(lps.double().sum()) - float(st[2])), 2e-3))
                cap = tokens2description(toks[0][0], _SparseVocab(i2w), m["sos"], m["eos"])
                out.append((f"demo/{name} [{what}] caption string differs from the reference's tokens2description", float(cap != m["captions"][n]), 0.0))
    e.close()
    return out


class _SparseVocab:
    """idx2word lookup restricted to the words the fixture's captions use (the 10 000-word pickle is reference data)."""

    def __init__(self, d):
        self.d = d

    def __getitem__(self, i):
        return self.d[int(i)]


def check_graph_pointer_independence() -> List[Triple]:
    """A caller that passes a FRESH device tensor (new address) for every batch must hit the captured CUDA graph: the
    input is staged into a handle-owned buffer first.  Five distinct tensors of the same shape: identical captions to a
    graph-free run, and from the third call on one graph launch replays the call (the per-call launch count stays the
    count of the captured graph; no new capture, no eager fallback).  Same for five distinct pinned host buffers through
    xn_caption_host (copy nodes patched in place)."""
    e, g, cfg, sd, x, pads = engine_for("tiny_e2e_peaky", "fp16")
    from on_device_image_captioning_b200 import synth
    m = g["meta"]
    xs = [synth.make_images(cfg, 6, seed=50 + i, kind="mixed") for i in range(5)]
    e.set_option("use_graph", 0)
    ref = [_tokens_list(*e.beam_search(t, None, m["sos"], m["eos"], 3, 1, 12)[:2]) for t in xs]
    e.set_option("use_graph", 1)
    bad, keep, counts = 0, [], []
    for rep in range(2):
        for i, t in enumerate(xs):
            d = t.cuda().clone()                  # a new allocation every call; kept alive so addresses never repeat
            keep.append(d)
            l0 = e.kernel_launches
            tok, ln, _ = e.beam_search(d, None, m["sos"], m["eos"], 3, 1, 12)
            counts.append(e.kernel_launches - l0)
            bad += sum(1 for a, b in zip(_tokens_list(tok, ln), ref[i]) if a != b)
    out = [("graph: captions with 10 distinct device tensors differing from the graph-free run", float(bad), 0.0),
           ("graph: distinct input addresses used", float(-len({d.data_ptr() for d in keep})), -10.0),
           ("graph: per-call launch counts of the replays differing from the captured call's (call 1 eager, call 2 captured, then replays; "
            "the captured graph has a few extra set-condition kernels of the early-exit IF nodes)", float(len(set(counts[1:])) - 1), 0.0)]
    hosts = [t.clone().pin_memory() for t in xs]
    badh = 0
    for rep in range(2):
        for i, hbuf in enumerate(hosts):
            tok, ln, _ = e.caption_host(hbuf, m["sos"], m["eos"], 3, 1, 12)
            badh += sum(1 for a, b in zip(_tokens_list(tok, ln), ref[i]) if a != b)
    out.append(("graph: captions through xn_caption_host with 5 distinct pinned buffers differing (copy nodes re-pointed)", float(badh), 0.0))
    # pipelined begin / end: two calls in flight
    outs = [(torch.empty(6, 1, 12, dtype=torch.int32).pin_memory(), torch.empty(6, 1, dtype=torch.int32).pin_memory(),
             torch.empty(6, 1, 12, dtype=torch.float32).pin_memory()) for _ in range(2)]
    badp = 0
    tick = e.caption_host_begin(hosts[0], m["sos"], m["eos"], 3, 1, 12, outs[0])
    for i in range(1, 5):
        t2 = e.caption_host_begin(hosts[i], m["sos"], m["eos"], 3, 1, 12, outs[i % 2])
        e.caption_host_end(tick)
        badp += sum(1 for a, b in zip(_tokens_list(outs[(i - 1) % 2][0], outs[(i - 1) % 2][1]), ref[i - 1]) if a != b)
        tick = t2
    e.caption_host_end(tick)
    badp += sum(1 for a, b in zip(_tokens_list(outs[0][0], outs[0][1]), ref[4]) if a != b)
    out.append(("pipelined caption_host_begin/_end (two in flight): captions differing", float(badp), 0.0))
    return out


def check_sampling() -> List[Triple]:
    """SURVEY.md 8f N4, second half: the sampling decode branches.  The draws cannot match torch's random stream, so the
    checks are distributional and structural:
      * Gumbel-top-k kernel: over 4000 seeds on one 50-word row, the first draw's frequencies match softmax(logits)
        (chi-square) and the second draw's match the without-replacement law p_j / (1 - p_i) aggregated over i;
      * mode='sampling': every returned log-prob equals the teacher-forced log-prob of that token under the model
        (forward_dec on the sampled prefix), sequences are cut at the first EOS, log-probs are zero after it, and the
        same seed reproduces the same samples while another seed does not;
      * beam_search(sample_or_max='sample'): returned sequences are valid (SOS first, length rules, log-probs consistent
        with forward_dec)."""
    e, g, cfg, sd, x, pads = engine_for("tiny_e2e_peaky", "fp32")
    m = g["meta"]
    out = []
    # ---- kernel-level distribution test through the sampling entry point is indirect; use xn_sample on a 1-step problem:
    # the first sampled word of every row is a Categorical draw from the step-0 distribution, identical for all rows of an image
    B = x.shape[0]
    with torch.no_grad():
        enc = e.forward_enc(x, None)
        sos = torch.full((B, 1), m["sos"], dtype=torch.int64)
        p0 = torch.softmax(e.forward_dec(enc, None, sos, [0] * B, False)[:, 0].double().cpu(), -1)      # (B, V)
    counts = torch.zeros(B, cfg.vocab, dtype=torch.float64)
    n_draw = 0
    for seed in range(250):
        tok, ln, lp = e.sample(x, None, m["sos"], m["eos"], num_outputs=8, max_len=1, seed=1000 + seed)
        first = tok[:, :, 1].cpu().long()                    # (B, 8)
        for b in range(B):
            counts[b] += torch.bincount(first[b], minlength=cfg.vocab).double()
        n_draw += 8
    # chi-square over the words with expected count >= 5, the rest pooled
    worst = 0.0
    for b in range(B):
        exp = p0[b] * n_draw
        big = exp >= 5
        o = torch.cat([counts[b][big], counts[b][~big].sum().reshape(1)])
        ex = torch.cat([exp[big], exp[~big].sum().reshape(1)])
        keep = ex > 0
        chi = float((((o - ex) ** 2)[keep] / ex[keep]).sum())
        dof = int(keep.sum()) - 1
        worst = max(worst, (chi - dof) / math.sqrt(2 * max(dof, 1)))       # standardised: ~N(0,1) under H0
    out.append((f"sampling: first-word frequencies vs softmax (chi-square z-score, {n_draw} draws per image)", worst, 5.0))
    # ---- structure of mode='sampling'
    tok, ln, lp = e.sample(x, None, m["sos"], m["eos"], num_outputs=4, max_len=10, seed=7)
    tok2, ln2, lp2 = e.sample(x, None, m["sos"], m["eos"], num_outputs=4, max_len=10, seed=7)
    tok3, _, _ = e.sample(x, None, m["sos"], m["eos"], num_outputs=4, max_len=10, seed=8)
    out.append(("sampling: same seed gives different samples", float(not (torch.equal(tok, tok2) and torch.equal(lp, lp2))), 0.0))
    out.append(("sampling: different seeds give identical samples", float(torch.equal(tok, tok3)), 0.0))
    tk, lnc, lpc = tok.cpu(), ln.cpu(), lp.cpu()
    bad_struct, worst_lp = 0, 0.0
    with torch.no_grad():
        for b in range(B):
            for j in range(4):
                n = int(lnc[b, j])
                seq = tk[b, j, :n].tolist()
                bad_struct += seq[0] != m["sos"] or n < 2 or n > 11
                if m["eos"] in seq[1:]:
                    bad_struct += seq.index(m["eos"], 1) != n - 1            # cut right after the first EOS
                elif n != 11:
                    bad_struct += 1
                bad_struct += int((tk[b, j, n:] != -1).any()) + int((lpc[b, j, n:] != 0).any()) + int(lpc[b, j, 0] != 0)
                full = e.forward_dec(enc[b:b + 1], None, torch.tensor([seq[:-1]]), [0], True)[0].cpu()      # (n-1, V)
                want = full[torch.arange(n - 1), torch.tensor(seq[1:])]
                worst_lp = max(worst_lp, float((want - lpc[b, j, 1:n]).abs().max()))
    out.append(("sampling: malformed sampled sequences (SOS first, cut after first EOS, padding)", float(bad_struct), 0.0))
    out.append(("sampling: returned log-probs vs teacher-forced log-probs of the sampled words max-abs", worst_lp, 2e-4))
    # ---- beam search with sampled candidates
    tokb, lnb, lpb = e.beam_search_sample(x, None, m["sos"], m["eos"], 3, 2, 12, seed=3)
    tokc, _, _ = e.beam_search_sample(x, None, m["sos"], m["eos"], 3, 2, 12, seed=3)
    tokd, _, _ = e.beam_search(x, None, m["sos"], m["eos"], 3, 2, 12)
    out.append(("beam_search(sample): same seed gives different captions", float(not torch.equal(tokb, tokc)), 0.0))
    tb, lb, pb = tokb.cpu(), lnb.cpu(), lpb.cpu()
    bad_b, worst_b = 0, 0.0
    with torch.no_grad():
        for b in range(B):
            for j in range(2):
                n = int(lb[b, j])
                seq = tb[b, j, :n].tolist()
                bad_b += seq[0] != m["sos"] or n < 2 or n > 12 or (m["eos"] in seq[1:-1])
                full = e.forward_dec(enc[b:b + 1], None, torch.tensor([seq[:-1]]), [0], True)[0].cpu()
                want = full[torch.arange(n - 1), torch.tensor(seq[1:])]
                worst_b = max(worst_b, float((want - pb[b, j, 1:n]).abs().max()))
    out.append(("beam_search(sample): malformed captions", float(bad_b), 0.0))
    out.append(("beam_search(sample): returned log-probs vs teacher-forced log-probs max-abs", worst_b, 2e-4))
    out.append(("beam_search(sample): captions identical to the arg-max search (informational)", float(torch.equal(tokb, tokd)), float("inf")))
    return out


def check_preprocess_batch() -> List[Triple]:
    """xn_preprocess_rgb8_batch: one launch pair for a batch of mixed sizes (host and device inputs), bit-exact vs the
    oracle; 80 distinct sizes through the single-image entry point cycle the 64-entry coefficient-table cache (the
    round-1 eviction bug freed a table still in use) and must stay bit-exact."""
    from oracle import preprocess_oracle as P
    from test_preprocess_oracle import synth_image
    e = bare_engine()
    out = []
    shapes = [(300, 400), (480, 640), (200, 200), (97, 1013), (1, 7), (640, 480), (333, 500), (1080, 1920)]
    imgs = [synth_image(h, w, 13 * h + w) for (h, w) in shapes]
    refs = [P.preprocess_rgb8(im, 96) for im in imgs]
    l0 = e.kernel_launches
    yh = e.preprocess_rgb8(imgs, 96).cpu().numpy()
    out.append(("preprocess batch (8 host images, mixed sizes): kernel launches", float(e.kernel_launches - l0), 2.0))
    out.append(("preprocess batch (host): elements differing from the oracle", float(sum(int((yh[i] != r).sum()) for i, r in enumerate(refs))), 0.0))
    yd = e.preprocess_rgb8([torch.from_numpy(im).cuda() for im in imgs], 96).cpu().numpy()
    out.append(("preprocess batch (device): elements differing from the oracle", float(sum(int((yd[i] != r).sum()) for i, r in enumerate(refs))), 0.0))
    mixed = [torch.from_numpy(im).cuda() if i % 2 else im for i, im in enumerate(imgs)]
    ym = e.preprocess_rgb8(mixed, 96).cpu().numpy()
    out.append(("preprocess batch (host and device mixed): elements differing", float(sum(int((ym[i] != r).sum()) for i, r in enumerate(refs))), 0.0))
    bad = 0
    sizes = [(40 + 3 * i, 50 + 2 * i) for i in range(40)]
    cyc = sizes + sizes[:5] + sizes[::-1]
    for (h, w) in cyc:                                   # > 64 distinct tables, with re-use of old and new entries
        im = synth_image(h, w, h * 7 + w)
        y = e.preprocess_rgb8_single(im, 48).cpu().numpy()
        bad += int((y != P.preprocess_rgb8(im, 48)).sum())
    out.append((f"preprocess single-image entry over {len(cyc)} calls / 80 distinct table sizes: elements differing", float(bad), 0.0))
    return out


def check_early_exit() -> List[Triple]:
    """Device-side early termination (reference captioning_model.py:397 breaks when no beam was extended): inside the
    captured call the decode steps are bodies of CUDA-graph IF nodes.  A model whose EOS bias makes every caption end after
    one word must (a) return exactly the captions of the uncaptured run and of the run with early exit disabled and
    (b) replay measurably faster than with early exit disabled (18 of 19 decode steps are skipped); a model whose captions
    do not end must be unaffected."""
    import time
    from on_device_image_captioning_b200 import synth
    from on_device_image_captioning_b200.config import features_only
    cfg = features_only(vocab=1000, max_seq_len=24)
    x = synth.make_features(cfg, 48, seed=5)
    out = []
    for label, boost in (("captions end at once", 60.0), ("captions run to max_len", -60.0)):
        sd = synth.make_state_dict(cfg, seed=0, profile="peaky", eos_idx=7)
        sd["vocab_linear.bias"][7] += boost
        e = Engine(cfg, 0)
        e.load_state_dict(sd, "fp16")
        res, ms = {}, {}
        for ee in (0, 4):
            e.set_option("early_exit", ee)
            e.set_option("use_graph", 0)
            t0_, l0_, _ = e.beam_search(x, [0] * 48, 5, 7, 3, 1, 20)
            eager = _tokens_list(t0_, l0_)
            e.set_option("use_graph", 1)
            for _ in range(3):                              # eager, capture, replay
                tk, ln, _ = e.beam_search(x, [0] * 48, 5, 7, 3, 1, 20)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                tk, ln, _ = e.beam_search(x, [0] * 48, 5, 7, 3, 1, 20)
            b.record()
            torch.cuda.synchronize()
            ms[ee] = a.elapsed_time(b) / 5
            res[ee] = _tokens_list(tk, ln)
            out.append((f"early exit [{label}] early_exit={ee}: replayed captions differing from the uncaptured run",
                        float(sum(1 for p_, q_ in zip(res[ee], eager) if p_ != q_)), 0.0))
        out.append((f"early exit [{label}]: captions differing between early_exit=1 and 0", float(sum(1 for p_, q_ in zip(res[0], res[4]) if p_ != q_)), 0.0))
        lens = sorted({len(c_) for c_ in res[4]})
        if boost > 0:
            out.append((f"early exit [{label}]: longest caption (tokens incl. SOS/EOS)", float(lens[-1]), 3.0))
            out.append((f"early exit [{label}]: replay time with early exit / without ({ms[4]:.2f} / {ms[0]:.2f} ms)", ms[4] / ms[0], 0.8))
        else:
            out.append((f"early exit [{label}]: shortest caption", float(-lens[0]), -20.0))
            out.append((f"early exit [{label}]: replay time with early exit / without ({ms[4]:.2f} / {ms[0]:.2f} ms) (the IF nodes' own cost)", ms[4] / ms[0], 1.08))
        e.close()
    return out


def check_nvjpeg_decode() -> List[Triple]:
    """SURVEY.md 8f N1, the decode: JPEG files decoded on the GPU by nvJPEG and preprocessed in the same call
    (xn_preprocess_jpeg_batch) against the PIL decode + the same GPU resize.  nvJPEG's IDCT / chroma upsampling is not
    libjpeg's, so the comparison is a closeness bound on the normalised tensors (one grey level = 0.017), plus the exact
    rules: decoded sizes, and the blank canvas for a grayscale (non-RGB) file."""
    import io
    import os
    from PIL import Image
    from conftest import GOLDEN_DIR
    from on_device_image_captioning_b200.image_utils import preprocess_images, preprocess_images_nvjpeg
    from test_preprocess_oracle import reference_preprocess, synth_image
    e = bare_engine()
    if not e.jpeg_available():
        return [("nvJPEG not present on this host (skipped)", 0.0, 0.0)]
    out = []
    paths = [os.path.join(GOLDEN_DIR, "demo_material", n) for n in ("tatin.jpg", "micheal.jpg")]
    a = preprocess_images_nvjpeg(paths, 384, e)
    sizes = list(e.last_jpeg_sizes)
    b = preprocess_images(paths, 384, e)
    d = (a - b).abs()
    out.append(("nvjpeg: decoded (H, W) differ from PIL's", float(sizes != [(960, 1280), (589, 880)]), 0.0))
    out.append(("nvjpeg decode + GPU resize vs PIL decode + GPU resize: mean |diff| of the normalised tensor", float(d.mean()), 0.02))
    out.append(("nvjpeg decode + GPU resize vs PIL decode + GPU resize: max |diff| (grey levels)", float(d.max()) * 0.225 * 255, 24.0))
    # a synthetic 4:4:4 quality-95 JPEG and a grayscale one
    rgb = synth_image(300, 420, 9)
    buf = io.BytesIO()
    Image.fromarray(rgb, "RGB").save(buf, format="JPEG", quality=95, subsampling=0)
    gray = io.BytesIO()
    Image.fromarray(rgb[..., 1], "L").save(gray, format="JPEG", quality=90)
    y = e.preprocess_jpeg([buf.getvalue(), gray.getvalue()], 96).cpu().numpy()
    pil = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB"))
    ref = reference_preprocess(pil, 96)
    out.append(("nvjpeg 4:4:4 q95 synthetic: mean |diff| vs PIL decode", float(np.abs(y[0] - ref).mean()), 0.02))
    blank = reference_preprocess(np.zeros((300, 420, 3), dtype=np.uint8), 96)
    out.append(("nvjpeg grayscale JPEG -> blank canvas rule: elements differing", float((y[1] != blank).sum()), 0.0))
    return out


def check_evaluate_model_loop() -> List[Triple]:
    """the evaluate_model batching loop (reference test.py:141-275) on the drop-in class: sub-batches of
    4 over 10 images (last one ragged) must give, per image, the caption a single-image call gives, in the reference's
    (pred_dict, gts_dict) layout with SOS / EOS stripped."""
    import argparse
    from on_device_image_captioning_b200 import synth
    from on_device_image_captioning_b200.evaluation import evaluate_model
    from on_device_image_captioning_b200.models import End_ExpansionNet_v2
    e0, g, cfg, sd, x, pads = engine_for("tiny_e2e_peaky", "fp32")
    m = g["meta"]
    words = [f"w{i}" for i in range(cfg.vocab)]
    da = argparse.Namespace(enc=0.0, dec=0.0, enc_input=0.0, dec_input=0.0, other=0.0)
    model = End_ExpansionNet_v2(swin_img_size=cfg.img_size, swin_patch_size=cfg.patch_size, swin_in_chans=cfg.in_chans,
                                swin_embed_dim=cfg.embed_dim, swin_depths=list(cfg.depths), swin_num_heads=list(cfg.swin_heads),
                                swin_window_size=cfg.window_size, swin_mlp_ratio=cfg.mlp_ratio, swin_qkv_bias=True, swin_qk_scale=None,
                                swin_drop_rate=0.0, swin_attn_drop_rate=0.0, swin_drop_path_rate=0.0, swin_norm_layer=torch.nn.LayerNorm,
                                swin_ape=False, swin_patch_norm=True, swin_use_checkpoint=False, final_swin_dim=cfg.feat_dim,
                                d_model=cfg.d_model, N_enc=cfg.n_enc, N_dec=cfg.n_dec, ff=cfg.ff, num_heads=cfg.num_heads,
                                num_exp_enc_list=list(cfg.num_exp_enc_list), num_exp_dec=cfg.num_exp_dec,
                                output_word2idx={w: i for i, w in enumerate(words)}, output_idx2word=words,
                                max_seq_len=cfg.max_seq_len, drop_args=da, rank=0, precision="fp32")
    model.load_state_dict(sd)
    model = model.to(0).eval()
    imgs = synth.make_images(cfg, 10, seed=77, kind="mixed")

    class Loader:
        def get_images_by_idx(self, i, dataset_split=None):
            return imgs[i]

        def get_captions_by_idx(self, i, dataset_split=None):
            return [f"reference caption {i} a", f"reference caption {i} b"]

    pred, gts = evaluate_model(model, words, beam_size=3, max_seq_len=12, sos_idx=m["sos"], eos_idx=m["eos"], rank=0,
                               parallel_batches=4, indexes=list(range(10)), data_loader=Loader(),
                               use_images_instead_of_features=True, verbose=False)
    bad = 0
    for i in range(10):
        one, _ = model(enc_x=imgs[i:i + 1].cuda(), enc_x_num_pads=[0], mode="beam_search", beam_size=3, beam_max_seq_len=12,
                       sample_or_max="max", how_many_outputs=1, sos_idx=m["sos"], eos_idx=m["eos"])
        want = " ".join(words[t] for t in one[0][0][1:-1])
        bad += pred[i][0]["caption"] != want or pred[i][0]["image_id"] != i or len(gts[i]) != 2
    return [("evaluate_model loop: images whose caption differs from the single-image call / layout errors", float(bad), 0.0),
            ("evaluate_model loop: predictions returned", float(-len(pred)), -10.0)]


ALL_FP32_MODEL_CASES = ["tiny_e2e_peaky", "tiny_e2e_xavier", "feat_peaky_b5", "feat_xavier_b1", "full_e2e_xavier", "full_e2e_peaky"]


def check_mega_decoder(name: str = "full_e2e_peaky", precision: str = "fp16") -> List[Triple]:
    """The persistent decoder-position kernel (csrc/decode_mega.cu, option use_mega=1) against the default
    one-kernel-per-operation path (same 16-bit arithmetic, different tiling / summation order) and its fused log-softmax +
    top-k against the separate kernel; also that the option really switches paths (launch counts)."""
    e, g, cfg, sd, x, pads = engine_for(name, precision)
    m = g["meta"]
    out = []
    with torch.no_grad():
        enc = O.forward_enc(sd, cfg, x, pads)
    tok = torch.from_numpy(g["dec_tokens"])
    dp = g["dec_pads"].tolist()
    try:
        e.set_option("use_graph", 0)
        e.set_option("use_mega", 1)
        l0 = e.kernel_launches
        lg_mega = e.forward_dec(enc, pads, tok, dp, False).cpu()
        n_mega = e.kernel_launches - l0
        cap_fused = unpack_beam_results(*e.beam_search(x, pads, m["sos"], m["eos"], m["beam"], m["how_many"], m["max_len"]))
        e.set_option("fuse_topk", 0)
        cap_unfused = unpack_beam_results(*e.beam_search(x, pads, m["sos"], m["eos"], m["beam"], m["how_many"], m["max_len"]))
        e.set_option("fuse_topk", 1)
        e.set_option("use_mega", 0)
        l0 = e.kernel_launches
        lg_ops = e.forward_dec(enc, pads, tok, dp, False).cpu()
        n_ops = e.kernel_launches - l0
        cap_ops = unpack_beam_results(*e.beam_search(x, pads, m["sos"], m["eos"], m["beam"], m["how_many"], m["max_len"]))
    finally:
        e.set_option("use_mega", 0)
        e.set_option("fuse_topk", 1)
        e.set_option("use_graph", 1)
    t = tok.shape[1]
    # one launch per position (+ the cross K/V projection and its cast) against ~33 per position
    out.append((f"{name}/{precision} persistent path: launches of a {t}-position teacher-forced decode ({n_mega} vs {n_ops} per-operation)",
                float(n_mega), float(t + 8)))
    out.append((f"{name}/{precision} persistent vs per-operation logits rel-max", rel_max(lg_mega, lg_ops), 2e-3))
    out.append((f"{name}/{precision} fused top-k: caption tokens differing from the separate log-softmax/top-k kernel",
                float(sum(a != b for a, b in zip(cap_fused[0], cap_unfused[0]))), 0.0))
    if tuple(cap_fused[1].shape) == tuple(cap_unfused[1].shape):
        out.append((f"{name}/{precision} fused top-k: caption log-probs vs the separate kernel max-abs",
                    float((cap_fused[1] - cap_unfused[1]).abs().max()), 2e-5))
    out.append((f"{name}/{precision} persistent vs per-operation caption tokens differing (informational)",
                float(sum(a != b for a, b in zip(cap_fused[0], cap_ops[0]))), float("inf")))
    return out


def check_mega_timeline() -> List[Triple]:
    """xn_mega_timeline: CTA 0's %globaltimer stamps of the last persistent-kernel launch are complete and ordered."""
    e, g, cfg, sd, x, pads = engine_for("full_e2e_peaky", "fp16")
    m = g["meta"]
    try:
        e.set_option("use_mega", 1)
        e.set_option("mega_dbg", 1)
        e.beam_search(x, pads, m["sos"], m["eos"], m["beam"], m["how_many"], m["max_len"])
        t = e.mega_timeline()
    finally:
        e.set_option("mega_dbg", 0)
        e.set_option("use_mega", 0)
    n_expected = 2 * (7 * cfg.n_dec + 2) + 2          # start, (before, after) per barrier, end
    out = [("mega timeline: stamps missing", float(abs(len(t) - n_expected)), 0.0),
           ("mega timeline: stamps out of order", float(sum(b < a for a, b in zip(t, t[1:]))), 0.0)]
    if len(t) > 1:
        out.append(("mega timeline: kernel longer than 5 ms", float(t[-1] - t[0]) * 1e-6, 5.0))
    return out


def check_mega_search(name: str = "full_e2e_peaky", precision: str = "fp16") -> List[Triple]:
    """Option mega_search=1 (all time steps of the 'max' search in one launch, bookkeeping and early exit inside the kernel)
    gives the captions of the default one-launch-per-position path."""
    e, g, cfg, sd, x, pads = engine_for(name, precision)
    m = g["meta"]
    args = (x, pads, m["sos"], m["eos"], m["beam"], m["how_many"], m["max_len"])
    e.set_option("use_mega", 1)
    ref = unpack_beam_results(*e.beam_search(*args))
    try:
        e.set_option("mega_search", 1)
        l0 = e.kernel_launches
        e.set_option("use_graph", 0)
        got = unpack_beam_results(*e.beam_search(*args))
        e.set_option("use_graph", 1)
        for _ in range(3):
            got_g = unpack_beam_results(*e.beam_search(*args))
    finally:
        e.set_option("mega_search", 0)
        e.set_option("use_mega", 0)
        e.set_option("use_graph", 1)
    out = [(f"{name}/{precision} whole-search kernel: caption tokens differing from the per-position path",
            float(sum(a != b for a, b in zip(got[0], ref[0]))), 0.0),
           (f"{name}/{precision} whole-search kernel (graph replay): caption tokens differing",
            float(sum(a != b for a, b in zip(got_g[0], ref[0]))), 0.0)]
    if tuple(got[1].shape) == tuple(ref[1].shape):
        out.append((f"{name}/{precision} whole-search kernel: caption log-probs max-abs diff", float((got[1] - ref[1]).abs().max()), 1e-6))
    return out
