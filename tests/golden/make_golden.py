"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the authoring container (needs /root/reference):

    python tests/golden/make_golden.py

For every case it builds the synthetic checkpoint (on_device_image_captioning_b200.synth),
loads it into the reference's own upstream classes (legacy_models/, imported as ``models``,
see oracle/ref_loader.py), runs the reference's forward_enc / forward_dec / beam_search on
seeded synthetic inputs and stores the outputs (sub-sampled where large) as
tests/golden/<case>.npz.  The oracle (oracle/xnv2_oracle.py) and the CUDA path are both
tested against these files.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from on_device_image_captioning_b200 import config as C, synth  # noqa: E402
from oracle import ref_loader as RL  # noqa: E402
from oracle import xnv2_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (cfg factory, profile, B, beam, max_len, how_many, enc_pads, input kind)
    "tiny_e2e_peaky": (lambda: C.swin_tiny_test(), "peaky", 3, 3, 12, 2, None, "mixed"),
    "tiny_e2e_xavier": (lambda: C.swin_tiny_test(), "xavier", 2, 2, 10, 1, None, "randn"),
    "full_e2e_xavier": (lambda: C.swin_l_384(), "xavier", 2, 3, 20, 1, None, "randn"),
    "full_e2e_peaky": (lambda: C.swin_l_384(), "peaky", 2, 3, 20, 2, None, "mixed"),
    "feat_peaky_b5": (lambda: C.features_only(), "peaky", 4, 5, 20, 2, [0, 5, 17, 0], "feat"),
    "feat_xavier_b1": (lambda: C.features_only(), "xavier", 2, 1, 16, 1, [0, 0], "feat"),
    # BASELINE.json configs[1] at its full batch: 64 images, beam 3, max_len 20, in ONE reference call
    "c2_b64_xavier": (lambda: C.swin_l_384(), "xavier", 64, 3, 20, 1, None, "randn"),
    "c2_b64_peaky": (lambda: C.swin_l_384(), "peaky", 64, 3, 20, 1, None, "mixed"),
    # the fork's fine-tuning geometry (train.py:381-387: swin_img_size=288, swin_patch_size=3 -> still 96x96 patches)
    # with the layer-removal configuration N_enc = N_dec = 2 (test.py:360-365) on the full Swin-L
    "full_p3_288_n2": (lambda: C.XNConfig(img_size=288, patch_size=3, n_enc=2, n_dec=2), "peaky", 2, 3, 20, 2, None, "mixed"),
}
# large cases keep only what the GPU test compares: sub-sampled features, captions, log-probs, margins
SLIM = {"c2_b64_xavier", "c2_b64_peaky"}


def weight_fingerprint(sd):
    tot = 0.0
    for k in sorted(sd):
        tot += float(sd[k].double().abs().sum())
    return tot


def sub(t: torch.Tensor, n: int = 4096) -> np.ndarray:
    f = t.reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step].contiguous().numpy().copy()


def pad_tokens(tok, how_many, max_len):
    B = len(tok)
    out = np.full((B, how_many, max_len), -1, dtype=np.int64)
    ln = np.zeros((B, how_many), dtype=np.int64)
    for b in range(B):
        for j in range(how_many):
            out[b, j, :len(tok[b][j])] = tok[b][j]
            ln[b, j] = len(tok[b][j])
    return out, ln


def run_case(name):
    mk, profile, B, beam, max_len, how_many, enc_pads, kind = CASES[name]
    cfg = mk()
    sos, eos = 79 % cfg.vocab, 77 % cfg.vocab
    sd = synth.make_state_dict(cfg, seed=0, profile=profile, eos_idx=eos)
    ref = RL.build_reference_model(cfg, sd)
    x = synth.make_features(cfg, B, seed=1) if kind == "feat" else synth.make_images(cfg, B, seed=1, kind=kind)
    pads = enc_pads if enc_pads is not None else [0] * B
    out = {}
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # encoder (reference forward_enc)
        if cfg.has_swin:
            sw = ref.swin_transf(x)
            out["swin_sub"] = sub(sw)
            out["swin_absmean"] = np.float64(sw.double().abs().mean())
        enc = ref.forward_enc(x, pads)
        out["enc_sub"] = sub(enc)
        out["enc_absmean"] = np.float64(enc.double().abs().mean())
        if enc.numel() <= 1 << 18:
            out["enc_full"] = enc.numpy().copy()
        # teacher-forced decoder (reference forward_dec), with decoder pads
        g = torch.Generator().manual_seed(5)
        t = 7
        nd = min(B, 8)                     # the teacher-forced check keeps at most 8 rows of a large batch
        tok = torch.randint(0, cfg.vocab, (nd, t), generator=g)
        dpads = [(3 * i) % 4 for i in range(nd)]
        lp = ref.forward_dec(enc[:nd], pads[:nd], tok, dpads, apply_log_softmax=True)
        lg = ref.forward_dec(enc[:nd], pads[:nd], tok, dpads, apply_log_softmax=False)
        out["dec_tokens"] = tok.numpy()
        out["dec_pads"] = np.array(dpads)
        out["dec_logprob_sub"] = sub(lp)
        out["dec_logits_sub"] = sub(lg)
        tv, ti = torch.topk(lp, 8, dim=-1)
        out["dec_top8_val"] = tv.numpy().copy()
        out["dec_top8_idx"] = ti.numpy().copy()
        # beam search through the reference's public call (upstream call style)
        r_tok, r_lp = ref(enc_x=x, enc_x_num_pads=pads, mode="beam_search", beam_size=beam,
                          beam_max_seq_len=max_len, sample_or_max="max", how_many_outputs=how_many,
                          sos_idx=sos, eos_idx=eos)
        tk, ln = pad_tokens(r_tok, how_many, max_len)
        out["beam_tokens"], out["beam_len"] = tk, ln
        out["beam_logprobs"] = r_lp.numpy().copy()
        # decision margins come from the oracle's trace (the oracle is checked to be
        # token- and logprob-identical to the reference right here)
        tr = {}
        o_tok, o_lp = O.beam_search(sd, cfg, x, pads, sos, eos, beam, how_many, max_len, trace=tr)
        assert o_tok == r_tok, f"{name}: oracle tokens differ from the reference"
        assert torch.equal(o_lp, r_lp), f"{name}: oracle log-probs differ from the reference"
        assert torch.equal(tr["enc_out"], enc)
        out["vocab_margin"] = tr["vocab_margin"].numpy()
        out["merge_margin"] = tr["merge_margin"].numpy()
        out["final_margin"] = tr["final_margin"].numpy()
    if name in SLIM:
        out.pop("enc_full", None)
    meta = dict(case=name, profile=profile, B=B, beam=beam, max_len=max_len, how_many=how_many,
                enc_pads=pads, kind=kind, sos=sos, eos=eos, cfg=cfg.to_dict(),
                weight_fingerprint=weight_fingerprint(sd), input_absmean=float(x.double().abs().mean()),
                torch=torch.__version__)
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "tokens", [r_tok[b][0] for b in range(B)], "lens", ln.tolist(),
          "margins", out["vocab_margin"].min(), out["merge_margin"].min(), out["final_margin"].min())


if __name__ == "__main__":
    torch.set_num_threads(8)
    for n in (sys.argv[1:] or list(CASES)):
        run_case(n)
