"""Golden fixture for the ensemble beam search from the UNMODIFIED reference class
legacy_models/ensemble_captioning_model.py:EsembleCaptioningModel (the class test.py:334 builds), run here on CPU on two
synthetic tiny models.  Writes tests/golden/ens_tiny_e2e.npz.   python tests/golden/make_ensemble_golden.py"""
import json, os, sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from on_device_image_captioning_b200 import config as C, synth  # noqa: E402
from oracle import ref_loader as RL  # noqa: E402
from oracle import xnv2_oracle as O  # noqa: E402

B, BEAM, L, HOW, SOS, EOS = 3, 3, 12, 2, 79, 77
cfg = C.swin_tiny_test()
sds = [synth.make_state_dict(cfg, seed=s, profile="peaky", eos_idx=EOS) for s in (0, 1)]
x = synth.make_images(cfg, B, seed=1, kind="mixed")
models = [RL.build_reference_model(cfg, sd) for sd in sds]
from models.ensemble_captioning_model import EsembleCaptioningModel  # noqa: E402  (staged legacy tree)
ens = EsembleCaptioningModel(models, "cpu")
with torch.no_grad():
    tok, lp = ens(enc_x=x, enc_x_num_pads=[0] * B, mode="beam_search", beam_size=BEAM, how_many_outputs=HOW,
                  beam_max_seq_len=L, sample_or_max="max", sos_idx=SOS, eos_idx=EOS)
    tr = {}
    o_tok, o_lp = O.beam_search(sds, cfg, x, [0] * B, SOS, EOS, BEAM, HOW, L, trace=tr)
assert o_tok == tok, "oracle ensemble differs from the reference class"
assert float((o_lp - lp).abs().max()) == 0.0
tokens = np.full((B, HOW, L), -1, dtype=np.int64)
lens = np.zeros((B, HOW), dtype=np.int64)
for b in range(B):
    for j in range(HOW):
        lens[b, j] = len(tok[b][j]); tokens[b, j, : lens[b, j]] = tok[b][j]
meta = dict(cfg=cfg.to_dict(), seeds=[0, 1], profile="peaky", B=B, beam=BEAM, max_len=L, how_many=HOW, sos=SOS, eos=EOS, kind="mixed")
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ens_tiny_e2e.npz"), beam_tokens=tokens, beam_len=lens,
                    beam_logprobs=lp.numpy(), vocab_margin=tr["vocab_margin"].numpy(), merge_margin=tr["merge_margin"].numpy(),
                    final_margin=tr["final_margin"].numpy(), meta_json=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
print("tokens", tok, "\nmargins", tr["vocab_margin"], tr["merge_margin"], tr["final_margin"])
