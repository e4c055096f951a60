"""BASELINE.json configs[0] known answers: the reference's four demo images, file -> caption.

Run in the authoring container (needs /root/reference):   python tests/golden/make_demo_golden.py

Mirrors demo.py:48-133 with the UNMODIFIED reference code: ``utils.image_utils.preprocess_image(path, 384)`` (PIL open,
torchvision Resize((384, 384)) + ToTensor + Normalize), the upstream End_ExpansionNet_v2 class with demo.py's
hyper-parameters, ``model(enc_x=image, enc_x_num_pads=[0], mode="beam_search", beam_size=3, beam_max_seq_len=20,
sample_or_max="max", how_many_outputs=1, sos_idx, eos_idx)`` with the vocabulary of demo_material/demo_coco_tokens.pickle
(SOS 79, EOS 77), and ``utils.language_utils.tokens2description``.  There is no checkpoint in the tree and no network, so
the weights are the deterministic synthetic checkpoint of on_device_image_captioning_b200.synth (profile ``xavier`` = the
reference's own init distributions); SURVEY.md Appendix E lists the same quantities for the reference's
``torch.manual_seed(0)`` init, whose 935 MB of weights cannot travel to the GPU box.

Stored per image (tests/golden/demo_c1.npz): the 384 x 384 RGB8 image after the reference's resize (so the GPU test can
feed the exact pixels for the two files that are too large to commit), the mean of the normalised tensor, mean |encoder
output|, caption tokens, log-probabilities and their sum, the oracle's decision margins, and the caption string.  The two
small JPEGs (tatin, micheal: 370 KB) are committed next to the fixture so the test also runs file -> decode -> GPU resize.
"""
from __future__ import annotations

import json
import os
import pickle
import shutil
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
DEMO = os.path.join(RL.REFERENCE_ROOT, "demo_material")
IMAGES = ["tatin.jpg", "micheal.jpg", "napoleon.jpg", "cat_girl.jpg"]
COMMIT_FILES = ["tatin.jpg", "micheal.jpg"]
BEAM, MAX_LEN = 3, 20


def main():
    torch.set_num_threads(8)
    with open(os.path.join(DEMO, "demo_coco_tokens.pickle"), "rb") as f:
        coco = pickle.load(f)
    w2i, i2w = coco["word2idx_dict"], coco["idx2word_list"]
    sos, eos = w2i[coco["sos_str"]], w2i[coco["eos_str"]]
    cfg = C.swin_l_384(vocab=len(w2i), max_seq_len=74)
    sd = synth.make_state_dict(cfg, seed=0, profile="xavier", eos_idx=eos)
    ref = RL.build_reference_model(cfg, sd, vocab_words=i2w)
    from utils.image_utils import preprocess_image      # the reference's own (staged) module
    from utils.language_utils import tokens2description
    from PIL import Image
    import torchvision

    out = {}
    meta = dict(images=IMAGES, beam=BEAM, max_len=MAX_LEN, sos=int(sos), eos=int(eos), cfg=cfg.to_dict(), profile="xavier",
                weight_fingerprint=sum(float(sd[k].double().abs().sum()) for k in sorted(sd)), torch=torch.__version__,
                captions=[], sizes=[], used_words={})
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for n, name in enumerate(IMAGES):
            path = os.path.join(DEMO, name)
            x = preprocess_image(path, cfg.img_size)                      # (1, 3, 384, 384), reference code
            pil = Image.open(path)
            assert pil.mode == "RGB"
            meta["sizes"].append(list(pil.size))
            u8 = np.asarray(torchvision.transforms.Resize((cfg.img_size, cfg.img_size))(pil), dtype=np.uint8)
            tail = torchvision.transforms.Compose([torchvision.transforms.ToTensor(),
                                                   torchvision.transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
            assert torch.equal(tail(Image.fromarray(u8)).unsqueeze(0), x), "resized pixels do not reproduce the reference tensor"
            pred, lp = ref(enc_x=x, enc_x_num_pads=[0], mode="beam_search", beam_size=BEAM, beam_max_seq_len=MAX_LEN,
                           sample_or_max="max", how_many_outputs=1, sos_idx=sos, eos_idx=eos)
            tr = {}
            o_tok, o_lp = O.beam_search(sd, cfg, x, [0], sos, eos, BEAM, 1, MAX_LEN, trace=tr)
            # batch 1: the oracle's decoder contractions take other BLAS shapes than the reference's, so the log-probs agree
            # to the last bits (measured 9.5e-7), not bit for bit; tokens and encoder output are identical
            assert o_tok == pred and float((o_lp - lp).abs().max()) <= 2e-6, f"{name}: oracle differs from the reference"
            assert torch.equal(tr["enc_out"], ref.forward_enc(x, [0]))
            toks = pred[0][0]
            cap = tokens2description(toks, i2w, sos, eos)
            for t in toks:
                meta["used_words"][str(int(t))] = i2w[t]
            meta["captions"].append(cap)
            out[f"u8_{n}"] = u8
            out[f"tokens_{n}"] = np.array(toks, dtype=np.int64)
            out[f"logprobs_{n}"] = lp[0, 0].numpy().copy()
            out[f"stats_{n}"] = np.array([float(x.double().mean()), float(tr["enc_out"].double().abs().mean()), float(lp.double().sum()),
                                          float(tr["vocab_margin"][0]), float(tr["merge_margin"][0]), float(tr["final_margin"][0])])
            print(name, pil.size, "mean(x) %.6f  mean|enc| %.6f  sum lp %.4f  margins %.2e / %.2e / %.2e" %
                  tuple(out[f"stats_{n}"].tolist()), "\n   ", toks, "\n   ", cap)
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "demo_c1.npz"), **out)
    os.makedirs(os.path.join(HERE, "demo_material"), exist_ok=True)
    for f in COMMIT_FILES:
        shutil.copyfile(os.path.join(DEMO, f), os.path.join(HERE, "demo_material", f))


if __name__ == "__main__":
    main()
