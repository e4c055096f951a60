"""Driver for an ncu capture of the persistent decoder-position kernel: eager (un-graphed) beam search from a resident
encoder output.  ncu --kernel-name regex:dec_step_mega --launch-skip 50 --launch-count 1 python tools/prof_mega.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from on_device_image_captioning_b200 import config as C, synth
from on_device_image_captioning_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = C.swin_l_384()
sd = synth.make_state_dict(cfg, 0, "xavier")
x = synth.make_images(cfg, B, 1, "randn").cuda()
e = Engine(cfg, 0)
e.load_state_dict(sd, "fp16")
e.set_option("use_mega", 1)
e.set_option("use_graph", 0)
enc = e.forward_enc(x)
for _ in range(4):
    e.beam_search(enc, None, 79, 77, 3, 1, 20, from_enc=True)
torch.cuda.synchronize()
print("done")
