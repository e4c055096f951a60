"""Per-launcher device time of one eager 64-image call (every launch event-timed, xn_profile_kernels), for A/B runs of an
option.   python tools/step_spans.py [batch] [option=value ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from on_device_image_captioning_b200 import config as C, synth
from on_device_image_captioning_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = C.swin_l_384()
e = Engine(cfg, 0)
e.load_state_dict(synth.make_state_dict(cfg, 0, "xavier"), "fp16")
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    e.set_option(k, int(v))
x = synth.make_images(cfg, B, 1, "randn").cuda()
for _ in range(2):
    e.beam_search(x, None, 79, 77, 3, 1, 20)
torch.cuda.synchronize()
e.set_option("profile", 2)
e.beam_search(x, None, 79, 77, 3, 1, 20)
sp = e.profile_kernels()
e.set_option("profile", 0)
tot = sum(v[1] for v in sp.values())
print(f"options {sys.argv[2:]}: eager step {tot:.3f} ms over {sum(v[0] for v in sp.values())} launches")
for k, (n, ms) in sorted(sp.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("SPANS_TOP", "14"))]:
    print(f"   {ms:8.3f} ms {n:5d}x  {k}")
# the tcgen05 GEMMs one by one
e.set_option("profile", 1)
e.forward_swin(x)
ms, fl, n = e.profile_read()
print(f"   Swin tcgen05 GEMMs: {n} launches, {ms:.3f} ms, {fl / ms / 1e9:.1f} TFLOP/s")
e.set_option("profile", 0)
