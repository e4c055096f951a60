"""One captioning step at the bench configuration, for ncu launch lists:
   python tools/prof_step.py [precision] [batch]   (2 warm-up steps, then 1 step between cudaProfilerStart/Stop)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from on_device_image_captioning_b200 import config as C, synth
from on_device_image_captioning_b200.engine import Engine

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg = C.swin_l_384()
e = Engine(cfg, 0)
e.load_state_dict(synth.make_state_dict(cfg, 0, "xavier"), prec)
e.set_option("use_graph", 0)          # kernels inside a replayed graph are profiled too, but keep names per launch simple
x = synth.make_images(cfg, B, 1, "randn").cuda()
for _ in range(2):
    e.beam_search(x, None, 79, 77, 3, 1, 20)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
e.beam_search(x, None, 79, 77, 3, 1, 20)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", e.kernel_launches)
