"""Swin-only device time per image at several chunk sizes (images per pass through the backbone): separates per-kernel
fixed costs (launch boundary, prologue, tail) from throughput.  python tools/swin_chunk_time.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from on_device_image_captioning_b200 import config as C, synth
from on_device_image_captioning_b200.engine import Engine

def timeit(fn, n=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = C.swin_l_384()
sd = synth.make_state_dict(cfg, 0, "xavier")
x = synth.make_images(cfg, B, 1, "randn").cuda()
e = Engine(cfg, 0)
e.load_state_dict(sd, "fp16")
for chunk in (16, 32, 64, 128, 256):
    if chunk > B: break
    e.set_option("swin_chunk", chunk)
    ms = timeit(lambda: e.forward_swin(x))
    print(f"swin B={B} chunk={chunk}: {ms:.2f} ms -> {ms / B * 64:.2f} ms per 64 images, {207.84 * B / ms:.0f} TFLOP/s", flush=True)
