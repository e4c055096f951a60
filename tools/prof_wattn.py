"""Window-attention kernels at the Swin-L/384 stage shapes: event-timed A/B of the tcgen05 and the mma.sync kernel
(and the launches ncu captures).   python tools/prof_wattn.py [batch] [tc|mma|both] [stage,...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import torch
from on_device_image_captioning_b200 import config as CFG
from on_device_image_captioning_b200.engine import Engine, _ptr

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
which = sys.argv[2] if len(sys.argv) > 2 else "both"
stages = [int(s) for s in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2, 3]
e = Engine(CFG.swin_tiny_test(), 0)
shapes = [(96, 6, 6), (48, 12, 6), (24, 24, 6), (24, 24, 0), (12, 48, 0)]          # (H, heads, shift)
shapes = [s for i, s in enumerate(shapes) if min(i, 3) in stages or (i == 3 and 2 in stages)]
g = torch.Generator(device="cuda").manual_seed(0)
for (H, heads, shift) in shapes:
    Cc = heads * 32
    n = B * H * H
    qkv = torch.randn(n, 3 * Cc, device="cuda", generator=g).half()
    out = torch.empty(n, Cc, device="cuda", dtype=torch.float16)
    table = (torch.randn(529, heads, device="cuda", generator=g) * 0.5)
    bias_t = torch.empty(2 * 532 * heads, device="cuda")
    lib = e.lib
    for kern in (["tc", "mma"] if which == "both" else [which]):
        e.set_option("attn_tc", 1 if kern == "tc" else 0)
        e.set_option("attn_tc_dbg", int(os.environ.get("WATTN_DBG", "0")))
        # raw 16-bit entry: xn_op_window_attention converts fp32 inputs, so time through a tiny loop of the op instead
        x32 = qkv.float()
        y = e.op_window_attention(x32, table, B, H, Cc, heads, shift, "fp16")            # warm (includes the converts)
        torch.cuda.synchronize()
        e.set_option("profile", 2)
        for _ in range(5):
            e.op_window_attention(x32, table, B, H, Cc, heads, shift, "fp16")
        spans = e.profile_kernels()
        e.set_option("profile", 0)
        k = [v for kk, v in spans.items() if "attn" in kk][0]
        algo = 8.0 * n * Cc                                                             # read QKV + write O, 16-bit
        us = k[1] / k[0] * 1e3
        print(f"H={H:3d} heads={heads:2d} shift={shift} B={B} {kern:3s}: {us:8.1f} us per launch  {algo / us / 1e3:7.1f} GB/s algorithmic", flush=True)
        del x32, y
