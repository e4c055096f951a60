"""Quick device timings of the main calls (CUDA events).  python tools/quick_time.py [B]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from on_device_image_captioning_b200 import config as C, synth
from on_device_image_captioning_b200.engine import Engine

def timeit(fn, n=5, warm=3):       # warm >= 2: the decode CUDA graph is captured on the second identical call
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    precs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["fp32", "bf16"]
    cfg = C.swin_l_384()
    sd = synth.make_state_dict(cfg, 0, "xavier")
    x = synth.make_images(cfg, B, 1, "randn").cuda()
    e = Engine(cfg, 0)
    for opt in ("pdl", "use_graph", "use_skinny", "use_mega", "fuse_topk", "mega_coop", "tc_pair", "decode_groups", "ln_on_load", "attn_tc", "early_exit", "ln_fuse", "se_tc", "pe_tc", "dec_splitk"):
        if os.environ.get("XNV2_" + opt.upper()) is not None:
            e.set_option(opt, int(os.environ["XNV2_" + opt.upper()]))
    for prec in precs:
        e.load_state_dict(sd, prec)
        l0 = e.kernel_launches
        ms = timeit(lambda: e.forward_swin(x))
        print(f"{prec} swin B={B}: {ms:.2f} ms  ({B/ms*1e3:.1f} img/s, {207.84*B/ms:.1f} TFLOP/s)", flush=True)
        ms = timeit(lambda: e.forward_enc(x))
        print(f"{prec} forward_enc B={B}: {ms:.2f} ms", flush=True)
        enc = e.forward_enc(x)
        ms = timeit(lambda: e.beam_search(enc, None, 79, 77, 3, 1, 20, from_enc=True))
        print(f"{prec} beam(from enc) B={B}: {ms:.2f} ms", flush=True)
        l1 = e.kernel_launches
        ms = timeit(lambda: e.beam_search(x, None, 79, 77, 3, 1, 20))
        print(f"{prec} caption e2e(device) B={B}: {ms:.2f} ms -> {B/ms*1e3:.1f} captions/s   launches/call={(e.kernel_launches-l1)//8}", flush=True)

if __name__ == "__main__":
    main()
