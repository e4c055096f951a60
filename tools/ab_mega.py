"""Same-process A/B of the persistent decoder kernel against the per-operation path: Swin+encoder, decode from a resident
encoder output, and the whole call, alternating the setting.  python tools/ab_mega.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from on_device_image_captioning_b200 import config as C, synth
from on_device_image_captioning_b200.engine import Engine


def timeit(fn, n=8, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    cfg = C.swin_l_384()
    sd = synth.make_state_dict(cfg, 0, "xavier")
    x = synth.make_images(cfg, B, 1, "randn").cuda()
    e = Engine(cfg, 0)
    e.load_state_dict(sd, "fp16")
    for opt in ("mega_coop", "early_exit", "fuse_topk", "mega_search", "decode_groups"):
        if os.environ.get("XNV2_" + opt.upper()) is not None:
            e.set_option(opt, int(os.environ["XNV2_" + opt.upper()]))
    enc = e.forward_enc(x)
    for rep in range(2):
        for mega in (1, 0):
            e.set_option("use_mega", mega)
            t_enc = timeit(lambda: e.forward_enc(x))
            t_dec = timeit(lambda: e.beam_search(enc, None, 79, 77, 3, 1, 20, from_enc=True))
            t_all = timeit(lambda: e.beam_search(x, None, 79, 77, 3, 1, 20))
            print(f"B={B} use_mega={mega}: swin+encoder {t_enc:.2f} ms, decode(from enc) {t_dec:.2f} ms, whole call {t_all:.2f} ms "
                  f"-> {B / t_all * 1e3:.1f} captions/s  (call - parts = {t_all - t_enc - t_dec:+.2f} ms)", flush=True)


if __name__ == "__main__":
    main()
