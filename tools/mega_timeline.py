"""Phase timeline of the persistent decoder-position kernel (decode_mega.cu): CTA 0's %globaltimer stamps before / after
every grid barrier of the LAST decoded position.  python tools/mega_timeline.py [B] [beam]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from on_device_image_captioning_b200 import config as C, synth
from on_device_image_captioning_b200.engine import Engine

PHASES = [f"L{l}.{n}" for l in range(3) for n in ("dyn5", "dynexp", "wq", "cross", "wo", "ff1", "ff2")] + ["reduce", "vocab", "merge+beam", "(finalise)"]

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    beam = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    cfg = C.swin_l_384()
    sd = synth.make_state_dict(cfg, 0, "xavier")
    x = synth.make_images(cfg, B, 1, "randn").cuda()
    e = Engine(cfg, 0)
    e.load_state_dict(sd, "fp16")
    e.set_option("use_mega", 1)
    for opt in ("fuse_topk", "mega_coop", "mega_dbg_mode", "mega_search", "early_exit"):
        if os.environ.get("XNV2_" + opt.upper()) is not None:
            e.set_option(opt, int(os.environ["XNV2_" + opt.upper()]))
    enc = e.forward_enc(x)
    e.set_option("mega_dbg", 1)
    for _ in range(4):
        e.beam_search(enc, None, 79, 77, beam, 1, 20, from_enc=True)
    torch.cuda.synchronize()
    t = e.mega_timeline()
    print(f"B={B} beam={beam}: {len(t)} stamps, kernel {1e-3 * (t[-1] - t[0]):.1f} us")
    # stamps: start, then (before barrier, after barrier) per phase, then end
    prev = t[0]
    i = 1
    k = 0
    while i + 1 < len(t):
        name = PHASES[k] if k < len(PHASES) else f"phase{k}"
        print(f"  {name:10s} work(CTA0) {1e-3 * (t[i] - prev):7.2f} us   barrier wait {1e-3 * (t[i + 1] - t[i]):7.2f} us")
        prev = t[i + 1]
        i += 2
        k += 1
    if i < len(t):
        name = PHASES[k] if k < len(PHASES) else f"phase{k}"
        print(f"  {name:10s} work(CTA0) {1e-3 * (t[i] - prev):7.2f} us   (last phase)")
    print("phase boundaries (us since kernel start): " + " ".join(f"{1e-3 * (v - t[0]):.2f}" for v in t))

    c0, c1 = e.mega_clock
    print(f"SM clock during the kernel: {(c1 - c0) / (t[-1] - t[0]) * 1e3:.0f} MHz ({c1 - c0} cycles)")
    f = e.mega_fine
    print("fine stamps (GEMM phases of layer 0, then reduce / vocab; us since kernel start):")
    print("  " + " ".join(f"{1e-3 * (v - t[0]):.2f}" for v in f))


if __name__ == "__main__":
    main()
