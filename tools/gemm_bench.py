"""Event-timed tcgen05 GEMM micro-benchmark over the Swin shapes, with the kernel's debug switches
(1 = no epilogue global traffic, 2 = no MMA issue, 4 = no TMA loads) to attribute time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from on_device_image_captioning_b200 import config as C
from on_device_image_captioning_b200.engine import Engine

def main():
    e = Engine(C.swin_tiny_test(), 0)
    if os.environ.get("XNV2_TC_PAIR") is not None:
        e.set_option("tc_pair", int(os.environ["XNV2_TC_PAIR"]))
    Bc = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    modes = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 4, 7]
    shapes = [("s1.qkv", Bc * 9216, 576, 192, 0, False), ("s1.proj", Bc * 9216, 192, 192, 0, True),
              ("s1.fc1", Bc * 9216, 768, 192, 1, False), ("s1.fc2", Bc * 9216, 192, 768, 0, True),
              ("s2.qkv", Bc * 2304, 1152, 384, 0, False), ("s2.fc1", Bc * 2304, 1536, 384, 1, False),
              ("s3.qkv", Bc * 576, 2304, 768, 0, False), ("s3.proj", Bc * 576, 768, 768, 0, True),
              ("s3.fc1", Bc * 576, 3072, 768, 1, False), ("s3.fc2", Bc * 576, 768, 3072, 0, True),
              ("s4.fc1", Bc * 144, 6144, 1536, 1, False), ("big", 8192, 8192, 8192, 0, False),
              # decoder-step shapes (rows = images x beam)
              ("dec.dyn5", 6 * Bc, 2560, 512, 0, False), ("dec.wq", 6 * Bc, 512, 512, 0, False),
              ("dec.wo", 6 * Bc, 512, 512, 0, True), ("dec.ff1", 6 * Bc, 2048, 512, 2, False),
              ("dec.ff2", 6 * Bc, 512, 2048, 0, True), ("dec.red", 6 * Bc, 512, 1536, 0, True),
              ("dec.vocab", 6 * Bc, 10000, 512, 0, False)]
    only = sys.argv[3].split(",") if len(sys.argv) > 3 and sys.argv[3] != "all" else None
    out16 = len(sys.argv) > 4 and sys.argv[4] == "out16"
    prec = "fp16" if out16 else "bf16"
    g = torch.Generator().manual_seed(0)
    for name, M, N, K, act, res in shapes:
        if only and name not in only:
            continue
        x = torch.randn(M, K, generator=g).cuda()
        w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
        b = torch.randn(N, generator=g).cuda()
        r = torch.randn(M, N, generator=g).cuda() if res else None
        if out16 and res:
            continue          # 16-bit outputs never carry the fp32 residual in the engine
        e.set_option("op_out16", 1 if out16 else 0)
        line = f"{name:8s} M={M:7d} N={N:5d} K={K:5d} act{act} res{int(res)} {'out16' if out16 else 'out32'}:"
        for mode in modes:
            e.set_option("tc_debug", mode)
            e.op_linear(x, w, b, r, act, prec)
            e.set_option("profile", 1)
            for _ in range(10 if name.startswith("dec") else 3):
                e.op_linear(x, w, b, r, act, prec)
            ms, fl, n = e.profile_read()
            e.set_option("profile", 0)
            line += f"  dbg{mode}: {ms / n * 1e3:8.1f} us {fl / ms / 1e9:7.1f} TF/s |"
        e.set_option("tc_debug", 0)
        print(line, flush=True)
        del x, w, b, r

if __name__ == "__main__":
    main()
