"""In-graph latency of the decoder-step GEMM shapes: a CUDA graph of `reps` dependent launches of one GEMM, replayed and
event-timed -> microseconds per launch as the decode loop sees them.  tcgen05 kernel vs the skinny kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from on_device_image_captioning_b200 import config as C
from on_device_image_captioning_b200.engine import Engine

def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 192
    reps = 50
    e = Engine(C.swin_tiny_test(), 0)
    shapes = [("dyn5", 2560, 512, 0, False, True), ("wq", 512, 512, 0, False, False), ("wo", 512, 512, 0, True, False),
              ("ff1", 2048, 512, 2, False, True), ("ff2", 512, 2048, 0, True, False), ("red", 512, 1536, 0, True, True),
              ("vocab", 10000, 512, 0, False, True)]
    g = torch.Generator().manual_seed(0)
    for name, N, K, act, res, a32ok in shapes:
        x32 = torch.randn(M, K, generator=g).cuda()
        x16 = x32.half()
        w16 = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().half()
        b = torch.randn(N, generator=g).cuda()
        r = torch.randn(M, N, generator=g).cuda() if res else None
        ga, be = torch.ones(K).cuda(), torch.zeros(K).cuda()
        y = torch.empty(M, N, device="cuda")
        line = f"{name:6s} M={M} N={N:5d} K={K:4d}:"
        variants = [("tcgen05", 0, x16, None), ("skinny16", 1, x16, None)]
        if a32ok:
            variants.append(("skinny32" + ("+ln" if K <= 1024 else ""), 2, x32, ga if K <= 1024 else None))
        if K == 512:
            variants.append(("tcgen05+ln-on-load", 3, x32, ga))
        for label, which, a, gam in variants:
            try:
                fn = lambda: e.op_gemm_raw(which, a, w16, b, r, y, act, "fp16", gam, be if gam is not None else None)
                fn(); torch.cuda.synchronize()
                s = torch.cuda.Stream()
                with torch.cuda.stream(s):
                    fn()
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr, stream=s):
                        for _ in range(reps):
                            fn()
                    for _ in range(3):
                        gr.replay()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(s)
                    for _ in range(5):
                        gr.replay()
                    e1.record(s)
                    torch.cuda.synchronize()
                line += f"  {label} {e0.elapsed_time(e1) / (5 * reps) * 1e3:6.2f} us |"
            except Exception as ex:
                line += f"  {label} n/a ({str(ex)[:40]}) |"
        print(line, flush=True)

if __name__ == "__main__":
    main()
