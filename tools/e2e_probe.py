"""Where does the host-buffer path spend its time?  Pure H2D bandwidth of the pinned image batch, the blocking call, the
pipelined begin/end call with and without overlap.   python tools/e2e_probe.py [batch]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from on_device_image_captioning_b200 import config as C, synth
from on_device_image_captioning_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = C.swin_l_384()
eng = Engine(cfg, 0)
eng.load_state_dict(synth.make_state_dict(cfg, 0, "xavier"), "fp16")
host = [synth.make_images(cfg, B, seed=100 + i, kind="randn").pin_memory() for i in range(3)]
dev = torch.empty_like(host[0], device="cuda")
outs = [(torch.empty(B, 1, 20, dtype=torch.int32).pin_memory(), torch.empty(B, 1, dtype=torch.int32).pin_memory(),
         torch.empty(B, 1, 20, dtype=torch.float32).pin_memory()) for _ in range(2)]


def timed(fn, n=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        fn(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


print("H2D %d MB pinned: %.2f ms per copy" % (host[0].numel() * 4 >> 20, timed(lambda i: dev.copy_(host[i % 3], non_blocking=True))))
for i in range(4):
    eng.beam_search(dev, None, 79, 77, 3, 1, 20)
print("device-input call: %.2f ms" % timed(lambda i: eng.beam_search(dev, None, 79, 77, 3, 1, 20)))
for i in range(4):
    eng.caption_host(host[i % 3], 79, 77, 3, 1, 20, out=outs[0])
print("blocking caption_host: %.2f ms" % timed(lambda i: eng.caption_host(host[i % 3], 79, 77, 3, 1, 20, out=outs[0])))


def serial(i):
    t = eng.caption_host_begin(host[i % 3], 79, 77, 3, 1, 20, outs[i % 2])
    eng.caption_host_end(t)


for i in range(6):
    serial(i)
print("begin+end back to back (no overlap): %.2f ms" % timed(serial))


def pipelined(n):
    tick = eng.caption_host_begin(host[0], 79, 77, 3, 1, 20, outs[0])
    for i in range(1, n + 1):
        t0 = time.perf_counter()
        nxt = eng.caption_host_begin(host[i % 3], 79, 77, 3, 1, 20, outs[i % 2]) if i < n else None
        t1 = time.perf_counter()
        eng.caption_host_end(tick)
        t2 = time.perf_counter()
        if i in (3, 4):
            print("   step %d: begin took %.2f ms on the host, end waited %.2f ms" % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
        tick = nxt


pipelined(6)
torch.cuda.synchronize()
t0 = time.perf_counter()
pipelined(10)
torch.cuda.synchronize()
print("pipelined begin/end: %.2f ms per step" % ((time.perf_counter() - t0) / 10 * 1e3))
