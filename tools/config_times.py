"""Device timings of the other BASELINE.json configurations (CUDA events, graphs on, 3 warm-up + 10 timed calls):
  C3  decoder-only ExpansionNet_v2 on precomputed (144 x 1536) features, batch 256, beam 5
  C4  per-GPU shape of the 8-GPU run: 512 images per call, beam 3
  C5  batch-1 latency, greedy and beam 3 (also in bench.py's line)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from on_device_image_captioning_b200 import config as C, synth
from on_device_image_captioning_b200.engine import Engine

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
OPTS = {o: int(os.environ["XNV2_" + o.upper()]) for o in ("use_mega", "mega_search", "fuse_topk", "decode_groups") if os.environ.get("XNV2_" + o.upper()) is not None}
print("options:", OPTS, flush=True)
cfg3 = C.features_only()
e3 = Engine(cfg3, 0)
for o, v in OPTS.items(): e3.set_option(o, v)
e3.load_state_dict(synth.make_state_dict(cfg3, 0, "xavier"), prec)
f = synth.make_features(cfg3, 256, 1).cuda()
ms = timeit(lambda: e3.beam_search(f, [0] * 256, 79, 77, 5, 1, 20))
print(f"C3 {prec} features-in B=256 beam 5 max_len 20: {ms:.2f} ms -> {256 / ms * 1e3:.0f} captions/s", flush=True)
e3.close()
cfg = C.swin_l_384()
e = Engine(cfg, 0)
for o, v in OPTS.items(): e.set_option(o, v)
e.load_state_dict(synth.make_state_dict(cfg, 0, "xavier"), prec)
x = synth.make_images(cfg, 512, 1, "randn").cuda()
ms = timeit(lambda: e.beam_search(x, None, 79, 77, 3, 1, 20), n=5)
print(f"C4 {prec} end-to-end B=512 beam 3 max_len 20 (per-GPU shape): {ms:.1f} ms -> {512 / ms * 1e3:.0f} captions/s, workspace {e.workspace_bytes / 2**30:.2f} GiB", flush=True)
x1 = x[:1].contiguous()
for name, bm in (("beam 3", 3), ("greedy", 1)):
    ms = timeit(lambda: e.beam_search(x1, None, 79, 77, bm, 1, 20), n=20)
    print(f"C5 {prec} batch-1 {name}: {ms:.2f} ms", flush=True)
