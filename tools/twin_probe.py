"""Two engine handles on one GPU, host-buffer calls alternating between them: does the latency-bound decode phase of
one call hide under the Swin phase of the next?   python tools/twin_probe.py [batch] [steps]
Prints captions/s of
  single : one handle, xn_caption_host_begin/_end pipelined (what bench.py's e2e times)
  twin   : two handles, calls i / i+1 on alternating handles, 2 or 4 calls in flight
and checks that the token ids do not depend on the call pattern.

The runs recorded under profiles/round2_s4_twin_handles_*.txt also swept two experimental library options that were
removed again after they measured no gain: `sm_reserve` (the persistent Swin kernels leave k SMs free) and `front_chain`
(a call's Swin + encoder part waits for the previous call's, of any handle, so that one call's decode runs beside the
next call's Swin instead of the two Swin phases interleaving)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from on_device_image_captioning_b200 import config as C, synth
from on_device_image_captioning_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cfg = C.swin_l_384()
sd = synth.make_state_dict(cfg, 0, "xavier")
NH = int(os.environ.get("XNV2_HANDLES", "2"))           # handles in the rotation of the "twin" pattern
engs = [Engine(cfg, 0) for _ in range(NH)]
streams = [torch.cuda.Stream() for _ in range(NH)]      # a handle's call is ordered on the caller's stream: one stream per handle
for e in engs:
    e.load_state_dict(sd, "fp16")
    if os.environ.get("XNV2_USE_GRAPH") is not None:      # 0: eager launches (plain stream-event semantics for front_chain)
        e.set_option("use_graph", int(os.environ["XNV2_USE_GRAPH"]))
host = [synth.make_images(cfg, B, 10 + i, "randn").pin_memory() for i in range(3)]
L = 20
outs = [(torch.empty(B, 1, L, dtype=torch.int32).pin_memory(), torch.empty(B, 1, dtype=torch.int32).pin_memory(),
         torch.empty(B, 1, L, dtype=torch.float32).pin_memory()) for _ in range(2 * max(NH, 2))]


def run(n, twin, depth):
    """`depth` calls in flight; call i goes to handle i % 2 when twin."""
    q = []
    for i in range(n):
        e = engs[i % NH] if twin else engs[0]
        with torch.cuda.stream(streams[i % NH] if twin else streams[0]):
            q.append((e, e.caption_host_begin(host[i % 3], 79, 77, 3, 1, L, outs[i % len(outs)])))
        if len(q) >= depth:
            e0, t0 = q.pop(0)
            e0.caption_host_end(t0)
    for e0, t0 in q:
        e0.caption_host_end(t0)


def timed(twin, depth):
    run(8, twin, depth)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run(STEPS, twin, depth)                # every call is ended (host-synchronised) inside run()
    torch.cuda.synchronize()
    b.record()
    torch.cuda.synchronize()
    return B * STEPS / (a.elapsed_time(b) / 1e3)


def tokens_of(twin, depth):
    """Token ids of three consecutive calls (host[0..2]) through the given call pattern."""
    res, q = [], []

    def end_one():
        e0, t0, i0 = q.pop(0)
        e0.caption_host_end(t0)
        res.append(outs[i0 % len(outs)][0].clone())

    for i in range(3):
        e = engs[i % NH] if twin else engs[0]
        with torch.cuda.stream(streams[i % NH] if twin else streams[0]):
            q.append((e, e.caption_host_begin(host[i % 3], 79, 77, 3, 1, L, outs[i % len(outs)]), i))
        if len(q) >= depth:
            end_one()
    while q:
        end_one()
    return res


want = tokens_of(False, 2)
s1 = timed(False, 2)
t2 = timed(True, NH)
t4 = timed(True, 2 * NH)
same = all(torch.equal(a, b) for a, b in zip(want, tokens_of(True, 2 * NH)))
print(f"{NH} handles: single {s1:7.1f}   rotation, {NH} in flight {t2:7.1f}   rotation, {2 * NH} in flight {t4:7.1f} captions/s"
      f"   tokens {'identical' if same else 'DIFFER'}", flush=True)
