#!/usr/bin/env python
"""Headline benchmark: captions/sec of the End_ExpansionNet_v2 (Swin-L/384, beam 3, max_len 20)
captioning path -- BASELINE.json's metric on its configs[1] (batch 64 synthetic 384x384 images
per GPU).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference ...                     # the reference algorithm on the host CPU

A "step" = caption one batch of 64 images per GPU (encoder + beam search, token ids out).  One
process per GPU (torchrun for N > 1); images are sharded by rank, weights replicated; the only
collective is the NCCL all-gather of the int32 caption tokens that ends every step.
`value` is timed with CUDA events with inputs resident in HBM; `e2e` is the same metric through
the C-ABI call that takes HOST buffers (xn_caption_host: H2D of the images and D2H of the
tokens inside the timed region).  The line also carries the tensor-core roofline of the
dominant kernel (the tcgen05 GEMM, event-timed per launch in an extra instrumented step),
a CPU baseline (the oracle port of the reference algorithm on this host's cores, bounded
sample) and the SM clocks seen during the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 64
BEAM, MAX_LEN, SOS, EOS = 3, 20, 79, 77
FLOPS_PER_CAPTION = 215.8e9          # Swin 207.84 + encoder 5.37 + cached decode 2.62 GFLOP
METRIC = "captions_per_sec"
UNIT = "captions/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_sustained=d.get("bf16_tflops_sustained"), bf16_burst=d.get("bf16_tflops"), hbm=d.get("hbm_gbs"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_sustained=1400.0, bf16_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def _traffic():
    """dram read+write bytes per launch of the dominant kernel, from the committed `ncu --set full` capture
    (profiles/gemm_tc_traffic.json, written by tools/ncu_traffic.py); None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "gemm_tc_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p))
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []
        self.active = threading.Event()          # set by the bench around each timed region: idle phases (weight loading,
                                                 # graph capture) read the maximum clock and would mask a power-capped run

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            p = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        while not self.stop_flag.is_set():
            line = p.stdout.readline()
            if not line:
                break
            if self.active.is_set():
                self.rows.append([c.strip() for c in line.split(",")])
        p.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


class CpuReference:
    """The reference's own CPU implementation of the path on the host cores: the UNMODIFIED upstream classes
    (legacy_models/, staged by oracle/stage_ref.py into the git-ignored oracle/_ref/, imported through oracle/ref_loader.py)
    when present -- kind "reference" -- else the oracle port of the same algorithm (kind "port")."""

    def __init__(self, threads: int):
        import torch
        from on_device_image_captioning_b200 import config as C, synth
        torch.set_num_threads(threads)
        self.cfg = C.swin_l_384()
        self.sd = synth.make_state_dict(self.cfg, 0, "xavier")
        self.synth = synth
        self.model = None
        self.kind = "port"
        try:
            from oracle import ref_loader as RL
            if RL.available():
                self.model = RL.build_reference_model(self.cfg, self.sd)
                self.kind = "reference"
        except Exception as ex:                                  # fall back to the port, say why
            print(f"[bench] unmodified reference not usable ({ex!r}); timing the oracle port", file=sys.stderr)
            self.model = None

    def describe(self):
        if self.kind == "reference":
            return ("unmodified reference classes (legacy_models/End_ExpansionNet_v2 + captioning_model.beam_search, torch CPU fp32, "
                    "whole-prefix re-decode), staged copy oracle/_ref")
        return "oracle/xnv2_oracle.py (torch CPU fp32 port of the reference algorithm: whole-prefix re-decode)"

    def captions_per_sec(self, n_images: int, seed: int = 1):
        import torch
        import warnings
        x = self.synth.make_images(self.cfg, n_images, seed, "randn")
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t0 = time.perf_counter()
            if self.model is not None:
                self.model(enc_x=x, enc_x_num_pads=[0] * n_images, mode="beam_search", beam_size=BEAM, beam_max_seq_len=MAX_LEN,
                           sample_or_max="max", how_many_outputs=1, sos_idx=SOS, eos_idx=EOS)
            else:
                from oracle import xnv2_oracle as O
                O.beam_search(self.sd, self.cfg, x, [0] * n_images, SOS, EOS, BEAM, 1, MAX_LEN)
            dt = time.perf_counter() - t0
        return n_images / dt, dt


REF_ARM_BATCH = 8            # the CPU arm runs a reduced batch of 8 per step, reported per caption


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n = REF_ARM_BATCH
    t_start = time.perf_counter()
    cpu = CpuReference(threads)
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt = cpu.captions_per_sec(n, seed=1 + i)
        if i >= args.warmup:
            vals.append(v)
        if vals and time.perf_counter() - t_start > 240:
            break
    val = statistics.mean(vals)
    line = dict(metric=METRIC, value=val, unit=UNIT, impl="reference", n_gpus=args.gpus, steps=len(vals), warmup=args.warmup,
                ms_per_step=1e3 * n / val, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic",
                config=dict(workload="End_ExpansionNet_v2 Swin-L/384, N_enc=3 N_dec=3 d=512, beam=3 max_len=20 (BASELINE.json configs[1]); "
                                     f"each step a bounded sample of {n} of the 64 images on the host CPU", batch=n, beam=BEAM, max_len=MAX_LEN,
                            weights="synthetic xavier-style random init (reference Q5), seed 0"),
                cpu_baseline=dict(value=val, unit=UNIT, cores=threads, kind=cpu.kind,
                                  sample=f"{n} synthetic 384x384 images per step, {cpu.describe()}"),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)
    return 0


def _pin_to_gpu_numa_node(local: int):
    """Bind this process to the CPUs of the GPU's NUMA node before the pinned host buffers are allocated (first touch),
    so that eight ranks feeding eight GPUs do not all pull their 113 MB per step across the socket interconnect."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dv = torch.cuda.get_device_properties(local).pci_device_id
        node_f = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dv:02x}.0/numa_node"
        node = int(open(node_f).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        return None
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from on_device_image_captioning_b200 import config as C, synth
    from on_device_image_captioning_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def note(msg):
        if args.verbose:
            print(f"[bench rank {rank}] {msg} (+{time.perf_counter() - t_begin:.1f}s)", file=sys.stderr, flush=True)

    t_begin = time.perf_counter()
    if world > 1:
        torch.set_num_threads(max(1, (os.cpu_count() or 8) // world))   # torchrun pins OMP_NUM_THREADS=1: weight synthesis is CPU work
        dist.init_process_group("nccl", device_id=dev)
        note("process group up")

    numa = _pin_to_gpu_numa_node(local) if world > 1 else None
    cfg = C.swin_l_384()
    eng = Engine(cfg, local)
    sd0 = synth.make_state_dict(cfg, 0, "xavier")
    eng.load_state_dict(sd0, args.precision)
    if args.swin_chunk:
        eng.set_option("swin_chunk", args.swin_chunk)
    if args.early_exit is not None:
        eng.set_option("early_exit", args.early_exit)
    pair = None                                          # the e2e leg's two handles: loaded here, outside the clock-sampled region
    if not args.no_pair:
        from on_device_image_captioning_b200.engine import EnginePair
        pair = EnginePair(cfg, local)
        pair.load_state_dict(sd0, args.precision)
        if args.swin_chunk:
            pair.set_option("swin_chunk", args.swin_chunk)
        if args.early_exit is not None:
            pair.set_option("early_exit", args.early_exit)
    del sd0
    B = args.batch
    n_rot = 3                                            # rotate inputs so they are never L2-resident
    host = [synth.make_images(cfg, B, seed=100 + rank * n_rot + i, kind="randn").pin_memory() for i in range(n_rot)]
    devin = [h.to(dev) for h in host]
    gathered = torch.empty(world, B, 1, MAX_LEN, dtype=torch.int32, device=dev)

    def step(i):
        # a caller-owned device tensor per batch; the engine stages it into its own buffer, so the call's CUDA graph
        # (captured during warm-up: first call eager, second captured, then replayed) does not depend on the address
        tok, ln, lp = eng.beam_search(devin[i % n_rot], None, SOS, EOS, BEAM, 1, MAX_LEN)
        if world > 1:
            dist.all_gather_into_tensor(gathered, tok)   # the path's only collective: caption token ids
        return tok

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    note("weights loaded, inputs resident")
    for i in range(args.warmup):                         # W >= 3: eager, capture, replay
        step(i)
        note(f"warm-up step {i} enqueued")
    barrier()
    note("warm-up done")
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.profile_region:                             # ncu --profile-from-start off: launch list of exactly the timed steps
        torch.cuda.cudart().cudaProfilerStart()
    sampler.active.set()                             # clocks are sampled inside the timed regions only
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    sampler.active.clear()
    if args.profile_region:
        torch.cuda.cudart().cudaProfilerStop()
    launches = eng.kernel_launches - l0
    ms = e0.elapsed_time(e1)
    note(f"timed region done: {ms:.1f} ms")
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end: host buffers in, host tokens out, through the C-ABI calls a streaming user makes
    # (xn_caption_host_begin / _end: every step's images cross PCIe from pinned host memory and its token ids come back,
    # all inside the timed region; the copy of step i+1 overlaps the compute of step i through the call's two staging slots)
    outs = [(torch.empty(B, 1, MAX_LEN, dtype=torch.int32).pin_memory(), torch.empty(B, 1, dtype=torch.int32).pin_memory(),
             torch.empty(B, 1, MAX_LEN, dtype=torch.float32).pin_memory()) for _ in range(2)]

    def e2e_run(n_steps):
        tick = eng.caption_host_begin(host[0], SOS, EOS, BEAM, 1, MAX_LEN, outs[0])
        for i in range(1, n_steps + 1):
            nxt = eng.caption_host_begin(host[i % n_rot], SOS, EOS, BEAM, 1, MAX_LEN, outs[i % 2]) if i < n_steps else None
            eng.caption_host_end(tick)                   # step i-1's token ids are in host memory now
            if world > 1:
                dist.all_gather_into_tensor(gathered, outs[(i - 1) % 2][0].to(dev, non_blocking=True))
            tick = nxt

    e2e_run(max(6, args.warmup))                        # both staging slots: eager, capture, replay
    barrier()
    sampler.active.set()                             # clocks are sampled inside the timed regions only
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    sampler.active.clear()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(t.item()) / 1e3)
    # the blocking single call (xn_caption_host: copy, compute, copy back, return) for comparison
    for i in range(3):
        eng.caption_host(host[i % n_rot], SOS, EOS, BEAM, 1, MAX_LEN, out=outs[0])
    barrier()
    sampler.active.set()                             # clocks are sampled inside the timed regions only
    e0.record()
    for i in range(args.steps):
        eng.caption_host(host[i % n_rot], SOS, EOS, BEAM, 1, MAX_LEN, out=outs[0])
    e1.record()
    barrier()
    sampler.active.clear()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_blocking = world * B * args.steps / (float(t.item()) / 1e3)
    # ---- the headline end-to-end number: the same loop through an EnginePair (two handles with the same weights taking
    # alternate xn_caption_host_begin/_end calls, four calls in flight), so that the decode chains of two calls run side by
    # side.  Same rules: every step's images come from pinned host memory and its token ids go back to the host inside
    # the timed region; every call is host-synchronised by its caption_host_end before the closing event is recorded.
    e2e_single = e2e_value
    e2e_api = "xn_caption_host_begin/_end (pinned host buffers; two calls in flight, copies inside the timed region)"
    if pair is not None:
        outs4 = outs + [tuple(torch.empty_like(o).pin_memory() for o in outs[0]) for _ in range(2)]

        def pair_run(n_steps):
            q = []

            def end_oldest():
                i, tk = q.pop(0)
                pair.caption_host_end(tk)                # step i's token ids are in host memory now
                if world > 1:
                    dist.all_gather_into_tensor(gathered, outs4[i % 4][0].to(dev, non_blocking=True))

            for i in range(n_steps):
                q.append((i, pair.caption_host_begin(host[i % n_rot], SOS, EOS, BEAM, 1, MAX_LEN, outs4[i % 4])))
                if len(q) == 4:
                    end_oldest()
            while q:
                end_oldest()

        pair_run(12)                                     # both handles, both slots: eager, capture, replay
        barrier()
        sampler.active.set()                             # clocks are sampled inside the timed regions only
        e0.record()
        pair_run(args.steps)
        e1.record()
        barrier()
        sampler.active.clear()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = world * B * args.steps / (float(t.item()) / 1e3)
        e2e_api = ("EnginePair: two xn handles with the same weights, alternate xn_caption_host_begin/_end calls, four calls in "
                   "flight (pinned host buffers, copies inside the timed region)")
        pair.close()
        del pair
    sampler.stop_flag.set()
    h2d = host[0].numel() * 4
    d2h = sum(o.numel() * o.element_size() for o in outs[0])


    # ---- BASELINE.json configs[3]'s per-GPU shape: 512 images per call and GPU (8 Swin chunks of 64, 1536 decoder rows)
    c4 = None
    if not args.no_config4:
        try:
            big = devin[0].repeat(8, 1, 1, 1)[: 512].contiguous() if B == 64 else None
            if big is not None:
                for _ in range(3):
                    eng.beam_search(big, None, SOS, EOS, BEAM, 1, MAX_LEN)
                barrier()
                e0.record()
                for _ in range(3):
                    eng.beam_search(big, None, SOS, EOS, BEAM, 1, MAX_LEN)
                e1.record()
                barrier()
                t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                c4 = dict(workload="512 images per call and GPU, beam 3, max_len 20 (BASELINE.json configs[3] per-GPU shape)",
                          captions_per_s=world * 512 * 3 / (float(t.item()) / 1e3), ms_per_call=float(t.item()) / 3, steps=3)
                del big
        except Exception as ex:
            note(f"config-4 leg skipped: {ex}")

    line = None
    if rank == 0:
        peaks = _peaks()
        roofline = None
        if args.precision != "fp32":
            # ---- roofline of the dominant kernel: every tcgen05 GEMM launch of one more step, event-timed
            runs = []
            for _ in range(3):                                  # median of three instrumented steps (one eager pass is noisy)
                eng.set_option("profile", 1)
                eng.beam_search(devin[0], None, SOS, EOS, BEAM, 1, MAX_LEN)     # rank-local: no collective outside the lock-step region
                runs.append(eng.profile_read() + eng.profile_read_min(1e9))
                eng.set_option("profile", 0)
            g_ms, g_fl, g_n, b_ms, b_fl, b_n = sorted(runs)[1]
            achieved = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
            roofline = dict(bound="tensor", kernel="gemm_tc_kernel (tcgen05.mma, TMA, TMEM)", achieved=achieved,
                            peak=peaks["bf16_sustained"], unit="TFLOP/s", frac=achieved / peaks["bf16_sustained"],
                            traffic=(_traffic() or {}).get("dram_bytes_per_launch"),
                            traffic_detail=_traffic(), peak_source=peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                            launches_timed=g_n, gemm_ms_per_step=g_ms, gemm_share_of_step=g_ms / (ms / args.steps),
                            share_note=("gemm_share_of_step divides the event-timed EAGER GEMM launches by the GRAPHED step time (the ~320 "
                                        "decoder-step launches take 2x longer eager than as graph nodes, so it overstates the share); the "
                                        "like-for-like share, comparable with the ncu launch lists under profiles/, is "
                                        "launcher_time_shares_eager_step.launch_gemm_tc"),
                            algorithmic_tflop_per_step=g_fl / 1e12,
                            # the same figure over the launches of >= 1 GFLOP (Swin, encoder, cross K/V: throughput-bound);
                            # the remainder are the decoder-step GEMMs, bounded by latency (5-13 us each for < 0.1 GFLOP)
                            large_gemms=dict(launches=b_n, ms_per_step=b_ms, achieved=(b_fl / (b_ms * 1e-3) / 1e12 if b_ms > 0 else 0.0),
                                             frac=(b_fl / (b_ms * 1e-3) / 1e12 / peaks["bf16_sustained"] if b_ms > 0 else 0.0)))
        # ---- bandwidth-bound kernels: one more instrumented step with EVERY launch event-timed (graphs off), attributed
        # to its launcher; achieved = algorithmic bytes (tensors that must cross HBM once) / summed time
        hbm_kernels = None
        if args.precision != "fp32":
            eng.set_option("profile", 2)
            eng.beam_search(devin[0], None, SOS, EOS, BEAM, 1, MAX_LEN)
            spans = eng.profile_kernels()
            eng.set_option("profile", 0)
            G = cfg.img_size // cfg.patch_size
            lc = sum(cfg.depths[i] * (G >> i) ** 2 * (cfg.embed_dim << i) for i in range(len(cfg.depths)))   # sum over blocks of L*C
            R = B * BEAM
            algo = {   # launcher -> (what, algorithmic bytes per step)
                "ActOps<T>::attn": ("window_attention_tc_kernel (tcgen05, scores in TMEM): read QKV + write O, 16-bit", 8.0 * lc * B),
                # (the Swin norm1 / norm2 launches of round 1 are gone: folded into the neighbouring tcgen05 GEMMs)
                "launch_logsoftmax_topk": ("logsoftmax_topk_reg_kernel: read R x V logits once", 4.0 * R * cfg.vocab * (MAX_LEN - 1)),
                "launch_cross_attn_step<T, T>": ("cross_attn_step16_kernel: K/V of one layer per launch (KV-cache read by beam rows)",
                                                 2.0 * 2 * cfg.enc_len * cfg.d_model * B * cfg.n_dec * (MAX_LEN - 1)),
                "launch_patch_embed4": ("patch_embed4_kernel: read image fp32, write tokens fp32 + the same rows in 16 bits",
                                       B * (4.0 * cfg.in_chans * cfg.img_size ** 2 + 6.0 * G * G * cfg.embed_dim)),
            }
            hbm_kernels = []
            for key, (what, nbytes) in algo.items():
                if key in spans and spans[key][1] > 0:
                    n, t_ms = spans[key]
                    gbs = nbytes / (t_ms * 1e-3) / 1e9
                    hbm_kernels.append(dict(kernel=what, launches=n, ms_per_step=t_ms, algorithmic_gb_per_step=nbytes / 1e9,
                                            achieved_gbs=gbs, frac_of_hbm_peak=gbs / peaks["hbm"]))
            step_ms = sum(v[1] for v in spans.values())
            shares = {k: round(v[1] / step_ms, 4) for k, v in sorted(spans.items(), key=lambda kv: -kv[1][1])[:12]}
            if roofline is not None:
                roofline["launcher_time_shares_eager_step"] = shares
        # ---- batch-1 latency (BASELINE.json configs[4])
        lat = {}
        one = devin[0][:1].contiguous()
        for name, bm in (("beam3", 3), ("greedy", 1)):
            ts = []
            for i in range(3 + 20):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                eng.beam_search(one, None, SOS, EOS, bm, 1, MAX_LEN)
                b.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(a.elapsed_time(b))
            lat[name + "_p50_ms"] = statistics.median(ts)
        # ---- the survey's stated bar: the reference in eager PyTorch on THIS B200, fp32, TF32 off --
        # the unmodified upstream classes moved to the GPU when staged (oracle/_ref), else the oracle port with CUDA
        # tensors.  Also the config-5 comparator (no TensorRT / ONNX runtime on this image).  First-class key
        # `eager_gpu_baseline`; a baseline next to `value`, never a product path.
        eager = None
        if not args.no_cpu and world == 1:
            try:
                import warnings
                torch.backends.cuda.matmul.allow_tf32 = False
                torch.backends.cudnn.allow_tf32 = False
                sd_host = synth.make_state_dict(cfg, 0, "xavier")
                ref_model, kind = None, "port (oracle/xnv2_oracle.py with CUDA tensors)"
                try:
                    from oracle import ref_loader as RL
                    if RL.available():
                        ref_model = RL.build_reference_model(cfg, sd_host, rank=dev)
                        kind = "reference (unmodified legacy_models classes on cuda, staged copy oracle/_ref)"
                except Exception as ex:
                    note(f"unmodified reference not usable on the GPU ({ex!r}); using the port")
                    ref_model = None
                sd_dev = None if ref_model is not None else {k: v.to(dev) for k, v in sd_host.items()}

                def eager_call(xin, n):
                    with torch.no_grad(), warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        if ref_model is not None:
                            return ref_model(enc_x=xin, enc_x_num_pads=[0] * n, mode="beam_search", beam_size=BEAM, beam_max_seq_len=MAX_LEN,
                                             sample_or_max="max", how_many_outputs=1, sos_idx=SOS, eos_idx=EOS)
                        from oracle import xnv2_oracle as O
                        with torch.device(dev):
                            return O.beam_search(sd_dev, cfg, xin, [0] * n, SOS, EOS, BEAM, 1, MAX_LEN)

                es = ClockSampler(local)
                es.start()
                ts = []
                for i in range(4):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    eager_call(one, 1)
                    torch.cuda.synchronize()
                    if i:
                        ts.append((time.perf_counter() - t0) * 1e3)
                eb = []
                for i in range(3):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    eager_call(devin[i % n_rot], B)
                    torch.cuda.synchronize()
                    if i:
                        eb.append(time.perf_counter() - t0)
                es.stop_flag.set()
                eager = dict(kind=kind, dtype="f32 (TF32 off)", value=B / statistics.median(eb), unit=UNIT, batch=B,
                             batch1_beam3_ms=statistics.median(ts), clocks=es.summary())
                lat["reference_eager_on_this_gpu_beam3_ms"] = statistics.median(ts)
                del sd_dev, ref_model
                torch.cuda.empty_cache()
            except Exception as ex:            # a comparator must never take the bench line down
                eager = dict(unavailable=repr(ex))
                note(f"torch-eager comparator skipped: {ex}")
        # ---- CPU baseline: bounded sample on this host's cores
        threads = os.cpu_count() or 1
        # reported at N = 1 only (the other ranks' processes share these host cores at N > 1)
        cpu = CpuReference(threads) if (not args.no_cpu and world == 1) else None
        cpu_v, cpu_dt = cpu.captions_per_sec(4) if cpu is not None else (None, None)
        clocks = sampler.summary()
        whole_tflops = value / world * FLOPS_PER_CAPTION / 1e12
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype=args.precision if args.precision != "fp32" else "f32", data="synthetic",
                    config=dict(workload="End_ExpansionNet_v2 Swin-L/384, N_enc=3 N_dec=3 d=512, beam=3 max_len=20, "
                                         f"batch {B} synthetic 384x384 images per GPU (BASELINE.json configs[1])",
                                batch_per_gpu=B, global_batch=B * world, beam=BEAM, max_len=MAX_LEN, parallelism=f"dp{world}",
                                weights="synthetic xavier-style random init (reference Q5), seed 0",
                                l2="working set (0.5-0.9 GB of weights + GBs of activations) >> 126 MB L2; inputs rotate over 3 batches",
                                graph="captured during the W warm-up steps (first call eager, second captured, then replayed); the "
                                      "caller's input tensor is staged into a handle-owned buffer, so any input address replays it",
                                precision_note=(f"{args.precision} operands on tcgen05/mma tensor cores with fp32 accumulation for every Linear "
                                                "and the window attention; LayerNorm, softmax, expansion normalisation, residual "
                                                "stream and beam bookkeeping in fp32 (fp16 meets the 2e-3 feature/logit parity "
                                                "target, bf16 measures 6e-3: DESIGN.md)")
                                if args.precision != "fp32" else "all fp32 (parity mode, CUDA-core FFMA GEMMs)"),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             api=e2e_api, single_handle_value=e2e_single,
                             blocking_call_value=e2e_blocking, numa_node=numa),
                    config4_batch512=c4,
                    gpu_launches=int(launches), roofline=roofline,
                    whole_path=dict(algorithmic_tflops_per_gpu=whole_tflops, frac_of_tensor_peak=whole_tflops / peaks["bf16_sustained"],
                                    flops_per_caption=FLOPS_PER_CAPTION),
                    cpu_baseline=(dict(value=cpu_v, unit=UNIT, cores=threads, kind=cpu.kind,
                                       sample=f"4 synthetic 384x384 images, beam 3, max_len 20, {cpu.describe()}, "
                                              f"{threads} threads, {cpu_dt:.1f} s") if cpu_v else None),
                    eager_gpu_baseline=eager,
                    hbm_kernels=hbm_kernels, latency_batch1=lat, clocks=clocks)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def _claim_stdout():
    """Libraries (NCCL prints its version banner) may write to fd 1; the contract wants exactly one JSON line on
    stdout, so fd 1 is pointed at stderr for the run and the JSON goes to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("XNV2_PRECISION", "fp16"), choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--swin-chunk", type=int, default=0)
    ap.add_argument("--early-exit", type=int, default=None, help="decode steps per conditional block (0 = no IF nodes; for profilers that cannot see into them)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-pair", action="store_true", help="e2e through one handle only (skip the EnginePair leg)")
    ap.add_argument("--no-config4", action="store_true", help="skip the extra batch-512 (BASELINE configs[3] per-GPU shape) leg")
    ap.add_argument("--profile-region", action="store_true", help="cudaProfilerStart/Stop around the timed steps (for ncu launch lists)")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
