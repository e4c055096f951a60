"""Python handle over the C ABI.  torch is used only for device memory and streams."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .config import XNConfig

PRECISIONS = {"fp32": _lib.XN_PREC_FP32, "bf16": _lib.XN_PREC_BF16, "fp16": _lib.XN_PREC_FP16}


def _cfg_struct(cfg: XNConfig) -> _lib.XnConfig:
    s = _lib.XnConfig()
    s.has_swin = int(cfg.has_swin)
    s.img_size, s.patch_size, s.in_chans, s.embed_dim = cfg.img_size, cfg.patch_size, cfg.in_chans, cfg.embed_dim
    s.n_stages = len(cfg.depths)
    for i, (d, h) in enumerate(zip(cfg.depths, cfg.swin_heads)):
        s.depths[i] = d
        s.swin_heads[i] = h
    s.window_size = cfg.window_size
    s.mlp_ratio = float(cfg.mlp_ratio)
    s.feat_dim, s.d_model, s.n_enc, s.n_dec, s.ff, s.num_heads = cfg.feat_dim, cfg.d_model, cfg.n_enc, cfg.n_dec, cfg.ff, cfg.num_heads
    s.n_exp_groups = len(cfg.num_exp_enc_list)
    for i, g in enumerate(cfg.num_exp_enc_list):
        s.exp_groups[i] = g
    s.num_exp_dec, s.vocab, s.max_seq_len, s.enc_len = cfg.num_exp_dec, cfg.vocab, cfg.max_seq_len, cfg.enc_len
    return s


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _int_array(v: Optional[Sequence[int]]):
    if v is None:
        return None
    arr = (C.c_int32 * len(v))(*[int(x) for x in v])
    return arr


class Engine:
    """One model instance bound to one CUDA device."""

    def __init__(self, cfg: XNConfig, device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("xnv2_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device("cuda", int(device))
        self._h = C.c_void_p()
        self._cs = _cfg_struct(cfg)
        rc = self.lib.xn_create(C.byref(self._cs), int(device), C.byref(self._h))
        if rc != 0:
            raise RuntimeError("xn_create failed: " + (self.lib.xn_last_error(None) or b"").decode())
        self.precision: Optional[str] = None

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): " + (self.lib.xn_last_error(self._h) or b"").decode())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.xn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.xn_kernel_launches(self._h))

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.xn_workspace_bytes(self._h))

    def set_option(self, name: str, value: int):
        self._check(self.lib.xn_set_option(self._h, name.encode(), int(value)), "xn_set_option")

    def profile_read_min(self, min_flops: float):
        """(ms, flops, launches) of the event-timed tcgen05 GEMMs of at least `min_flops` FLOP each."""
        ms, fl, n = C.c_double(), C.c_double(), C.c_int64()
        self._check(self.lib.xn_profile_read_min(self._h, float(min_flops), C.byref(ms), C.byref(fl), C.byref(n)), "xn_profile_read_min")
        return ms.value, fl.value, n.value

    def mega_timeline(self):
        """%globaltimer stamps (ns) of CTA 0 in the last persistent decoder-position launch (set_option('mega_dbg', 1))."""
        buf = (C.c_uint64 * 128)()
        n = self.lib.xn_mega_timeline(self._h, buf, 128)
        if n < 0:
            self._check(n, "xn_mega_timeline")
        self.mega_clock = (int(buf[124]), int(buf[125]))                       # clock64 at the kernel's start / end (CTA 0)
        self.mega_fine = [int(buf[i]) for i in range(64, 124) if buf[i]]     # intra-phase stamps of the GEMM phases (layer 0, tail)
        return [int(buf[i]) for i in range(n)]

    def profile_kernels(self):
        """{launcher: (launches, total_ms)} of every kernel launch since set_option('profile', 2)."""
        buf = C.create_string_buffer(1 << 16)
        self._check(self.lib.xn_profile_kernels(self._h, buf, len(buf)), "xn_profile_kernels")
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.split("\t")
            out[name.strip()] = (int(n), float(ms))
        return out

    def profile_read(self):
        """(ms, flops, launches) of the event-timed tcgen05 GEMMs since set_option('profile', 1)."""
        ms, fl, n = C.c_double(), C.c_double(), C.c_int64()
        self._check(self.lib.xn_profile_read(self._h, C.byref(ms), C.byref(fl), C.byref(n)), "xn_profile_read")
        return ms.value, fl.value, n.value

    def _f32(self, t: torch.Tensor) -> torch.Tensor:
        return t.to(device=self.device, dtype=torch.float32).contiguous()

    # -- weights ----------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor], precision: str = "fp32"):
        """Replaces model.load_state_dict(checkpoint['model_state_dict']) (reference demo.py:100-104)."""
        for k, v in sd.items():
            if k.endswith("relative_position_index") or k.endswith("attn_mask"):
                continue
            if v.is_sparse:                       # pruned checkpoints store sparse tensors (reference test.py:455-457)
                v = v.to_dense()
            t = v.detach().to(dtype=torch.float32).contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            self._check(self.lib.xn_load_tensor(self._h, k.encode(), C.c_void_p(t.data_ptr()), _lib.XN_DTYPE_F32,
                                                shape, t.dim()), f"xn_load_tensor({k})")
        self.finalize(precision)

    def finalize(self, precision: str):
        self._check(self.lib.xn_finalize_weights(self._h, PRECISIONS[precision]), "xn_finalize_weights")
        self.precision = precision

    # -- model calls ------------------------------------------------------------------------
    def forward_swin(self, images: torch.Tensor) -> torch.Tensor:
        x = self._f32(images)
        B = x.shape[0]
        Hl = (self.cfg.img_size // self.cfg.patch_size) >> (len(self.cfg.depths) - 1)
        out = torch.empty(B, Hl * Hl, self.cfg.feat_dim, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._check(self.lib.xn_forward_swin(self._h, _ptr(x), B, _ptr(out), self._stream()), "xn_forward_swin")
        return out

    def forward_enc(self, enc_input: torch.Tensor, enc_pads: Optional[Sequence[int]] = None) -> torch.Tensor:
        x = self._f32(enc_input)
        B = x.shape[0]
        out = torch.empty(B, self.cfg.enc_len, self.cfg.d_model, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._check(self.lib.xn_forward_enc(self._h, _ptr(x), B, _int_array(enc_pads), _ptr(out), self._stream()),
                        "xn_forward_enc")
        return out

    def forward_dec(self, cross: torch.Tensor, enc_pads: Optional[Sequence[int]], tokens: torch.Tensor,
                    dec_pads: Optional[Sequence[int]], apply_log_softmax: bool = False) -> torch.Tensor:
        cr = self._f32(cross)
        tok = tokens.to(device=self.device, dtype=torch.int64).contiguous()
        R, t = tok.shape
        out = torch.empty(R, t, self.cfg.vocab, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._check(self.lib.xn_forward_dec(self._h, _ptr(cr), R, _int_array(enc_pads), _ptr(tok), t,
                                                _int_array(dec_pads), int(apply_log_softmax), _ptr(out), self._stream()),
                        "xn_forward_dec")
        return out

    def beam_search(self, enc_input: torch.Tensor, enc_pads: Optional[Sequence[int]], sos_idx: int, eos_idx: int,
                    beam_size: int = 3, how_many: int = 1, max_len: int = 20,
                    from_enc: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Device results: tokens (B,how_many,max_len) int32 (-1 padded), lengths (B,how_many) int32,
        log-probs (B,how_many,max_len) f32 (0 padded)."""
        x = self._f32(enc_input)
        B = x.shape[0]
        tok = torch.empty(B, how_many, max_len, device=self.device, dtype=torch.int32)
        ln = torch.empty(B, how_many, device=self.device, dtype=torch.int32)
        lp = torch.empty(B, how_many, max_len, device=self.device, dtype=torch.float32)
        fn = self.lib.xn_beam_search_from_enc if from_enc else self.lib.xn_beam_search
        with torch.cuda.device(self.device):
            self._check(fn(self._h, _ptr(x), B, _int_array(enc_pads), int(beam_size), int(max_len), int(how_many),
                           int(sos_idx), int(eos_idx), _ptr(tok), _ptr(ln), _ptr(lp), self._stream()), "xn_beam_search")
        return tok, ln, lp

    def preprocess_rgb8(self, images, img_size: Optional[int] = None) -> torch.Tensor:
        """Decoded RGB8 images (list of (H, W, 3) uint8 numpy arrays / CPU or CUDA tensors, any sizes) ->
        (B, 3, S, S) float32 on the device: Resize((S, S)) + ToTensor + Normalize of the reference's
        utils/image_utils.py:preprocess_image, bit-identical to Pillow/torchvision (csrc/preprocess.cu).
        The whole list goes through ONE launch pair (xn_preprocess_rgb8_batch); a list mixing host and device
        images is split into two such calls."""
        S = int(img_size or self.cfg.img_size)
        out = torch.empty(len(images), 3, S, S, device=self.device, dtype=torch.float32)
        ts = []
        for im in images:
            t = im if isinstance(im, torch.Tensor) else torch.from_numpy(im)
            assert t.dtype == torch.uint8 and t.dim() == 3 and t.shape[2] == 3, "expected (H, W, 3) uint8 RGB"
            ts.append(t.contiguous())
        with torch.cuda.device(self.device):
            for on_dev in (False, True):
                idx = [i for i, t in enumerate(ts) if t.is_cuda == on_dev]
                if not idx:
                    continue
                if len(idx) == len(ts):
                    dst = out
                else:
                    dst = torch.empty(len(idx), 3, S, S, device=self.device, dtype=torch.float32)
                ptrs = (C.c_void_p * len(idx))(*[ts[i].data_ptr() for i in idx])
                hs = (C.c_int32 * len(idx))(*[int(ts[i].shape[0]) for i in idx])
                ws = (C.c_int32 * len(idx))(*[int(ts[i].shape[1]) for i in idx])
                self._check(self.lib.xn_preprocess_rgb8_batch(self._h, ptrs, int(on_dev), hs, ws, len(idx), _ptr(dst), S, self._stream()),
                            "xn_preprocess_rgb8_batch")
                if dst is not out:
                    out[torch.tensor(idx, device=self.device)] = dst
            if any(not t.is_cuda for t in ts):
                torch.cuda.current_stream().synchronize()      # host buffers were read asynchronously
        return out

    def jpeg_available(self) -> bool:
        return bool(self.lib.xn_jpeg_available())

    def preprocess_jpeg(self, jpeg_streams, img_size: Optional[int] = None) -> torch.Tensor:
        """JPEG byte strings -> (B, 3, S, S) float32 on the device: nvJPEG decode on the GPU, then the batched resize /
        ToTensor / Normalize (xn_preprocess_jpeg_batch).  Raises if libnvjpeg is not present."""
        S = int(img_size or self.cfg.img_size)
        n = len(jpeg_streams)
        out = torch.empty(n, 3, S, S, device=self.device, dtype=torch.float32)
        bufs = [bytes(b) for b in jpeg_streams]
        keep = [C.create_string_buffer(b, len(b)) for b in bufs]
        ptrs = (C.c_void_p * n)(*[C.addressof(k) for k in keep])
        sizes = (C.c_int64 * n)(*[len(b) for b in bufs])
        hs, ws = (C.c_int32 * n)(), (C.c_int32 * n)()
        with torch.cuda.device(self.device):
            self._check(self.lib.xn_preprocess_jpeg_batch(self._h, ptrs, sizes, n, _ptr(out), S, hs, ws, self._stream()), "xn_preprocess_jpeg_batch")
            torch.cuda.current_stream().synchronize()          # the streams were read asynchronously
        self.last_jpeg_sizes = [(int(h), int(w)) for h, w in zip(hs, ws)]
        return out

    def preprocess_rgb8_single(self, image, img_size: Optional[int] = None) -> torch.Tensor:
        """One image through the single-image entry point (xn_preprocess_rgb8): (3, S, S) float32."""
        S = int(img_size or self.cfg.img_size)
        t = image if isinstance(image, torch.Tensor) else torch.from_numpy(image)
        assert t.dtype == torch.uint8 and t.dim() == 3 and t.shape[2] == 3, "expected (H, W, 3) uint8 RGB"
        t = t.contiguous()
        out = torch.empty(3, S, S, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._check(self.lib.xn_preprocess_rgb8(self._h, _ptr(t), int(t.is_cuda), int(t.shape[0]), int(t.shape[1]), _ptr(out), S,
                                                    self._stream()), "xn_preprocess_rgb8")
            if not t.is_cuda:
                torch.cuda.current_stream().synchronize()
        return out

    def beam_search_sample(self, enc_input: torch.Tensor, enc_pads: Optional[Sequence[int]], sos_idx: int, eos_idx: int,
                           beam_size: int = 3, how_many: int = 1, max_len: int = 20, seed: int = 0):
        """beam_search with sample_or_max='sample' (candidates drawn without replacement); same returns as beam_search."""
        x = self._f32(enc_input)
        B = x.shape[0]
        tok = torch.empty(B, how_many, max_len, device=self.device, dtype=torch.int32)
        ln = torch.empty(B, how_many, device=self.device, dtype=torch.int32)
        lp = torch.empty(B, how_many, max_len, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._check(self.lib.xn_beam_search_sample(self._h, _ptr(x), B, _int_array(enc_pads), int(beam_size), int(max_len),
                                                       int(how_many), int(sos_idx), int(eos_idx), C.c_uint64(seed & (2 ** 64 - 1)),
                                                       _ptr(tok), _ptr(ln), _ptr(lp), self._stream()), "xn_beam_search_sample")
        return tok, ln, lp

    def sample(self, enc_input: torch.Tensor, enc_pads: Optional[Sequence[int]], sos_idx: int, eos_idx: int, num_outputs: int = 1,
               max_len: int = 20, seed: int = 0):
        """mode='sampling' (get_batch_multiple_sampled_prediction): tokens (B,num_outputs,max_len+1) int32 (-1 padded),
        lengths (B,num_outputs), log-probs (B,num_outputs,max_len+1)."""
        x = self._f32(enc_input)
        B = x.shape[0]
        tok = torch.empty(B, num_outputs, max_len + 1, device=self.device, dtype=torch.int32)
        ln = torch.empty(B, num_outputs, device=self.device, dtype=torch.int32)
        lp = torch.empty(B, num_outputs, max_len + 1, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._check(self.lib.xn_sample(self._h, _ptr(x), B, _int_array(enc_pads), int(num_outputs), int(max_len), int(sos_idx),
                                           int(eos_idx), C.c_uint64(seed & (2 ** 64 - 1)), _ptr(tok), _ptr(ln), _ptr(lp), self._stream()),
                        "xn_sample")
        return tok, ln, lp

    def overflow_flag(self, clear: bool = True) -> bool:
        """True when a 16-bit-mode call since the last clear produced a non-finite encoder output (fp16 overflow)."""
        f = C.c_int(0)
        self._check(self.lib.xn_overflow_flag(self._h, C.byref(f), int(clear)), "xn_overflow_flag")
        return bool(f.value)

    def caption_host_begin(self, inputs_host: torch.Tensor, sos_idx: int, eos_idx: int, beam_size: int, how_many: int, max_len: int,
                           out: Tuple[torch.Tensor, torch.Tensor, torch.Tensor]) -> int:
        """Pipelined host-buffer call: returns a ticket; the results are in `out` after caption_host_end(ticket)."""
        assert inputs_host.device.type == "cpu" and inputs_host.dtype == torch.float32 and inputs_host.is_contiguous()
        tok, ln, lp = out
        with torch.cuda.device(self.device):
            t = self.lib.xn_caption_host_begin(self._h, _ptr(inputs_host), inputs_host.shape[0], int(beam_size), int(max_len), int(how_many),
                                               int(sos_idx), int(eos_idx), _ptr(tok), _ptr(ln), _ptr(lp), self._stream())
        if t < 0:
            self._check(t, "xn_caption_host_begin")
        return t

    def caption_host_end(self, ticket: int):
        self._check(self.lib.xn_caption_host_end(self._h, int(ticket)), "xn_caption_host_end")

    @staticmethod
    def ensemble_beam_search(engines: Sequence["Engine"], enc_input: torch.Tensor, enc_pads: Optional[Sequence[int]], sos_idx: int,
                             eos_idx: int, beam_size: int = 3, how_many: int = 1, max_len: int = 20):
        """Beam search over an ensemble (reference EsembleCaptioningModel, test.py:334): same returns as beam_search."""
        e0 = engines[0]
        x = e0._f32(enc_input)
        B = x.shape[0]
        tok = torch.empty(B, how_many, max_len, device=e0.device, dtype=torch.int32)
        ln = torch.empty(B, how_many, device=e0.device, dtype=torch.int32)
        lp = torch.empty(B, how_many, max_len, device=e0.device, dtype=torch.float32)
        arr = (C.c_void_p * len(engines))(*[e._h.value for e in engines])
        with torch.cuda.device(e0.device):
            e0._check(e0.lib.xn_ensemble_beam_search(arr, len(engines), _ptr(x), B, _int_array(enc_pads), int(beam_size), int(max_len),
                                                     int(how_many), int(sos_idx), int(eos_idx), _ptr(tok), _ptr(ln), _ptr(lp),
                                                     e0._stream()), "xn_ensemble_beam_search")
        return tok, ln, lp

    def caption_host(self, inputs_host: torch.Tensor, sos_idx: int, eos_idx: int, beam_size: int = 3, how_many: int = 1,
                     max_len: int = 20, out: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None):
        """Host buffers in, host buffers out (H2D/D2H inside the call).  `inputs_host` should be pinned."""
        assert inputs_host.device.type == "cpu" and inputs_host.dtype == torch.float32 and inputs_host.is_contiguous()
        B = inputs_host.shape[0]
        if out is None:
            out = (torch.empty(B, how_many, max_len, dtype=torch.int32).pin_memory(),
                   torch.empty(B, how_many, dtype=torch.int32).pin_memory(),
                   torch.empty(B, how_many, max_len, dtype=torch.float32).pin_memory())
        tok, ln, lp = out
        with torch.cuda.device(self.device):
            self._check(self.lib.xn_caption_host(self._h, _ptr(inputs_host), B, int(beam_size), int(max_len), int(how_many),
                                                 int(sos_idx), int(eos_idx), _ptr(tok), _ptr(ln), _ptr(lp), self._stream()),
                        "xn_caption_host")
        return tok, ln, lp

    # -- single operators (kernel-level tests) ------------------------------------------------
    def op_layernorm(self, x, gamma, beta):
        x = self._f32(x); y = torch.empty_like(x)
        g, b = self._f32(gamma), self._f32(beta)          # keep the temporaries alive across the launch
        rows = x.numel() // x.shape[-1]
        self._check(self.lib.xn_op_layernorm(self._h, _ptr(x), _ptr(g), _ptr(b), _ptr(y), rows,
                                             x.shape[-1], self._stream()), "xn_op_layernorm")
        return y

    def op_linear(self, x, w, bias=None, residual=None, act: int = 0, precision: str = "fp32"):
        x = self._f32(x); w = self._f32(w)
        M, K = x.shape
        N = w.shape[0]
        b = self._f32(bias) if bias is not None else None
        r = self._f32(residual) if residual is not None else None
        y = torch.empty(M, N, device=self.device, dtype=torch.float32)
        self._check(self.lib.xn_op_linear(self._h, _ptr(x), _ptr(w), _ptr(b), _ptr(r), _ptr(y), M, N, K, act,
                                          PRECISIONS[precision], self._stream()), "xn_op_linear")
        return y

    def op_linear_skinny(self, x, w, bias=None, residual=None, act: int = 0, precision: str = "fp16", gamma=None, beta=None,
                         x_is_16bit: bool = False):
        x = self._f32(x); w = self._f32(w)
        M, K = x.shape
        N = w.shape[0]
        b = self._f32(bias) if bias is not None else None
        r = self._f32(residual) if residual is not None else None
        g = self._f32(gamma) if gamma is not None else None
        be = self._f32(beta) if beta is not None else None
        y = torch.empty(M, N, device=self.device, dtype=torch.float32)
        self._check(self.lib.xn_op_linear_skinny(self._h, _ptr(x), _ptr(g), _ptr(be), _ptr(w), _ptr(b), _ptr(r), _ptr(y), M, N, K,
                                                 act, int(x_is_16bit), PRECISIONS[precision], self._stream()), "xn_op_linear_skinny")
        return y

    def op_gemm_raw(self, which: int, a, w16, bias, residual, y, act: int = 0, precision: str = "fp16", gamma=None, beta=None):
        """Measurement hook: one GEMM launch on device operands as they are (a: 16-bit for which 0/1, fp32 for which 2)."""
        M, K = a.shape
        N = w16.shape[0]
        self._check(self.lib.xn_op_gemm_raw(self._h, which, _ptr(a), _ptr(gamma), _ptr(beta), _ptr(w16), _ptr(bias), _ptr(residual),
                                            _ptr(y), M, N, K, act, PRECISIONS[precision], self._stream()), "xn_op_gemm_raw")
        return y

    def op_window_attention(self, qkv, bias_table, B, H, C_, heads, shift, precision: str = "fp32"):
        qkv = self._f32(qkv); bt = self._f32(bias_table)
        out = torch.empty(B * H * H, C_, device=self.device, dtype=torch.float32)
        self._check(self.lib.xn_op_window_attention(self._h, _ptr(qkv), _ptr(bt), _ptr(out), B, H, C_, heads, shift,
                                                    PRECISIONS[precision], self._stream()), "xn_op_window_attention")
        return out

    def op_logsoftmax_topk(self, logits, k, want_logprob=False):
        x = self._f32(logits)
        rows, V = x.shape
        tv = torch.empty(rows, max(k, 1), device=self.device, dtype=torch.float32)
        ti = torch.empty(rows, max(k, 1), device=self.device, dtype=torch.int32)
        lp = torch.empty_like(x) if want_logprob else None
        self._check(self.lib.xn_op_logsoftmax_topk(self._h, _ptr(x), rows, V, k, _ptr(tv), _ptr(ti), _ptr(lp), self._stream()),
                    "xn_op_logsoftmax_topk")
        return tv, ti, lp


class EnginePair:
    """Two engine handles with the same weights on one GPU that take alternate host-buffer calls, each on its own CUDA
    stream, up to four calls in flight (two per handle: xn_caption_host_begin's staging slots).

    One call is a throughput-bound Swin phase followed by a latency-bound decode phase (two chains of ~300 small dependent
    kernels); with two handles the decode chains of two calls run side by side, which measured +5.5 ... 6.7 % captions/s
    over the pipelined single handle at batch 64 (profiles/round2_s4_twin_handles_*.txt, tools/twin_probe.py).  The results
    do not depend on the call pattern (every handle is deterministic and batch-invariant).  Costs a second copy of the
    weights (0.5 GB in 16-bit) and a second workspace."""

    def __init__(self, cfg: XNConfig, device: int = 0):
        self.engines = [Engine(cfg, device), Engine(cfg, device)]
        self.device = self.engines[0].device
        with torch.cuda.device(self.device):
            self.streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        self._next = 0
        self._inflight: Dict[int, Tuple[int, int]] = {}
        self._serial = 0

    def load_state_dict(self, sd: Dict[str, torch.Tensor], precision: str = "fp32"):
        for e in self.engines:
            e.load_state_dict(sd, precision)

    def set_option(self, name: str, value: int):
        for e in self.engines:
            e.set_option(name, value)

    @property
    def kernel_launches(self) -> int:
        return sum(e.kernel_launches for e in self.engines)

    def caption_host_begin(self, inputs_host: torch.Tensor, sos_idx: int, eos_idx: int, beam_size: int, how_many: int, max_len: int,
                           out: Tuple[torch.Tensor, torch.Tensor, torch.Tensor]) -> int:
        """Same contract as Engine.caption_host_begin; tickets are ended in any order, at most four in flight."""
        k = self._next
        with torch.cuda.stream(self.streams[k]):
            t = self.engines[k].caption_host_begin(inputs_host, sos_idx, eos_idx, beam_size, how_many, max_len, out)
        self._next = k ^ 1
        self._serial += 1
        self._inflight[self._serial] = (k, t)
        return self._serial

    def caption_host_end(self, ticket: int):
        if ticket not in self._inflight:
            raise RuntimeError(f"caption_host ticket {ticket} is not in flight")
        k, t = self._inflight.pop(ticket)
        self.engines[k].caption_host_end(t)

    def close(self):
        for e in self.engines:
            e.close()


def unpack_beam_results(tok: torch.Tensor, ln: torch.Tensor, lp: torch.Tensor) -> Tuple[List[List[List[int]]], torch.Tensor]:
    """Device/host result buffers -> the reference's return convention: nested token lists (SOS..EOS
    inclusive, cut at length) and a (B, how_many, max_len_in_batch) zero-padded log-prob tensor
    (reference models/captioning_model.py:401-425)."""
    tok_c, ln_c = tok.cpu(), ln.cpu()
    B, H, _ = tok_c.shape
    res = [[tok_c[b, j, :int(ln_c[b, j])].tolist() for j in range(H)] for b in range(B)]
    mx = int(ln_c.max()) if ln_c.numel() else 0
    return res, lp[:, :, :mx]
