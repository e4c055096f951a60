"""CPU-side checks added in round 2: the configs[0] demo fixture against the oracle, the evaluate_model batching loop
(incl. its world-size-2 gloo sharding), the staged-reference manifest, and the drop-in classes' pad / sampling surface."""
import json
import os
import socket

import numpy as np
import pytest
import torch

from conftest import load_golden, GOLDEN_DIR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_demo_fixture_oracle_reproduces_reference_caption():
    """tests/golden/demo_c1.npz (made by the unmodified reference on demo_material/tatin.jpg): the oracle, fed the stored
    384x384 pixels through torchvision's ToTensor + Normalize, reproduces the caption tokens and the log-prob sum; the
    committed JPEG decodes + resizes (PIL / torchvision, the reference's own preprocessing stack) to the stored pixels."""
    import torchvision
    from PIL import Image
    from on_device_image_captioning_b200 import synth
    from on_device_image_captioning_b200.config import XNConfig
    from oracle import xnv2_oracle as O
    g = load_golden("demo_c1")
    m = g["meta"]
    assert m["images"] == ["tatin.jpg", "micheal.jpg", "napoleon.jpg", "cat_girl.jpg"] and m["sos"] == 79 and m["eos"] == 77
    cfg = XNConfig(**m["cfg"])
    sd = synth.make_state_dict(cfg, seed=0, profile=m["profile"], eos_idx=m["eos"])
    tail = torchvision.transforms.Compose([torchvision.transforms.ToTensor(),
                                           torchvision.transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    for n, name in enumerate(m["images"][:2]):
        pil = Image.open(os.path.join(GOLDEN_DIR, "demo_material", name))
        u8 = np.asarray(torchvision.transforms.Resize((cfg.img_size, cfg.img_size))(pil), dtype=np.uint8)
        assert np.array_equal(u8, g[f"u8_{n}"]), f"{name}: decoded + resized pixels differ from the fixture"
    torch.set_num_threads(max(1, min(8, torch.get_num_threads())))
    x = tail(Image.fromarray(g["u8_0"])).unsqueeze(0)
    assert abs(float(x.double().mean()) - float(g["stats_0"][0])) < 1e-9
    with torch.no_grad():
        tok, lp = O.beam_search(sd, cfg, x, [0], m["sos"], m["eos"], m["beam"], 1, m["max_len"])
    assert tok[0][0] == g["tokens_0"].tolist()
    assert abs(float(lp.double().sum()) - float(g["stats_0"][2])) < 1e-4
    from on_device_image_captioning_b200.language_utils import tokens2description
    words = {int(k): v for k, v in m["used_words"].items()}

    class V:
        def __getitem__(self, i):
            return words[int(i)]
    assert tokens2description(tok[0][0], V(), m["sos"], m["eos"]) == m["captions"][0]


def test_sub_batch_ranges_and_num_pads():
    from on_device_image_captioning_b200.evaluation import sub_batch_ranges, compute_num_pads
    assert sub_batch_ranges(10, 4) == [(0, 4), (4, 8), (8, 10)]
    assert sub_batch_ranges(8, 4) == [(0, 4), (4, 8)]
    assert sub_batch_ranges(1, 16) == [(0, 1)]
    assert compute_num_pads([[1, 2, 3], [1], [1, 2]]) == [0, 2, 1]          # reference utils/language_utils.py:4-13


class _FakeModel:
    """Deterministic stand-in with the drop-in call surface: caption = f(image content)."""
    training = False

    def eval(self):
        return self

    def train(self):
        return self

    def __call__(self, enc_x, enc_x_num_pads, mode, **kw):
        assert mode == "beam_search" and len(enc_x_num_pads) == enc_x.shape[0]
        out = []
        for i in range(enc_x.shape[0]):
            k = int(enc_x[i].sum().round().item()) % 50
            out.append([[kw["sos_idx"], 10 + k, 11 + k, kw["eos_idx"]]])
        return out, None


class _Loader:
    def get_images_by_idx(self, i, dataset_split=None):
        return torch.full((3, 2, 2), float(i))

    def get_captions_by_idx(self, i, dataset_split=None):
        return [f"gt {i}"]


def _expected(n):
    return {i: f"w{10 + (12 * i) % 50} w{11 + (12 * i) % 50}" for i in range(n)}


def test_evaluate_model_loop_cpu():
    from on_device_image_captioning_b200.evaluation import evaluate_model
    words = [f"w{i}" for i in range(100)]
    pred, gts = evaluate_model(_FakeModel(), words, 3, 12, 79, 77, "cpu", parallel_batches=4, indexes=list(range(10)),
                               data_loader=_Loader(), use_images_instead_of_features=True, verbose=False)
    assert {i: pred[i][0]["caption"] for i in pred} == _expected(10)
    assert all(gts[i] == [{"image_id": i, "caption": f"gt {i}"}] for i in range(10))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _eval_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from on_device_image_captioning_b200.evaluation import evaluate_model
    words = [f"w{i}" for i in range(100)]
    pred, gts = evaluate_model(_FakeModel(), words, 3, 12, 79, 77, "cpu", parallel_batches=3, indexes=list(range(11)),
                               data_loader=_Loader(), use_images_instead_of_features=True, verbose=False, shard=True)
    q.put((rank, None if pred is None else {i: pred[i][0]["caption"] for i in pred}))
    dist.barrier()
    dist.destroy_process_group()


def test_evaluate_model_sharded_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_eval_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res[0] == _expected(11) and res[1] is None


def test_staged_reference_manifest_is_unmodified():
    """oracle/_ref (git-ignored, staged by oracle/stage_ref.py) must be a byte-for-byte copy: every file's SHA-256 is in
    its manifest.  Skipped where nothing is staged."""
    import hashlib
    ref = os.path.join(ROOT, "oracle", "_ref")
    man = os.path.join(ref, "MANIFEST.json")
    if not os.path.exists(man):
        pytest.skip("oracle/_ref not staged")
    files = json.load(open(man))["files"]
    assert any(f["staged"] == os.path.join("models", "captioning_model.py") for f in files)
    for f in files:
        assert hashlib.sha256(open(os.path.join(ref, f["staged"]), "rb").read()).hexdigest() == f["sha256"], f["staged"]


def test_staged_reference_runs_and_matches_a_fixture():
    """The staged copy imported through oracle/ref_loader.py is the reference arm bench.py times: on the tiny fixture it
    must reproduce the committed golden captions (which were made from /root/reference)."""
    from oracle import ref_loader as RL
    if not RL.available():
        pytest.skip("no reference available")
    from conftest import golden_setup
    g, cfg, sd, x, pads = golden_setup("tiny_e2e_peaky")
    m = g["meta"]
    import warnings
    ref = RL.build_reference_model(cfg, sd)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tok, lp = ref(enc_x=x, enc_x_num_pads=pads, mode="beam_search", beam_size=m["beam"], beam_max_seq_len=m["max_len"],
                      sample_or_max="max", how_many_outputs=m["how_many"], sos_idx=m["sos"], eos_idx=m["eos"])
    for b in range(m["B"]):
        for j in range(m["how_many"]):
            assert tok[b][j] == g["beam_tokens"][b, j, : int(g["beam_len"][b, j])].tolist()


def test_pad_rule_for_end_to_end_models():
    from on_device_image_captioning_b200.models import _check_no_enc_pads
    _check_no_enc_pads(None, "m")
    _check_no_enc_pads([0], "m")                 # demo.py / benchmarking.py pass the default [0] for any batch
    _check_no_enc_pads([0, 0, 0], "m")
    with pytest.raises(AssertionError, match="End to End case have no padding"):
        _check_no_enc_pads([0, 2], "End to End case have no padding")


def test_engine_pair_ticket_bookkeeping_without_a_gpu():
    """EnginePair's host logic (alternation between the two handles, tickets ended in any order, unknown tickets refused)
    on stand-in handles: the real handles need a GPU (tests/test_gpu_parity.py::test_engine_pair_alternating_pipelined_calls)."""
    from on_device_image_captioning_b200.engine import EnginePair

    class FakeHandle:
        def __init__(self, name):
            self.name, self.begun, self.ended, self.slot = name, [], [], 0

        def caption_host_begin(self, x, *a):
            self.begun.append(x)
            self.slot ^= 1
            return self.slot ^ 1                       # the C ABI's ticket: the staging slot, 0 / 1 alternating

        def caption_host_end(self, t):
            self.ended.append(t)

    pair = object.__new__(EnginePair)
    pair.engines = [FakeHandle("a"), FakeHandle("b")]
    pair.streams = [None, None]                        # torch.cuda.stream(None) is a no-op context
    pair.device, pair._next, pair._inflight, pair._serial = 0, 0, {}, 0
    tickets = [pair.caption_host_begin(i, 1, 2, 3, 1, 20, None) for i in range(4)]
    assert len(set(tickets)) == 4
    assert pair.engines[0].begun == [0, 2] and pair.engines[1].begun == [1, 3]
    pair.caption_host_end(tickets[2])                  # out of order: handle a's second slot
    pair.caption_host_end(tickets[1])
    assert pair.engines[0].ended == [1] and pair.engines[1].ended == [0]
    with pytest.raises(RuntimeError):
        pair.caption_host_end(tickets[2])              # already ended
    with pytest.raises(RuntimeError):
        pair.caption_host_end(999)
    pair.caption_host_end(tickets[0])
    pair.caption_host_end(tickets[3])
    assert not pair._inflight


def test_bench_clock_sampler_summary_uses_only_active_samples():
    """bench.py's clock sampler: rows are kept only while a timed region is active, the reported SM clock is their plain
    median (idle phases read the maximum clock and must not mask a power-capped run), throttle reasons are collected."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(GOLDEN_DIR), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.ClockSampler(0)
    assert not s.active.is_set()
    assert s.summary()["sm_mhz"] is None and s.summary()["samples"] == 0
    na, act = "Not Active", "Active"
    s.rows = [["1800", "1965", "900.0", na, na, na, act], ["1750", "1965", "950.0", na, na, na, act],
              ["1965", "1965", "300.0", na, na, na, na]]
    out = s.summary()
    assert out["sm_mhz"] == 1800.0 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 3
    assert out["reasons"] == ["sw_power_cap"]
