"""GPU parity tests: the CUDA library (through the C ABI) vs the CPU oracle and the golden
fixtures produced by the unmodified reference.  Run on the B200 box: pytest -m gpu."""
import pytest

pytestmark = pytest.mark.gpu


def _assert(triples):
    bad = [(l, v, t) for (l, v, t) in triples if not (v <= t)]
    assert not bad, "\n".join(f"{l}: {v:.3e} > {t:.3e}" for l, v, t in bad)


def test_layernorm():
    import gpu_checks as G
    _assert(G.check_layernorm())


def test_linear_fp32():
    import gpu_checks as G
    _assert(G.check_linear_fp32())


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_linear_16bit_tcgen05(precision):
    import gpu_checks as G
    _assert(G.check_linear_bf16(precision))


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_linear_16bit_skinny_decoder_step(precision):
    import gpu_checks as G
    _assert(G.check_linear_skinny(precision))


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_linear_16bit_layernorm_on_load(precision):
    import gpu_checks as G
    _assert(G.check_linear_ln_on_load(precision))


@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16"])
def test_window_attention(precision):
    import gpu_checks as G
    _assert(G.check_window_attention(precision))


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_window_attention_mma_sync_kernel(precision):
    """The mma.sync kernel stays the fallback (odd head counts): keep it under test."""
    import gpu_checks as G
    _assert(G.check_window_attention(precision, "mma"))


def test_logsoftmax_topk():
    import gpu_checks as G
    _assert(G.check_logsoftmax_topk())


CASES = ["tiny_e2e_peaky", "tiny_e2e_xavier", "feat_peaky_b5", "feat_xavier_b1", "full_e2e_xavier", "full_e2e_peaky",
         "full_p3_288_n2"]      # last: swin_patch_size=3 / img 288 (train.py:381-387) with N_enc = N_dec = 2 on Swin-L


@pytest.mark.parametrize("name", CASES)
def test_encoder_fp32(name):
    import gpu_checks as G
    _assert(G.check_encoder(name, "fp32"))


@pytest.mark.parametrize("name", CASES)
def test_decoder_fp32(name):
    import gpu_checks as G
    _assert(G.check_decoder(name, "fp32"))


@pytest.mark.parametrize("name", CASES)
def test_beam_search_fp32_bit_exact_captions(name):
    import gpu_checks as G
    _assert(G.check_beam(name, "fp32"))


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
@pytest.mark.parametrize("name", ["tiny_e2e_peaky", "full_e2e_xavier", "full_e2e_peaky"])
def test_encoder_16bit(name, precision):
    import gpu_checks as G
    _assert(G.check_encoder(name, precision))


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
@pytest.mark.parametrize("name", ["tiny_e2e_peaky", "feat_peaky_b5", "full_e2e_peaky"])
def test_decoder_16bit(name, precision):
    import gpu_checks as G
    _assert(G.check_decoder(name, precision))


@pytest.mark.parametrize("name", ["tiny_e2e_peaky", "feat_peaky_b5"])
def test_beam_search_fp16_runs(name):
    import gpu_checks as G
    _assert(G.check_beam(name, "fp16"))


def _shim_model(cfg, sd, precision):
    import torch
    from argparse import Namespace
    from on_device_image_captioning_b200 import End_ExpansionNet_v2, ExpansionNet_v2
    words = [f"w{i}" for i in range(cfg.vocab)]
    w2i = {w: i for i, w in enumerate(words)}
    da = Namespace(enc=0.0, dec=0.0, enc_input=0.0, dec_input=0.0, other=0.0)
    if cfg.has_swin:
        m = End_ExpansionNet_v2(
            swin_img_size=cfg.img_size, swin_patch_size=cfg.patch_size, swin_in_chans=cfg.in_chans, swin_embed_dim=cfg.embed_dim,
            swin_depths=cfg.depths, swin_num_heads=cfg.swin_heads, swin_window_size=cfg.window_size, swin_mlp_ratio=cfg.mlp_ratio,
            swin_qkv_bias=True, swin_qk_scale=None, swin_drop_rate=0.0, swin_attn_drop_rate=0.0, swin_drop_path_rate=0.0,
            swin_norm_layer=torch.nn.LayerNorm, swin_ape=False, swin_patch_norm=True, swin_use_checkpoint=False,
            final_swin_dim=cfg.feat_dim, d_model=cfg.d_model, N_enc=cfg.n_enc, N_dec=cfg.n_dec, ff=cfg.ff, num_heads=cfg.num_heads,
            num_exp_enc_list=cfg.num_exp_enc_list, num_exp_dec=cfg.num_exp_dec, output_word2idx=w2i, output_idx2word=words,
            max_seq_len=cfg.max_seq_len, drop_args=da, rank=0, precision=precision)
    else:
        m = ExpansionNet_v2(d_model=cfg.d_model, N_enc=cfg.n_enc, N_dec=cfg.n_dec, ff=cfg.ff, num_heads=cfg.num_heads,
                            num_exp_enc_list=cfg.num_exp_enc_list, num_exp_dec=cfg.num_exp_dec, output_word2idx=w2i,
                            output_idx2word=words, max_seq_len=cfg.max_seq_len, drop_args=da, img_feature_dim=cfg.feat_dim,
                            rank=0, precision=precision)
    m.load_state_dict(sd)
    return m.to("cuda:0").eval()


@pytest.mark.parametrize("name", ["tiny_e2e_peaky", "feat_peaky_b5"])
def test_drop_in_classes_both_call_styles(name):
    """The reference's two call styles (demo.py:124 upstream style; quantization.py:133 Captioner style) and the
    teacher-forced forward (test.py:119) through the shim classes, against the reference's golden outputs."""
    import numpy as np
    import torch
    from conftest import golden_setup
    from on_device_image_captioning_b200 import E2E_ExpansionNet_Captioner
    g, cfg, sd, x, pads = golden_setup(name)
    meta = g["meta"]
    m = _shim_model(cfg, sd, "fp32")
    kw = dict(beam_size=meta["beam"], beam_max_seq_len=meta["max_len"], sample_or_max="max", how_many_outputs=meta["how_many"],
              sos_idx=meta["sos"], eos_idx=meta["eos"])
    with torch.no_grad():
        pred, lp = m(enc_x=x.cuda(), enc_x_num_pads=pads, mode="beam_search", **kw)
        cap = E2E_ExpansionNet_Captioner(kw, model=m, rank=0)
        pred2, lp2 = cap(x.cuda(), enc_x_num_pads=pads, mode="beam_search")
        tok = torch.from_numpy(g["dec_tokens"])
        logits = m(enc_x=x.cuda(), dec_x=tok.cuda(), enc_x_num_pads=pads, dec_x_num_pads=g["dec_pads"].tolist(),
                   apply_log_softmax=False)
    assert pred == pred2
    for b in range(meta["B"]):
        for j in range(meta["how_many"]):
            n = int(g["beam_len"][b, j])
            assert pred[b][j] == g["beam_tokens"][b, j, :n].tolist()
    assert lp.shape == tuple(g["beam_logprobs"].shape)
    np.testing.assert_allclose(lp.cpu().numpy(), g["beam_logprobs"], rtol=0, atol=3e-4)
    from conftest import sub
    err = np.abs(sub(logits.cpu()).numpy() - g["dec_logits_sub"]).max()
    assert err < 5e-5, err


def test_config3_features_batch256_beam5():
    """BASELINE.json configs[2] at full size: oracle parity on the first images, batch invariance on the rest."""
    import gpu_checks as G
    _assert(G.check_config3_features_beam5(256))


def test_config4_batch512_per_gpu():
    """BASELINE.json configs[3] per-GPU shape: 512 images per call through the chunked Swin / 1536-row decoder."""
    import gpu_checks as G
    _assert(G.check_config4_batch512_chunking())


def test_preprocess_bit_exact():
    """GPU resize/normalise vs the Pillow-pinned oracle, bit-exact."""
    import gpu_checks as G
    _assert(G.check_preprocess())


def test_feature_extraction_layout():
    """bulk Swin feature extraction in the reference's storage layout."""
    import gpu_checks as G
    _assert(G.check_feature_extraction())


def test_ensemble_beam_search():
    """ensemble beam search vs the reference's EsembleCaptioningModel fixture."""
    import gpu_checks as G
    _assert(G.check_ensemble())


def test_caption_host_paths():
    """The host-buffer C-ABI call (bench.py's e2e): pinned (in-graph chunked copies) and pageable inputs."""
    import gpu_checks as G
    _assert(G.check_caption_host())


def test_engine_pair_alternating_pipelined_calls():
    """Two handles taking alternate host-buffer calls (four in flight) return the single handle's captions."""
    import gpu_checks as G
    _assert(G.check_engine_pair())


@pytest.mark.parametrize("name", ["c2_b64_peaky", "c2_b64_xavier"])
def test_config2_batch64_fp32_vs_reference(name):
    """BASELINE.json configs[1] at batch 64, fp32: bit-exact captions vs the unmodified reference's batch-64 run."""
    import gpu_checks as G
    _assert(G.check_config2_batch64(name, "fp32"))


def test_config2_batch64_fp16():
    import gpu_checks as G
    _assert(G.check_config2_batch64("c2_b64_peaky", "fp16"))


@pytest.mark.parametrize("name", ["full_e2e_peaky", "full_e2e_xavier", "full_p3_288_n2"])
def test_fp16_image_to_logits_rel_max(name):
    """The 16-bit mode end to end on the GPU (decoder fed by the GPU's own encoder output), rel-max vs the fp32 oracle."""
    import gpu_checks as G
    _assert(G.check_image_to_logits_16bit(name, "fp16"))


@pytest.mark.parametrize("name", ["full_e2e_xavier", "feat_peaky_b5"])
def test_round2_tensor_core_options_vs_round1_kernels(name):
    """Static expansion on tcgen05 (se_tc) and the TF32 patch embedding (pe_tc) against the kernels they replace."""
    import gpu_checks as G
    _assert(G.check_tensor_core_options(name, "fp16"))


@pytest.mark.parametrize("name", ["full_e2e_peaky", "feat_peaky_b5"])
def test_decoder_split_k_linears(name):
    """ff2 / reduce-group linears of the decoder step as K slices (batched tcgen05 GEMM + summing LayerNorm)."""
    import gpu_checks as G
    _assert(G.check_decoder_splitk(name, "fp16"))


def test_bf16_image_to_logits_regression_bound():
    """bf16 operands do NOT meet north_star's 2e-3 (one rounding of an 8-bit significand is already 2e-3): this is a
    regression bound on the measured error, not a parity claim."""
    import gpu_checks as G
    _assert(G.check_image_to_logits_16bit("full_e2e_peaky", "bf16"))


def test_fp16_saturation_and_overflow_flag():
    import gpu_checks as G
    _assert(G.check_fp16_saturation())


def test_demo_images_known_answers_config1():
    """BASELINE.json configs[0]: the four demo_material images, file -> caption, vs the unmodified reference."""
    import gpu_checks as G
    _assert(G.check_demo_known_answers())


def test_graph_replays_for_any_input_pointer():
    import gpu_checks as G
    _assert(G.check_graph_pointer_independence())


def test_sampling_modes():
    """mode='sampling' and beam_search(sample_or_max='sample')."""
    import gpu_checks as G
    _assert(G.check_sampling())


def test_preprocess_batch_and_table_cache():
    import gpu_checks as G
    _assert(G.check_preprocess_batch())


def test_evaluate_model_loop():
    """test.py's evaluate_model batching loop on the drop-in class."""
    import gpu_checks as G
    _assert(G.check_evaluate_model_loop())


def test_nvjpeg_decode_path():
    """JPEG decode on the GPU (nvJPEG) feeding the batched preprocessing."""
    import gpu_checks as G
    _assert(G.check_nvjpeg_decode())


def test_device_side_early_exit():
    """reference captioning_model.py:397: stop decoding when every beam has ended -- as graph IF nodes on the device."""
    import gpu_checks as G
    _assert(G.check_early_exit())


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_persistent_decoder_kernel_vs_per_operation_path(precision):
    import gpu_checks as G
    _assert(G.check_mega_decoder("full_e2e_peaky", precision))


def test_persistent_decoder_kernel_features_in_model_with_pads():
    """features-in model: encoder pads (n_valid) and decoder pads (row_len) through the persistent kernel"""
    import gpu_checks as G
    _assert(G.check_mega_decoder("feat_peaky_b5", "fp16"))


def test_persistent_decoder_kernel_timeline_hook():
    import gpu_checks as G
    _assert(G.check_mega_timeline())


def test_whole_search_kernel_option():
    import gpu_checks as G
    _assert(G.check_mega_search())
