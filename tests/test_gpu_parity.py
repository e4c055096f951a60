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


@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16"])
def test_window_attention(precision):
    import gpu_checks as G
    _assert(G.check_window_attention(precision))


def test_logsoftmax_topk():
    import gpu_checks as G
    _assert(G.check_logsoftmax_topk())


CASES = ["tiny_e2e_peaky", "tiny_e2e_xavier", "feat_peaky_b5", "feat_xavier_b1", "full_e2e_xavier", "full_e2e_peaky"]


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
