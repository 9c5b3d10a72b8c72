"""GPU parity of the ragged (mixed-length) path -- BASELINE.json config 5 -- through the C-ABI
(ditto_forward_ragged / ditto_p_sample_ragged).  The reference has no padding masks, so the oracle runs every
utterance ALONE at its own length (B = 1), exactly what `SpeechGenerator.__sample_latents` does per utterance
(src/model/SpeechGenerator.py:150-164); the packed CUDA path must reproduce each of them.
Bars (BASELINE.json): rel-L2 <= 1e-4 fp32 path, <= 2e-2 bf16 path."""
import pytest
import torch

import ditto_tts_b200 as D
from ditto_tts_b200.ragged import RaggedBatch
from oracle import ditto_oracle as O  # checker only

pytestmark = pytest.mark.gpu
BAR = {"fp32": 1e-4, "bf16": 2e-2}


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def build_model(cfg, sd, precision, dev, **kw):
    m = D.DiTTO(hidden_dim=cfg.hidden_dim, num_layers=cfg.num_layers, num_heads=cfg.num_heads, time_dim=cfg.time_dim,
                text_dim=cfg.text_dim, diffusion_steps=cfg.diffusion_steps, precision=precision, **kw)
    m.load_state_dict(sd, strict=True)
    return m.to(dev)


def utterances(lengths, text_lens, cfg, seed):
    g = torch.Generator().manual_seed(seed)
    xs = [torch.randn(T, cfg.hidden_dim, generator=g) for T in lengths]
    texts = [torch.randn(S, cfg.text_dim, generator=g) for S in text_lens]
    return xs, texts


def rel(a, b):
    return O.rel_l2(a.float().cpu(), b.float().cpu())


LENGTHS = [37, 64, 37, 150, 9, 64, 1]
TEXT_LENS = [5, 12, 5, 20, 3, 11, 1]       # utterances 0 and 2 share a group; 1 and 5 differ in S only


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("heads", [1, 4])
def test_ragged_forward_vs_oracle_per_utterance(dev, precision, heads):
    cfg = O.OracleConfig(256, 2, heads, 64, 256, 20)
    sd = O.make_state_dict(cfg, 41)
    xs, texts = utterances(LENGTHS, TEXT_LENS, cfg, 42)
    t = torch.tensor([3, 19, 0, 7, 7, 12, 5])
    m = build_model(cfg, sd, precision, dev)
    outs = m.forward_ragged([x.to(dev) for x in xs], [c.to(dev) for c in texts], t.to(dev))
    assert len(outs) == len(xs)
    for i, (x, c) in enumerate(zip(xs, texts)):
        ref = O.ditto_forward(sd, cfg, x[None], c[None], t[i:i + 1])[0]
        assert tuple(outs[i].shape) == tuple(ref.shape)
        assert rel(outs[i], ref) <= BAR[precision], (i, rel(outs[i], ref))


def test_ragged_single_group_equals_uniform_batch(dev):
    """Equal lengths -> one group -> the same kernels as the uniform batch: bitwise equal."""
    cfg = O.OracleConfig(256, 2, 1, 64, 256, 20)
    sd = O.make_state_dict(cfg, 43)
    m = build_model(cfg, sd, "bf16", dev)
    x, text, _ = O.make_inputs(3, 50, 8, cfg, 44)
    t = torch.tensor([1, 2, 3])
    uni = m(x.to(dev), text.to(dev), t.to(dev))
    rag = m.forward_ragged([v.to(dev) for v in x], [c.to(dev) for c in text], t.to(dev))
    assert torch.equal(torch.stack(rag), uni)


def test_ragged_matches_per_utterance_cuda_runs_at_default_width(dev):
    """Repo-default width (H = 768, fused RoPE epilogue with the position table, folded cross-attention): every
    utterance of a packed mixed-length batch equals the same utterance run alone through the uniform path."""
    cfg = O.OracleConfig(768, 2, 1, 256, 768, 50)
    sd = O.make_state_dict(cfg, 45)
    lengths, text_lens = [150, 300, 225, 150, 750], [13, 26, 19, 13, 64]
    xs, texts = utterances(lengths, text_lens, cfg, 46)
    t = torch.tensor([49, 10, 0, 3, 25])
    m = build_model(cfg, sd, "bf16", dev)
    outs = m.forward_ragged([x.to(dev) for x in xs], [c.to(dev) for c in texts], t.to(dev))
    for i in range(len(xs)):
        alone = m(xs[i][None].to(dev), texts[i][None].to(dev), t[i:i + 1].to(dev))[0]
        assert rel(outs[i], alone) <= 1e-5, (i, rel(outs[i], alone))
    ref = O.ditto_forward(sd, cfg, xs[1][None], texts[1][None], t[1:2])[0]
    assert rel(outs[1], ref) <= BAR["bf16"]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("use_graph", [False, True])
def test_ragged_cfg_sampling_vs_oracle(dev, precision, use_graph):
    """Full CFG sampling loop (10 steps, w = 3, uncond = zero text) on a mixed-length batch, noise replayed."""
    steps, w = 10, 3.0
    cfg = O.OracleConfig(256, 2, 1, 64, 256, steps)
    sd = O.make_state_dict(cfg, 47)
    lengths, text_lens = [40, 24, 40, 57], [6, 4, 6, 9]
    xs, texts = utterances(lengths, text_lens, cfg, 48)
    g = torch.Generator().manual_seed(49)
    noise = [torch.randn(steps, T, cfg.hidden_dim, generator=g) for T in lengths]
    s = D.DiTTOSampler(build_model(cfg, sd, precision, dev), guidance_scale=w)
    outs = s.sample_latents_ragged([c.to(dev) for c in texts], lengths, x_init=[x.to(dev) for x in xs],
                                   noise=[z.to(dev) for z in noise], use_graph=use_graph)
    for i in range(len(xs)):
        ref = O.sample_latents(sd, cfg, texts[i][None], xs[i][None], noise[i][:, None], guidance_scale=w)
        ref = ref[0] if isinstance(ref, tuple) else ref
        assert rel(outs[i], ref[0]) <= BAR[precision], (i, rel(outs[i], ref[0]))


def test_ragged_rejects_bad_input(dev):
    cfg = O.OracleConfig(256, 1, 1, 64, 256, 10)
    m = build_model(cfg, O.make_state_dict(cfg, 50), "bf16", dev)
    with pytest.raises(D.DittoError):
        RaggedBatch(m, [torch.zeros(4, 256, device=dev)], [0])
    with pytest.raises(D.DittoError):
        RaggedBatch(m, [torch.zeros(4, 256)], [8])            # CPU tensor: no fallback
    with pytest.raises(D.DittoError):
        RaggedBatch(m, [], [])


def test_ragged_cross_fusion_is_decided_per_group(dev):
    """One utterance with more than 64 text tokens (cross_fused.cu covers S <= 64) must not push the whole mixed batch onto the
    three-kernel cross-attention: the short-text groups keep the fused kernel, the long one takes the flash-style cluster kernel
    (flash_attn768q<CROSS>, which also writes its norm3) -- every utterance still equals the per-utterance (unpadded) oracle."""
    from ditto_tts_b200 import _lib
    cfg = O.OracleConfig(768, 2, 1, 256, 768, 20)
    sd = O.make_state_dict(cfg, 61)
    lengths, text_lens = [150, 300, 150, 90], [13, 100, 13, 64]
    xs, texts = utterances(lengths, text_lens, cfg, 62)
    t = torch.tensor([3, 19, 0, 7])
    m = build_model(cfg, sd, "bf16", dev)
    _lib.profile_start()
    outs = m.forward_ragged([x.to(dev) for x in xs], [c.to(dev) for c in texts], t.to(dev))
    prof = _lib.profile_stop()
    # the long-text group: the flash-style cluster kernel (64 < S <= 256), not the three-kernel composition
    assert "tc_gemm.cross_fused_ln" in prof and "tc_gemm.cross_flash_ln" in prof and "tc_gemm.cross_pv" not in prof, sorted(prof)
    assert prof["tc_gemm.cross_fused_ln"]["launches"] == 2 * cfg.num_layers      # groups (150, 13) and (90, 64)
    assert prof["tc_gemm.cross_flash_ln"]["launches"] == cfg.num_layers          # group (300, 100)
    for i, (x, c) in enumerate(zip(xs, texts)):
        ref = O.ditto_forward(sd, cfg, x[None], c[None], t[i:i + 1])[0]
        assert rel(outs[i], ref) <= BAR["bf16"], (i, rel(outs[i], ref))
