"""GPU parity of the sampler variants: DDIM / fewer steps / CFG scale 5 / shifted schedule through
DiTTOSampler (C-ABI: ditto_engine_load_update_table + ditto_p_sample) against the oracle's textbook restatements."""
import pytest
import torch

import ditto_tts_b200 as D
from oracle import ditto_oracle as O

pytestmark = pytest.mark.gpu
BAR = {"fp32": 1e-4, "bf16": 2e-2}      # BASELINE.json: rel-L2 vs the fp32 oracle


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def setup():
    cfg = O.OracleConfig(256, 2, 2, 64, 256, 20)
    sd = O.make_state_dict(cfg, 11)
    x, text, noise = O.make_inputs(2, 40, 9, cfg, 12, steps_noise=20)
    return cfg, sd, x, text, noise


def model(cfg, sd, precision, dev):
    m = D.DiTTO(hidden_dim=cfg.hidden_dim, num_layers=cfg.num_layers, num_heads=cfg.num_heads, time_dim=cfg.time_dim,
                text_dim=cfg.text_dim, diffusion_steps=cfg.diffusion_steps, precision=precision)
    m.load_state_dict(sd)
    return m.to(dev)


VARIANTS = [dict(method="ddim", num_steps=10, eta=0.0), dict(method="ddim", num_steps=20, eta=0.0),
            dict(method="ddim", num_steps=7, eta=0.7), dict(method="ddpm", num_steps=10),
            dict(method="ddpm", num_steps=5, schedule_scale=0.3), dict(method="ddim", num_steps=10, schedule_scale=0.3)]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("kw", VARIANTS, ids=lambda k: "-".join(f"{a}={b}" for a, b in k.items()))
@pytest.mark.parametrize("use_graph", [True, False])
def test_variant_sampling_vs_oracle(dev, setup, kw, precision, use_graph):
    cfg, sd, x, text, noise = setup
    w = 5.0                                                     # the paper's guidance scale
    want = O.sample_latents_variant(sd, cfg, text, x, noise, w, **kw)
    s = D.DiTTOSampler(model(cfg, sd, precision, dev), guidance_scale=w, **kw)
    got = s.sample_latents(text.to(dev), x_init=x.to(dev), noise=noise.to(dev), use_graph=use_graph)
    assert O.rel_l2(got.cpu(), want) <= BAR[precision]


def test_two_samplers_share_one_model(dev, setup):
    """The update table is engine state: alternating a DDIM sampler, the reference sampler and q_sample on ONE model
    must give each its own tables back."""
    cfg, sd, x, text, noise = setup
    m = model(cfg, sd, "fp32", dev)
    ref = D.DiTTOSampler(m, guidance_scale=3.0)
    ddim = D.DiTTOSampler(m, guidance_scale=3.0, method="ddim", num_steps=5)
    xd, td, zd = x.to(dev), text.to(dev), noise.to(dev)
    a1 = ref.sample_latents(td, x_init=xd, noise=zd)
    b1 = ddim.sample_latents(td, x_init=xd, noise=zd)
    t = torch.tensor([3, 17], device=dev)
    q = m.q_sample(xd, t, zd[0])
    a2 = ref.sample_latents(td, x_init=xd, noise=zd)
    b2 = ddim.sample_latents(td, x_init=xd, noise=zd)
    assert torch.equal(a1, a2) and torch.equal(b1, b2)
    assert O.rel_l2(a1.cpu(), O.sample_latents(sd, cfg, text, x, noise, 3.0)) <= 1e-4
    assert O.rel_l2(b1.cpu(), O.sample_latents_variant(sd, cfg, text, x, noise, 3.0, method="ddim", num_steps=5)) <= 1e-4
    assert O.rel_l2(q.cpu(), O.q_sample(sd, x, t.cpu(), noise[0])) <= 1e-6


def test_variant_ragged_batch(dev, setup):
    """DDIM over a sub-sequence on a mixed-length batch == every utterance sampled alone by the oracle."""
    cfg, sd, _, _, _ = setup
    g = torch.Generator().manual_seed(5)
    lengths, S_i = [24, 40, 9], [5, 9, 3]
    texts = [torch.randn(s_, cfg.text_dim, generator=g) for s_ in S_i]
    xs = [torch.randn(t_, cfg.hidden_dim, generator=g) for t_ in lengths]
    zs = [torch.randn(cfg.diffusion_steps, t_, cfg.hidden_dim, generator=g) for t_ in lengths]
    s = D.DiTTOSampler(model(cfg, sd, "fp32", dev), guidance_scale=5.0, method="ddim", num_steps=6)
    got = s.sample_latents_ragged([t.to(dev) for t in texts], lengths, x_init=[x.to(dev) for x in xs],
                                  noise=[z.to(dev) for z in zs])
    for i in range(3):
        want = O.sample_latents_variant(sd, cfg, texts[i][None], xs[i][None], zs[i][:, None], 5.0, method="ddim", num_steps=6)
        assert O.rel_l2(got[i].cpu()[None], want) <= 1e-4


def test_update_table_argument_checks(dev, setup):
    """ditto_engine_load_update_table: wrong shape is refused on the host; the C entry point reports bad sizes / a missing
    schedule through its status code (no exception crosses the ABI)."""
    import ctypes as C
    from ditto_tts_b200 import _lib
    cfg, sd, *_ = setup
    m = model(cfg, sd, "fp32", dev)
    with pytest.raises(D.DittoError):
        m.load_update_table(torch.zeros(cfg.diffusion_steps - 1, 3))
    lib = _lib.load()
    tab = torch.zeros(cfg.diffusion_steps, 3, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    eng = m.engine()
    assert lib.ditto_engine_load_update_table(eng, None, cfg.diffusion_steps, st) == -1          # DITTO_E_BADARG
    assert lib.ditto_engine_load_update_table(eng, C.c_void_p(tab.data_ptr()), 3, st) == -1
    assert b"steps" in lib.ditto_last_error()
    m2 = model(cfg, sd, "fp32", dev)                                                             # no schedule loaded yet
    assert lib.ditto_engine_load_update_table(m2.engine(), C.c_void_p(tab.data_ptr()), cfg.diffusion_steps, st) == -4  # DITTO_E_STATE
