"""CPU tests of the sampler-variant tables (ditto_tts_b200/schedules.py, host logic only) against the oracle's
step-by-step restatements, and of the reductions that pin them to the reference's sampler (SpeechGenerator.py:131-164)."""
import pytest
import torch

from ditto_tts_b200 import schedules as S
from oracle import ditto_oracle as O


@pytest.mark.parametrize("T,K", [(50, 50), (50, 25), (50, 7), (1000, 25), (1000, 50), (10, 1), (8, 8)])
def test_spaced_timesteps(T, K):
    tau = S.spaced_timesteps(T, K).tolist()
    assert tau == O.spaced_timesteps(T, K)
    assert len(tau) == K and tau[0] == T - 1 and (K == 1 or tau[-1] == 0)
    assert all(a > b for a, b in zip(tau, tau[1:]))
    if K == T:
        assert tau == list(range(T - 1, -1, -1))          # the reference's reversed(range(DIFFUSION_STEPS))


def test_spaced_timesteps_rejects():
    with pytest.raises(ValueError):
        S.spaced_timesteps(10, 11)
    with pytest.raises(ValueError):
        S.spaced_timesteps(10, 0)


@pytest.mark.parametrize("steps,rtol", [(8, 2e-5), (50, 2e-5), (1000, 5e-3)])
def test_full_ddpm_table_is_the_reference_table(steps, rtol):
    """ddpm_coef over every timestep == the three factors of SpeechGenerator.py:143-145 (incl. the t > 0 mask).
    (1000 steps: the reference's fp32 ``1 - alphas_cumprod`` loses ~3 digits near t = 0, hence the looser bound.)"""
    betas, alphas, acp = O.sampler_tables(steps)
    c = S.ddpm_coef(torch.cumprod(1.0 - betas.double(), 0), list(range(steps - 1, -1, -1))).flip(0)
    ref = torch.stack([1 / alphas.sqrt(), (1 - alphas) / (1 - acp).sqrt(), betas.sqrt() * (torch.arange(steps) > 0)], 1)
    assert torch.allclose(c.float(), ref, rtol=rtol, atol=1e-7)


@pytest.mark.parametrize("method,eta,K", [("ddim", 0.0, 50), ("ddim", 0.0, 10), ("ddim", 0.5, 25), ("ddim", 1.0, 25),
                                           ("ddpm", 0.0, 25), ("ddpm", 0.0, 5)])
def test_coef_form_equals_textbook_update(method, eta, K):
    """x_prev = c1 (x - c2 eps) + c3 z with the table rows == the oracle's textbook update, for every visited step."""
    steps = 50
    betas = O.cosine_beta_schedule(steps)
    acp = torch.cumprod(1.0 - betas.double(), 0)
    taus = S.spaced_timesteps(steps, K).tolist()
    coef = S.ddim_coef(acp, taus, eta) if method == "ddim" else S.ddpm_coef(acp, taus)
    g = torch.Generator().manual_seed(0)
    x, e, z = (torch.randn(3, 5, 16, generator=g, dtype=torch.float64) for _ in range(3))
    for i, t in enumerate(taus):
        last = i == K - 1
        want = O.variant_update(x, e, z, float(acp[t]), 1.0 if last else float(acp[taus[i + 1]]), method, eta, last)
        got = coef[i, 0] * (x - coef[i, 1] * e) + coef[i, 2] * z
        assert torch.allclose(got, want, rtol=1e-9, atol=1e-9 * float(coef[i, 0]))
    tab = S.coef_table(steps, taus, coef)
    assert tab.shape == (steps, 3) and tab.dtype == torch.float32
    assert torch.equal(tab[taus], coef.float())


def test_ddim_eta1_last_step_is_x0_prediction():
    acp = torch.cumprod(1.0 - O.cosine_beta_schedule(50).double(), 0)
    c = S.ddim_coef(acp, [0], 1.0)[0]
    assert c[2] == 0.0 and abs(float(c[0]) - float(1 / acp[0].sqrt())) < 1e-12
    assert abs(float(c[1]) - float((1 - acp[0]).sqrt())) < 1e-12


@pytest.mark.parametrize("steps", [50, 1000])
def test_shifted_cosine_schedule(steps):
    ref = O.cosine_beta_schedule(steps)
    assert torch.allclose(S.shifted_cosine_betas(steps, 1.0), ref, rtol=1e-4, atol=2e-6)   # scale 1 == DiTTO.py:96-104 (fp32 there)
    for scale in (0.3, 0.5, 2.0):
        b = S.shifted_cosine_betas(steps, scale)
        assert torch.allclose(b, O.shifted_cosine_betas(steps, scale), atol=1e-7)
        assert float(b.min()) >= 1e-4 - 1e-8 and float(b.max()) <= 0.9999 + 1e-7
    # a scale below 1 lowers the SNR of every level: abar' < abar
    lo = torch.cumprod(1 - S.shifted_cosine_betas(steps, 0.3).double(), 0)
    hi = torch.cumprod(1 - ref.double(), 0)
    assert bool((lo[: steps // 2] < hi[: steps // 2]).all())


def test_oracle_variant_sampler_reduces_to_reference_sampler():
    """All timesteps + ancestral update: sample_latents_variant == sample_latents (the reference restatement)."""
    cfg = O.OracleConfig(64, 1, 1, 16, 64, 6)
    sd = O.make_state_dict(cfg, 3)
    x, text, noise = O.make_inputs(2, 12, 4, cfg, 4, steps_noise=6)
    a = O.sample_latents(sd, cfg, text, x, noise, 2.0)
    b = O.sample_latents_variant(sd, cfg, text, x, noise, 2.0, method="ddpm")
    assert O.rel_l2(b, a) <= 1e-5
