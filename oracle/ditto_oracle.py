"""CPU oracle for the DiTTo-TTS DiT denoiser hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32 or fp64) restatement of the reference
algorithm.  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product path (``ditto_tts_b200``) never imports anything from ``oracle/`` and
fails loudly when the CUDA library is missing.

Parity status: PINNED.  The reference has no tests or golden vectors of its own
(SURVEY.md section 4), so the pin is "outputs of the reference itself run here":
``tests/golden/make_golden.py`` imports the unmodified reference modules from
``/root/reference/src`` (NAC stubbed), runs them on seeded
weights/inputs and stores the results under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this restatement against those fixtures
(fp32: rel-L2 <= 2e-6, i.e. re-association noise only).

Every function cites the reference file:line it follows
(paths relative to the reference repo root).

All tensors: x [n,T,H], text_emb [n,S,text_dim], t [n] int64.  ``sd`` is a
state_dict with the reference's key names.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass(frozen=True)
class OracleConfig:
    """Shapes of one DiTTO instance.  Defaults = ConfigDiTTO, src/utils/Config.py:109-116."""

    hidden_dim: int = 768
    num_layers: int = 5
    num_heads: int = 1
    time_dim: int = 256
    text_dim: int = 768
    diffusion_steps: int = 1000

    @property
    def head_dim(self) -> int:
        return self.hidden_dim // self.num_heads


# --------------------------------------------------------------------------
# schedule  (src/model/DiTTO.py:96-104, src/model/SpeechGenerator.py:70-72)
# --------------------------------------------------------------------------
def cosine_beta_schedule(timesteps: int, s: float = 0.008) -> Tensor:
    """Returns the clipped *betas* (the reference's name notwithstanding).  DiTTO.py:96-104."""
    steps = timesteps + 1
    x = torch.linspace(0, timesteps, steps)
    alphas_cumprod = torch.cos(((x / timesteps) + s) / (1 + s) * torch.pi * 0.5) ** 2
    alphas_cumprod = alphas_cumprod / alphas_cumprod[0]
    betas = 1 - (alphas_cumprod[1:] / alphas_cumprod[:-1])
    return torch.clip(betas, 0.0001, 0.9999)


def sampler_tables(timesteps: int):
    """betas, alphas, alphas_cumprod as the sampler builds them.  SpeechGenerator.py:70-72."""
    betas = cosine_beta_schedule(timesteps)
    alphas = 1.0 - betas
    alphas_cumprod = torch.cumprod(alphas, dim=0)
    return betas, alphas, alphas_cumprod


# --------------------------------------------------------------------------
# components  (src/components/DiT.py)
# --------------------------------------------------------------------------
def rotary_angles(seq_len: int, head_dim: int, dtype=torch.float32) -> Tensor:
    """RotaryEmbedding.__init__ + forward: [T, head_dim] angles = cat(freqs, freqs).  DiT.py:46-59."""
    inv_freq = 1.0 / (10000 ** (torch.arange(0, head_dim, 2).float() / head_dim))  # DiT.py:49 (fp32 buffer)
    inv_freq = inv_freq.to(dtype)
    t = torch.arange(seq_len).to(dtype)                                            # DiT.py:57
    freqs = torch.einsum("i,j->ij", t, inv_freq)                                   # DiT.py:58
    return torch.cat((freqs, freqs), dim=-1)                                       # DiT.py:59


def rotate_half(x: Tensor) -> Tensor:
    """DiT.py:52-54."""
    x1, x2 = x.chunk(2, dim=-1)
    return torch.cat((-x2, x1), dim=-1)


def apply_rope(pos: Tensor, t: Tensor) -> Tensor:
    """t: [n, T, heads, d]; pos: [T, d].  DiT.py:61-72."""
    pos = pos.unsqueeze(0).unsqueeze(2)
    return t * pos.cos() + rotate_half(t) * pos.sin()


def global_adaln(sd: Dict[str, Tensor], x: Tensor, time_emb: Tensor, text_emb: Tensor) -> Tensor:
    """GlobalAdaLN.forward.  DiT.py:25-40."""
    text_mean = torch.mean(text_emb, dim=1)                                                       # :27
    tm = F.linear(F.silu(time_emb), sd["ada_ln.time_mlp.1.weight"], sd["ada_ln.time_mlp.1.bias"])  # :30
    xm = F.linear(F.silu(text_mean), sd["ada_ln.text_mlp.1.weight"], sd["ada_ln.text_mlp.1.bias"])  # :31
    time_scale, time_shift = tm.chunk(2, dim=-1)
    text_scale, text_shift = xm.chunk(2, dim=-1)
    scale = 1 + time_scale + text_scale                                                           # :34
    shift = time_shift + text_shift                                                               # :35
    x = F.layer_norm(x, (x.shape[-1],), None, None, 1e-5)                                         # :38 (no affine)
    return x * scale.unsqueeze(1) + shift.unsqueeze(1)                                            # :39


def self_attention(sd, prefix: str, x: Tensor, rotary_pos: Tensor, num_heads: int) -> Tensor:
    """LN1 -> manual q/k/v from attn.in_proj -> RoPE(q,k) -> softmax(qk^T/sqrt(d))v -> +residual.
    No out_proj, no mask.  DiT.py:103-139."""
    n, T, H = x.shape
    d = H // num_heads
    residual = x
    u = F.layer_norm(x, (H,), sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], 1e-5)  # :105
    w = sd[prefix + "attn.in_proj_weight"]
    b = sd[prefix + "attn.in_proj_bias"]
    q = F.linear(u, w[:H], b[:H])                     # :112
    k = F.linear(u, w[H:2 * H], b[H:2 * H])           # :113
    v = F.linear(u, w[2 * H:], b[2 * H:])             # :114
    q = q.reshape(n, T, num_heads, d)                 # :117-119 (einops 'b n (h d) -> b n h d')
    k = k.reshape(n, T, num_heads, d)
    v = v.reshape(n, T, num_heads, d)
    q = apply_rope(rotary_pos, q)                     # :122
    k = apply_rope(rotary_pos, k)                     # :123
    q, k, v = (z.permute(0, 2, 1, 3) for z in (q, k, v))                # :126-128
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(d)        # :131-132
    attn = torch.softmax(scores, dim=-1)                                # :133
    out = torch.matmul(attn, v)                                         # :134
    out = out.permute(0, 2, 1, 3).reshape(n, T, H)                      # :137-138
    return out + residual                                               # :139


def cross_attention(sd, prefix: str, x: Tensor, text_emb: Tensor, num_heads: int) -> Tensor:
    """LN2 -> nn.MultiheadAttention(q=x, k=v=text) math path (eval: dropout off) -> +residual.
    DiT.py:141-148; torch/nn/functional.py multi_head_attention_forward, need_weights branch:
    q is pre-scaled by sqrt(1/d), softmax over S, out_proj applied; the head-averaged
    weights are computed and dropped by the caller ([0])."""
    n, T, H = x.shape
    S = text_emb.shape[1]
    d = H // num_heads
    residual = x
    u = F.layer_norm(x, (H,), sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], 1e-5)  # :143
    w = sd[prefix + "cross_attn.in_proj_weight"]
    b = sd[prefix + "cross_attn.in_proj_bias"]
    q = F.linear(u, w[:H], b[:H])
    k = F.linear(text_emb, w[H:2 * H], b[H:2 * H])
    v = F.linear(text_emb, w[2 * H:], b[2 * H:])
    q = q.reshape(n, T, num_heads, d).permute(0, 2, 1, 3)
    k = k.reshape(n, S, num_heads, d).permute(0, 2, 1, 3)
    v = v.reshape(n, S, num_heads, d).permute(0, 2, 1, 3)
    q = q * math.sqrt(1.0 / d)
    attn = torch.softmax(torch.matmul(q, k.transpose(-2, -1)), dim=-1)
    out = torch.matmul(attn, v).permute(0, 2, 1, 3).reshape(n, T, H)
    out = F.linear(out, sd[prefix + "cross_attn.out_proj.weight"], sd[prefix + "cross_attn.out_proj.bias"])
    return out + residual                                                                     # :148


def gated_mlp(sd, prefix: str, x: Tensor) -> Tensor:
    """LN3 -> fc2(GELU_erf(fc1(h)) * sigmoid(gate(h))) + residual.  DiT.py:150-155."""
    H = x.shape[-1]
    residual = x
    u = F.layer_norm(x, (H,), sd[prefix + "norm3.weight"], sd[prefix + "norm3.bias"], 1e-5)  # :152
    a = F.gelu(F.linear(u, sd[prefix + "mlp_fc1.weight"], sd[prefix + "mlp_fc1.bias"]))       # :153 (exact erf)
    g = torch.sigmoid(F.linear(u, sd[prefix + "gate.weight"], sd[prefix + "gate.bias"]))      # :154
    return F.linear(a * g, sd[prefix + "mlp_fc2.weight"], sd[prefix + "mlp_fc2.bias"]) + residual  # :155


def dit_block(sd, i: int, x: Tensor, text_emb: Tensor, rotary_pos: Tensor, num_heads: int) -> Tensor:
    """DiT.forward; time_emb is accepted and ignored by the reference.  DiT.py:100-157."""
    p = f"blocks.{i}."
    x = self_attention(sd, p, x, rotary_pos, num_heads)
    x = cross_attention(sd, p, x, text_emb, num_heads)
    return gated_mlp(sd, p, x)


# --------------------------------------------------------------------------
# model  (src/model/DiTTO.py)
# --------------------------------------------------------------------------
def time_embedding(sd, t: Tensor) -> Tensor:
    """t_embedding lookup + time_embed MLP.  DiTTO.py:75-76 (params :37-44)."""
    e = sd["t_embedding.weight"][t]
    e = F.linear(e, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    e = F.silu(e)
    return F.linear(e, sd["time_embed.2.weight"], sd["time_embed.2.bias"])


def ditto_forward(sd: Dict[str, Tensor], cfg: OracleConfig, x: Tensor, text_emb: Tensor, t: Tensor,
                  taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """DiTTO.forward.  DiTTO.py:66-94.  ``taps`` (optional dict) receives intermediates for kernel tests."""
    dtype = x.dtype
    te = time_embedding(sd, t)                                             # :75-76
    rotary_pos = rotary_angles(x.shape[1], cfg.head_dim, dtype)            # :79-80
    x_skip = F.linear(x, sd["proj_in.weight"], sd["proj_in.bias"])         # :83
    h = global_adaln(sd, x, te, text_emb)                                  # :86
    if taps is not None:
        taps["time_emb"] = te
        taps["x_skip"] = x_skip
        taps["adaln"] = h
    for i in range(cfg.num_layers):                                        # :89-90
        h = dit_block(sd, i, h, text_emb, rotary_pos, cfg.num_heads)
        if taps is not None:
            taps[f"block{i}"] = h
    out = F.linear(h, sd["proj_out.weight"], sd["proj_out.bias"])          # :93
    return x_skip + out                                                    # :94


def q_sample(sd, x_start: Tensor, t: Tensor, noise: Tensor) -> Tensor:
    """DiTTO.q_sample incl. the reference quirk: the ``alphas_cumprod`` buffer holds the clipped
    betas (DiTTO.py:63-64 registers cosine_beta_schedule's return value).  DiTTO.py:106-126."""
    ac = sd["alphas_cumprod"][t.long()]
    a = (ac ** 0.5).reshape(-1, 1, 1)
    b = ((1 - ac) ** 0.5).reshape(-1, 1, 1)
    return a * x_start + b * noise


# --------------------------------------------------------------------------
# sampler  (src/model/SpeechGenerator.py:131-164), plus the CFG extension
# --------------------------------------------------------------------------
def predict_noise(sd, cfg: OracleConfig, x: Tensor, text_emb: Tensor, t: Tensor,
                  guidance_scale: Optional[float] = None, null_text_emb: Optional[Tensor] = None) -> Tensor:
    """eps_hat.  Without guidance this is the reference call SpeechGenerator.py:135.
    EXTENSION (not in the reference): classifier-free guidance
    eps = eps_u + w (eps_c - eps_u), unconditional branch = zero text embedding unless given."""
    eps_c = ditto_forward(sd, cfg, x, text_emb, t)
    if guidance_scale is None:
        return eps_c
    null = torch.zeros_like(text_emb) if null_text_emb is None else null_text_emb
    eps_u = ditto_forward(sd, cfg, x, null, t)
    return eps_u + guidance_scale * (eps_c - eps_u)


def p_sample_update(x: Tensor, noise_pred: Tensor, noise: Tensor, t: Tensor,
                    betas: Tensor, alphas: Tensor, alphas_cumprod: Tensor) -> Tensor:
    """The DDPM ancestral update, term by term as the reference writes it.  SpeechGenerator.py:137-147."""
    beta_t = betas[t].view(-1, 1, 1)
    alpha_t = alphas[t].view(-1, 1, 1)
    alpha_cumprod_t = alphas_cumprod[t].view(-1, 1, 1)
    mask = (t > 0).to(x.dtype).view(-1, 1, 1)
    return (1 / torch.sqrt(alpha_t)) * (
        x - (1 - alpha_t) / torch.sqrt(1 - alpha_cumprod_t) * noise_pred
    ) + mask * torch.sqrt(beta_t) * noise


def sample_latents(sd, cfg: OracleConfig, text_emb: Tensor, x_init: Tensor, noise: Tensor,
                   guidance_scale: Optional[float] = None, record: Optional[List[Tensor]] = None) -> Tensor:
    """__sample_latents: for t = steps-1 .. 0: x = p_sample(x, t).  SpeechGenerator.py:150-164.
    ``noise`` [steps, B, T, H] replaces the reference's in-loop randn_like (index = t_val), so that
    CPU and CUDA runs see the same draws.  ``record`` collects the per-step eps_hat."""
    steps = cfg.diffusion_steps
    betas, alphas, alphas_cumprod = (z.to(x_init.dtype) for z in sampler_tables(steps))
    x = x_init
    for t_val in reversed(range(steps)):                                       # :161
        t = torch.full((x.shape[0],), t_val, dtype=torch.long)                 # :162
        eps = predict_noise(sd, cfg, x, text_emb, t, guidance_scale)           # :135
        if record is not None:
            record.append(eps)
        x = p_sample_update(x, eps, noise[t_val], t, betas, alphas, alphas_cumprod)  # :137-147
    return x


# --------------------------------------------------------------------------
# sampler variants.  EXTENSIONS: the reference has one sampler (above); these are the textbook
# forms (DDIM: Song et al. 2021 eq. 12; sub-sequence DDPM: Nichol & Dhariwal 2021 sec. 4; SNR-shifted cosine schedule:
# Hoogeboom et al. 2023 sec. 3.1), written step by step in float64 and NOT through the c1/c2/c3 rewrite the product uses.
# "Parity unpinned" for these: there is no reference implementation to mint goldens from; the pin is that with all
# timesteps / eta-free DDPM they reduce to p_sample_update above (tests/test_schedules.py).
# --------------------------------------------------------------------------
def spaced_timesteps(train_steps: int, num_steps: int) -> List[int]:
    """num_steps timesteps from train_steps-1 down to 0, evenly spaced, rounded half up (integer arithmetic)."""
    if num_steps == 1:
        return [train_steps - 1]
    d = num_steps - 1
    return [(2 * (train_steps - 1) * (d - i) + d) // (2 * d) for i in range(num_steps)]


def shifted_cosine_betas(timesteps: int, scale: float, s: float = 0.008) -> Tensor:
    """cosine_beta_schedule with SNR(t) multiplied by scale^2; same clip as DiTTO.py:104."""
    acp = [math.cos(((i / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2 for i in range(timesteps + 1)]
    acp = [a / acp[0] for a in acp]
    out = []
    for a in acp:
        snr_num = scale * scale * a
        out.append(snr_num / (snr_num + (1.0 - a)))
    betas = [min(max(1.0 - out[i + 1] / out[i], 0.0001), 0.9999) for i in range(timesteps)]
    return torch.tensor(betas, dtype=torch.float64).to(torch.float32)


def variant_update(x: Tensor, eps: Tensor, z: Tensor, acp_t: float, acp_prev: float, method: str, eta: float,
                   last: bool) -> Tensor:
    """One reverse step from abar_t to abar_prev in float64 (result cast back to x.dtype)."""
    xd, ed, zd = x.double(), eps.double(), z.double()
    if method == "ddim":
        x0 = (xd - math.sqrt(1.0 - acp_t) * ed) / math.sqrt(acp_t)
        sigma = eta * math.sqrt((1.0 - acp_prev) / (1.0 - acp_t)) * math.sqrt(max(1.0 - acp_t / acp_prev, 0.0))
        out = math.sqrt(acp_prev) * x0 + math.sqrt(max(1.0 - acp_prev - sigma * sigma, 0.0)) * ed + sigma * zd
    else:  # ancestral, sigma^2 = beta' (the reference's variance choice, SpeechGenerator.py:143-145)
        alpha = acp_t / acp_prev
        out = (xd - (1.0 - alpha) / math.sqrt(1.0 - acp_t) * ed) / math.sqrt(alpha)
        if not last:
            out = out + math.sqrt(1.0 - alpha) * zd
    return out.to(x.dtype)


def sample_latents_variant(sd, cfg: OracleConfig, text_emb: Tensor, x_init: Tensor, noise: Tensor,
                           guidance_scale: Optional[float] = None, method: str = "ddpm", num_steps: Optional[int] = None,
                           eta: float = 0.0, schedule_scale: Optional[float] = None,
                           record: Optional[List[Tensor]] = None) -> Tensor:
    """sample_latents over a sub-sequence of timesteps / with DDIM / with the shifted schedule.  noise index = t_val."""
    steps = cfg.diffusion_steps
    betas = cosine_beta_schedule(steps) if schedule_scale is None else shifted_cosine_betas(steps, schedule_scale)
    acp = torch.cumprod(1.0 - betas.double(), dim=0).tolist()
    taus = spaced_timesteps(steps, steps if num_steps is None else num_steps)
    x = x_init
    for i, t_val in enumerate(taus):
        t = torch.full((x.shape[0],), t_val, dtype=torch.long)
        eps = predict_noise(sd, cfg, x, text_emb, t, guidance_scale)
        if record is not None:
            record.append(eps)
        last = i == len(taus) - 1
        acp_prev = 1.0 if last else acp[taus[i + 1]]
        x = variant_update(x, eps, noise[t_val], acp[t_val], acp_prev, method, eta, last)
    return x


# --------------------------------------------------------------------------
# deterministic synthetic weights (used by tests, bench and the golden script alike)
# --------------------------------------------------------------------------
def state_dict_keys(cfg: OracleConfig):
    """(key, shape, kind) for every tensor of the reference DiTTO state_dict that the hot path reads
    (SURVEY.md section 8b), plus the dead self-attn out_proj so that load_state_dict(strict) works."""
    H, Td, Xd, L, St = cfg.hidden_dim, cfg.time_dim, cfg.text_dim, cfg.num_layers, cfg.diffusion_steps
    ks = [("t_embedding.weight", (St, Td), "emb"),
          ("time_embed.0.weight", (Td, Td), "w"), ("time_embed.0.bias", (Td,), "b"),
          ("time_embed.2.weight", (Td, Td), "w"), ("time_embed.2.bias", (Td,), "b"),
          ("ada_ln.time_mlp.1.weight", (2 * H, Td), "w"), ("ada_ln.time_mlp.1.bias", (2 * H,), "b"),
          ("ada_ln.text_mlp.1.weight", (2 * H, Xd), "w"), ("ada_ln.text_mlp.1.bias", (2 * H,), "b")]
    for i in range(L):
        p = f"blocks.{i}."
        ks += [(p + "norm1.weight", (H,), "g"), (p + "norm1.bias", (H,), "b"),
               (p + "attn.in_proj_weight", (3 * H, H), "w"), (p + "attn.in_proj_bias", (3 * H,), "b"),
               (p + "attn.out_proj.weight", (H, H), "w"), (p + "attn.out_proj.bias", (H,), "b"),
               (p + "norm2.weight", (H,), "g"), (p + "norm2.bias", (H,), "b"),
               (p + "cross_attn.in_proj_weight", (3 * H, H), "w"), (p + "cross_attn.in_proj_bias", (3 * H,), "b"),
               (p + "cross_attn.out_proj.weight", (H, H), "w"), (p + "cross_attn.out_proj.bias", (H,), "b"),
               (p + "norm3.weight", (H,), "g"), (p + "norm3.bias", (H,), "b"),
               (p + "mlp_fc1.weight", (4 * H, H), "w"), (p + "mlp_fc1.bias", (4 * H,), "b"),
               (p + "gate.weight", (4 * H, H), "w"), (p + "gate.bias", (4 * H,), "b"),
               (p + "mlp_fc2.weight", (H, 4 * H), "w"), (p + "mlp_fc2.bias", (H,), "b")]
    ks += [("proj_in.weight", (H, H), "w"), ("proj_in.bias", (H,), "b"),
           ("proj_out.weight", (H, H), "w"), ("proj_out.bias", (H,), "b")]
    return ks


def make_state_dict(cfg: OracleConfig, seed: int = 0) -> Dict[str, Tensor]:
    """Seeded random-init weights with torch-default-like scales (uniform +-1/sqrt(fan_in) for
    matrices, N(0,1) embedding) but NON-trivial biases and LayerNorm affine parameters, so that a
    dropped bias / gamma cannot hide.  Pure CPU torch.Generator => identical on every box."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for key, shape, kind in state_dict_keys(cfg):
        if kind == "emb":
            v = torch.randn(shape, generator=g)
        elif kind == "w":
            bound = 1.0 / math.sqrt(shape[1])
            v = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "g":
            v = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            v = 0.05 * torch.randn(shape, generator=g)
        sd[key] = v
    sd["alphas_cumprod"] = cosine_beta_schedule(cfg.diffusion_steps)        # DiTTO.py:63-64 (betas!)
    inv = 1.0 / (10000 ** (torch.arange(0, cfg.head_dim, 2).float() / cfg.head_dim))
    sd["rotary.inv_freq"] = inv.clone()
    for i in range(cfg.num_layers):
        sd[f"blocks.{i}.rotary.inv_freq"] = inv.clone()
    return sd


def make_inputs(B: int, T: int, S: int, cfg: OracleConfig, seed: int = 1, steps_noise: int = 0):
    """x_T ~ N(0,1) [B,T,H], text_emb ~ N(0,1) [B,S,text_dim], optional noise [steps,B,T,H].
    SURVEY.md section 8d 'Synthetic inputs'."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, cfg.hidden_dim, generator=g)
    text = torch.randn(B, S, cfg.text_dim, generator=g)
    noise = torch.randn(steps_noise, B, T, cfg.hidden_dim, generator=g) if steps_noise else None
    return x, text, noise


# --------------------------------------------------------------------------
# hand-off steps either side of the loop
# --------------------------------------------------------------------------
def make_codebook(codebook_size: int, latent_dim: int, seed: int = 0) -> Tensor:
    """A seeded codebook with the reference's init distribution (xavier_uniform_, VectorQuantizer.py:19-20),
    drawn from a private CPU generator so that it regenerates identically on every box."""
    g = torch.Generator().manual_seed(seed)
    bound = math.sqrt(6.0 / (codebook_size + latent_dim))
    return (torch.rand(codebook_size, latent_dim, generator=g) * 2 - 1) * bound


def vq_distances(codebook: Tensor, latents_flat: Tensor) -> Tensor:
    """|z|^2 - 2 z C^T + |c|^2, in the reference's association.  VectorQuantizer.py:34-38."""
    return (
        torch.sum(latents_flat ** 2, dim=1, keepdim=True)
        - 2 * torch.matmul(latents_flat, codebook.T)
        + torch.sum(codebook ** 2, dim=1)
    )


def vq_indices(codebook: Tensor, latents: Tensor) -> Tensor:
    """VectorQuantizer.forward: latents [B,C,T,D] -> argmin indices int64 [B,C,T].  VectorQuantizer.py:22-43."""
    batch_size, num_channels, num_frames, latent_dim = latents.shape
    flat = latents.reshape(-1, latent_dim)                                     # :31
    indices = torch.argmin(vq_distances(codebook, flat), dim=-1)               # :40
    return indices.view(batch_size, num_channels, num_frames)                  # :41


def latents_to_codes(codebook: Tensor, latents: Tensor, channels: int = 2) -> Tensor:
    """The step right after the sampling loop.  SpeechGenerator.py:117-118."""
    return vq_indices(codebook, latents.unsqueeze(1).repeat(1, channels, 1, 1))


def vq_near_tie(codebook: Tensor, latents_flat: Tensor, got: Tensor, want: Tensor, rel_tol: float = 1e-5) -> Tensor:
    """For rows where ``got`` != ``want``: True where the two candidates' distances (fp64) differ by less than
    rel_tol * |distance| -- i.e. the disagreement is an fp32 re-association tie, not a wrong winner."""
    d = vq_distances(codebook.double(), latents_flat.double())
    dg = d.gather(1, got.reshape(-1, 1)).squeeze(1)
    dw = d.gather(1, want.reshape(-1, 1)).squeeze(1)
    return (dg - dw).abs() <= rel_tol * dw.abs().clamp_min(1e-30)


def pool_latents(audio_latents: Tensor, max_length: int) -> Tensor:
    """audio_latents[:, :, :max_length].mean(dim=1).  TrainDiTTO.py:70-71 (validation: :113-114)."""
    return audio_latents[:, :, :max_length].mean(dim=1)


def mse_loss(pred: Tensor, target: Tensor) -> Tensor:
    """nn.MSELoss() (mean reduction).  TrainDiTTO.py:51,87,126."""
    return F.mse_loss(pred, target)


def validation_step(sd, cfg: OracleConfig, audio_latents: Tensor, text_embeddings: Tensor, t: Tensor, noise: Tensor,
                    max_length: int = 1024):
    """One iteration of the validation loop after the NAC encoder.  TrainDiTTO.py:113-127."""
    lat = pool_latents(audio_latents, max_length)                              # :113-114
    text = text_embeddings[:, :lat.size(1)]                                    # :115
    noisy = q_sample(sd, lat, t, noise)                                        # :123
    pred = ditto_forward(sd, cfg, noisy, text, t)                              # :124
    return mse_loss(pred, noise), pred                                         # :126


def rel_l2(a: Tensor, b: Tensor) -> float:
    """||a-b||_2 / ||b||_2 in fp64 (b = reference)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def forward_flops(T: int, S: int, cfg: OracleConfig) -> float:
    """Algorithmic flops per sequence per forward, SURVEY.md section 8a:
    L(34 T H^2 + 4 T^2 H + 4 T S H) + 4 T H^2 (dead self-attn out_proj excluded)."""
    H, L = cfg.hidden_dim, cfg.num_layers
    return L * (34.0 * T * H * H + 4.0 * T * T * H + 4.0 * T * S * H) + 4.0 * T * H * H


# ------------------------------------------------------------------------------------------------
# Noise stream of the fused update kernel (extension: the reference calls torch.randn_like, SpeechGenerator.py:145,
# whose CPU and CUDA streams differ anyway).  Philox4x32-10 as published (Salmon, Moraes, Dror, Shaw: "Parallel random
# numbers: as easy as 1, 2, 3", SC'11; Random123 known-answer vectors in tests/test_oracle_golden.py) + Box-Muller,
# restated in numpy exactly as csrc/elementwise.cu:philox_normal4 forms it: counter = (vector index lo, hi, draw lo, hi),
# key = (seed lo, hi); u = (r >> 8) 2^-24 + 2^-25; z = sqrt(-2 ln u0) (cos, sin)(2 pi (r1 >> 8) 2^-24), two pairs per vector.
# ------------------------------------------------------------------------------------------------
def philox4x32_10(counter, key):
    """counter [..., 4] uint32, key [..., 2] uint32 -> [..., 4] uint32."""
    import numpy as np
    c = [np.asarray(counter[..., i], dtype=np.uint64) for i in range(4)]
    k = [np.asarray(key[..., i], dtype=np.uint64) for i in range(2)]
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k[0], lo1, hi0 ^ c[3] ^ k[1], lo0]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return np.stack(c, axis=-1).astype(np.uint32)


def philox_normal(seed: int, draw: int, n: int, elem_offset: int = 0):
    """The n normals the CUDA update kernel draws for elements elem_offset .. elem_offset + n (float64 arithmetic)."""
    import numpy as np
    assert elem_offset % 4 == 0
    nv = (n + 3) // 4
    v = np.arange(nv, dtype=np.uint64) + np.uint64(elem_offset // 4)
    ctr = np.stack([v & np.uint64(0xFFFFFFFF), v >> np.uint64(32), np.full(nv, draw & 0xFFFFFFFF, np.uint64),
                    np.full(nv, (draw >> 32) & 0xFFFFFFFF, np.uint64)], axis=-1).astype(np.uint32)
    key = np.broadcast_to(np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32), (nv, 2))
    r = philox4x32_10(ctr, key).astype(np.float64)
    r = np.floor(r / 256.0)
    u0, u1 = r[:, 0] * 2.0 ** -24 + 2.0 ** -25, r[:, 2] * 2.0 ** -24 + 2.0 ** -25
    a0, a1 = r[:, 1] * (2.0 * np.pi * 2.0 ** -24), r[:, 3] * (2.0 * np.pi * 2.0 ** -24)
    m0, m1 = np.sqrt(-2.0 * np.log(u0)), np.sqrt(-2.0 * np.log(u1))
    z = np.stack([m0 * np.cos(a0), m0 * np.sin(a0), m1 * np.cos(a1), m1 * np.sin(a1)], axis=-1).reshape(-1)
    return torch.from_numpy(z[:n].astype(np.float32))
