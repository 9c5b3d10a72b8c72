"""Hand-off steps either side of the sampling loop, on the device.

* ``VectorQuantizer`` -- drop-in for ``components.VectorQuantizer.VectorQuantizer`` (src/components/VectorQuantizer.py:4-43):
  same constructor, same ``codebook`` parameter (state_dict key ``codebook``; ``nac.vector_quantizer.codebook`` in a
  reference checkpoint), ``forward(latents[B,C,T,D]) -> int64 [B,C,T]``.
* ``latents_to_codes`` -- the step right after the loop (SpeechGenerator.py:117-118): ``latents.unsqueeze(1).repeat(1,2,1,1)``
  through the quantiser, without quantising the duplicated channel twice.
* ``pool_latents`` -- ``audio_latents[:, :, :max_length].mean(dim=1)`` (TrainDiTTO.py:70-71).
* ``mse_loss`` / ``validation_step`` -- q_sample -> forward -> nn.MSELoss (TrainDiTTO.py:116-127), inference only.

All arithmetic runs in libditto_b200.so (csrc/vq.cu); CPU tensors raise, there is no fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from ._lib import DittoError
from .model import _need_cuda_f32, _ptr, _stream

__all__ = ["VectorQuantizer", "latents_to_codes", "pool_latents", "mse_loss", "validation_step"]


class VectorQuantizer(nn.Module):
    """Latents -> codebook indices (reference: VectorQuantizer.py:4-43).  Same init: randn then xavier_uniform_."""

    def __init__(self, codebook_size, latent_dim):
        super().__init__()
        self.codebook_size = codebook_size
        self.latent_dim = latent_dim
        self.codebook = nn.Parameter(torch.randn(codebook_size, latent_dim))
        nn.init.xavier_uniform_(self.codebook)
        self._sq = None
        self._sq_key = None

    def _sqnorm(self) -> torch.Tensor:
        """|c|^2 per code (VectorQuantizer.py:37), cached until the codebook changes."""
        cb = self.codebook
        if not cb.is_cuda:
            raise DittoError("VectorQuantizer.codebook is on the CPU: move the module to a B200; there is no CPU fallback")
        key = (cb.data_ptr(), cb._version)
        if self._sq is None or self._sq_key != key:
            if cb.dtype != torch.float32 or not cb.is_contiguous():
                raise DittoError("codebook must be a contiguous float32 tensor")
            sq = torch.empty((self.codebook_size,), dtype=torch.float32, device=cb.device)
            with torch.cuda.device(cb.device):
                _lib.check(_lib.load().ditto_vq_code_sqnorm(_ptr(cb.detach()), self.codebook_size, self.latent_dim, _ptr(sq),
                                                            _stream()), "ditto_vq_code_sqnorm")
            self._sq, self._sq_key = sq, key
        return self._sq

    @torch.no_grad()
    def encode(self, latents: torch.Tensor, repeat_channels: int = 1) -> torch.Tensor:
        """latents [B, T, D] -> indices int64 [B, repeat_channels, T] (every channel carries the same indices)."""
        latents = _need_cuda_f32("latents", latents)
        if latents.dim() != 3 or latents.shape[-1] != self.latent_dim:
            raise DittoError(f"expected latents [B, T, {self.latent_dim}], got {tuple(latents.shape)}")
        B, T, D = latents.shape
        sq = self._sqnorm()
        out = torch.empty((B, repeat_channels, T), dtype=torch.int64, device=latents.device)
        with torch.cuda.device(latents.device):
            _lib.check(_lib.load().ditto_vq_encode(_ptr(latents), B, T, D, _ptr(self.codebook.detach()), self.codebook_size,
                                                   _ptr(sq), repeat_channels, _ptr(out), _stream()), "ditto_vq_encode")
        return out

    @torch.no_grad()
    def forward(self, latents: torch.Tensor) -> torch.Tensor:
        """latents [B, C, T, D] -> indices [B, C, T] (VectorQuantizer.py:22-43)."""
        latents = _need_cuda_f32("latents", latents)
        if latents.dim() != 4:
            raise DittoError("expected latents [batch, channels, frames, latent_dim]")
        B, Cn, T, D = latents.shape
        return self.encode(latents.view(B * Cn, T, D)).view(B, Cn, T)


def latents_to_codes(vq: VectorQuantizer, latents: torch.Tensor, channels: int = 2) -> torch.Tensor:
    """``vq(latents.unsqueeze(1).repeat(1, channels, 1, 1))`` (SpeechGenerator.py:117-118): [B,T,D] -> int64 [B,channels,T].
    The duplicated channel is quantised once and its indices written ``channels`` times."""
    return vq.encode(latents, repeat_channels=channels)


@torch.no_grad()
def pool_latents(audio_latents: torch.Tensor, max_length: int) -> torch.Tensor:
    """``audio_latents[:, :, :max_length].mean(dim=1)`` (TrainDiTTO.py:70-71): [B,C,T,D] -> [B,min(T,max_length),D]."""
    audio_latents = _need_cuda_f32("audio_latents", audio_latents)
    if audio_latents.dim() != 4:
        raise DittoError("expected audio_latents [batch, channels, frames, dim]")
    B, Cn, T, D = audio_latents.shape
    out = torch.empty((B, min(T, int(max_length)), D), dtype=torch.float32, device=audio_latents.device)
    with torch.cuda.device(audio_latents.device):
        _lib.check(_lib.load().ditto_pool_latents(_ptr(audio_latents), B, Cn, T, D, int(max_length), _ptr(out), _stream()),
                   "ditto_pool_latents")
    return out


@torch.no_grad()
def mse_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """``nn.MSELoss()(pred, target)`` (TrainDiTTO.py:51): 0-dim fp32 CUDA tensor, deterministic reduction."""
    pred = _need_cuda_f32("pred", pred)
    target = _need_cuda_f32("target", target)
    if pred.shape != target.shape:
        raise DittoError("mse_loss: shapes differ")
    lib = _lib.load()
    ws = torch.empty((lib.ditto_mse_workspace_bytes() // 8,), dtype=torch.float64, device=pred.device)
    out = torch.empty((1,), dtype=torch.float32, device=pred.device)
    with torch.cuda.device(pred.device):
        _lib.check(lib.ditto_mse_loss(_ptr(pred), _ptr(target), pred.numel(), _ptr(out), _ptr(ws), ws.numel() * 8, _stream()),
                   "ditto_mse_loss")
    return out[0]


@torch.no_grad()
def validation_step(model, audio_latents: torch.Tensor, text_embeddings: torch.Tensor, t: torch.Tensor,
                    noise: torch.Tensor = None, max_length: int = 1024):
    """One iteration of the reference's validation loop after the NAC (TrainDiTTO.py:113-127): pool the encoder
    latents over channels, add noise at step t, predict it, MSE.  Returns (loss, noise_pred).  ``text_embeddings`` is
    truncated to the latent length like the reference's ``text_input[:, :audio_latents.size(1)]`` (:115)."""
    lat = pool_latents(audio_latents, max_length) if audio_latents.dim() == 4 else _need_cuda_f32("audio_latents", audio_latents)
    text_embeddings = _need_cuda_f32("text_embeddings", text_embeddings)[:, :lat.shape[1]].contiguous()
    if noise is None:
        noise = torch.randn_like(lat)
    noisy = model.q_sample(lat, t, noise)
    pred = model(noisy, text_embeddings, t)
    return mse_loss(pred, noise), pred
