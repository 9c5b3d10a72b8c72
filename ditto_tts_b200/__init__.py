"""ditto_tts_b200 -- B200-native (sm_100a) DiT denoiser + DDPM/CFG sampler of DiTTo-TTS.

Host side: Python mirrors of the reference's module and sampler signatures (model.py, sampler.py).
Device side: libditto_b200.so (csrc/, C-ABI in include/ditto_b200.h), hand-written CUDA only.
"""
from ._lib import DittoError, LIB_PATH, launch_count  # noqa: F401
from .config import ConfigDiTTO  # noqa: F401
from .model import DiT, DiTTO, GlobalAdaLN, RotaryEmbedding  # noqa: F401
from .sampler import DiTTOSampler  # noqa: F401
from .codec import VectorQuantizer, latents_to_codes, pool_latents, mse_loss, validation_step  # noqa: F401

__all__ = ["DiTTO", "DiT", "GlobalAdaLN", "RotaryEmbedding", "DiTTOSampler", "ConfigDiTTO", "DittoError",
           "launch_count", "LIB_PATH", "VectorQuantizer", "latents_to_codes", "pool_latents", "mse_loss", "validation_step"]
