"""Import the UNMODIFIED reference modules of the hot path (TEST / BASELINE INFRASTRUCTURE ONLY).

Looks in oracle/_ref/src (placed by oracle/build_ref.py; present on the GPU box) and falls back to /root/reference/src
(the build container).  ``model.NeuralAudioCodec`` -- which downloads GPT-2 / EnCodec from the HF hub in its constructor
(NeuralAudioCodec.py:15-18) -- is replaced by a stub module before ``model.DiTTO`` is imported, and ``torch.load`` of the NAC
checkpoint (DiTTO.py:24) is stubbed during construction: exactly the recipe of SURVEY.md appendix B.  Everything the hot
path executes (DiTTO.forward, GlobalAdaLN, DiT, RotaryEmbedding) is the reference's own code.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = [os.path.join(HERE, "_ref", "src"), "/root/reference/src"]


class _StubNAC(nn.Module):
    def __init__(self, **kw):
        super().__init__()
        self.language_model = nn.Identity()
        self.audio_encoder = nn.Identity()

    def load_state_dict(self, *a, **k):
        return None


def find_source():
    for c in CANDIDATES:
        if os.path.exists(os.path.join(c, "model", "DiTTO.py")) and os.path.exists(os.path.join(c, "components", "DiT.py")):
            return c
    return None


def available() -> bool:
    return find_source() is not None


def import_ditto():
    """-> the reference's ``model.DiTTO`` module (cached in sys.modules under its own name)."""
    src = find_source()
    if src is None:
        raise ImportError("reference sources not found: run `python oracle/build_ref.py` in the build container")
    if "model.DiTTO" in sys.modules and getattr(sys.modules["model.DiTTO"], "_ditto_ref_src", None) == src:
        return sys.modules["model.DiTTO"]
    if src not in sys.path:
        sys.path.insert(0, src)
    if "model.NeuralAudioCodec" not in sys.modules:      # never import the real one: it pulls pretrained weights
        stub = types.ModuleType("model.NeuralAudioCodec")
        stub.NAC = _StubNAC
        sys.modules["model.NeuralAudioCodec"] = stub
    import model.DiTTO as M   # noqa: E402  (namespace package `model` under `src`)
    M.NAC = _StubNAC
    M._ditto_ref_src = src
    return M


def build_reference(cfg, sd=None):
    """Reference DiTTO (eval, CPU fp32) with the shapes of an OracleConfig; ``sd`` = state_dict to load (oracle.make_state_dict)."""
    M = import_ditto()
    real_load = torch.load
    torch.load = lambda *a, **k: {"model_state_dict": {}}
    try:
        with contextlib.redirect_stdout(io.StringIO()):      # the constructor prints "[INFO] Loading NAC model..."
            ref = M.DiTTO(hidden_dim=cfg.hidden_dim, num_layers=cfg.num_layers, num_heads=cfg.num_heads, time_dim=cfg.time_dim,
                          text_dim=cfg.text_dim, diffusion_steps=cfg.diffusion_steps, nac_model_path="unused").eval()
    finally:
        torch.load = real_load
    if sd is not None:
        missing, unexpected = ref.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.startswith("nac.") for k in missing), (missing, unexpected)
    return ref
