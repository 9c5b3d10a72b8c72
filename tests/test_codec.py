"""Hand-off steps either side of the sampling loop: latents -> codebook indices
(VectorQuantizer.py:22-43 at SpeechGenerator.py:117-118), channel-mean pooling (TrainDiTTO.py:70-71), the validation
iteration q_sample -> forward -> MSELoss (TrainDiTTO.py:113-127).

CPU part: the oracle restatement against tests/golden/codec.npz (minted from the unmodified reference by
tests/golden/make_golden_codec.py).  GPU part (-m gpu): the CUDA path through the C-ABI against the same goldens and the oracle.

Index parity bar: EXACT, except that a row whose two candidates are an fp32 re-association tie (their fp64 distances agree
to 1e-5 relative) may resolve either way; the goldens' smallest margin is reported by the fixture."""
import ctypes as C

import numpy as np
import pytest
import torch

import ditto_tts_b200 as D
from ditto_tts_b200 import _lib, codec
from oracle import ditto_oracle as O


def T_(a):
    return torch.from_numpy(np.asarray(a))


def full_case(g):
    K, Dm, B, T, cseed, lseed = (int(v) for v in g["full_meta"])
    cb = O.make_codebook(K, Dm, seed=cseed)
    lat = torch.randn(B, T, Dm, generator=torch.Generator().manual_seed(lseed)) * 0.05
    assert abs(float(cb.double().sum()) - float(g["full_codebook_sum"][0])) < 1e-9, "seeded codebook does not regenerate"
    assert abs(float(lat.double().sum()) - float(g["full_latents_sum"][0])) < 1e-9, "seeded latents do not regenerate"
    return cb, lat, T_(g["full_indices"]).long()


def assert_indices(cb, lat_flat, got, want, what):
    got, want = got.reshape(-1).cpu(), want.reshape(-1)
    bad = (got != want).nonzero().flatten()
    if bad.numel():
        ok = O.vq_near_tie(cb, lat_flat[bad], got[bad], want[bad])
        assert bool(ok.all()), f"{what}: {int((~ok).sum())} wrong winners (of {bad.numel()} differing rows)"
        assert bad.numel() <= max(1, want.numel() // 1000), f"{what}: {bad.numel()} near-tie rows is implausibly many"


# ------------------------------------------------------------------------------------------ CPU: oracle vs reference goldens
def test_oracle_vq_tiny_exact(golden):
    g = golden("codec.npz")
    cb, lat, want = T_(g["tiny_codebook"]), T_(g["tiny_latents"]), T_(g["tiny_indices"])
    got = O.vq_indices(cb, lat)
    assert torch.equal(got, want)
    assert int(got[0, 0, 0]) == 3 and int(got[1, 2, 5]) == 5     # exact ties resolve to the lowest index


def test_oracle_vq_full_size_exact(golden):
    g = golden("codec.npz")
    cb, lat, want = full_case(g)
    got = O.latents_to_codes(cb, lat, channels=2)
    assert got.shape == want.shape == (2, 2, 750)
    assert torch.equal(got, want)
    assert torch.equal(got[:, 0], got[:, 1])
    assert float(g["full_margin"].min()) > 0


def test_oracle_validation_step(golden):
    g = golden("codec.npz")
    tiny = golden("tiny_full.npz")
    cfg = O.OracleConfig(*[int(v) for v in tiny["cfg"]])
    sd = {k[4:]: T_(tiny[k]) for k in tiny.files if k.startswith("sd::")}
    a = T_(g["val_audio_latents"])
    assert torch.equal(O.pool_latents(a, int(g["val_max_len"][0])), T_(g["val_pooled"]))
    assert torch.equal(O.pool_latents(T_(g["pool3_in"]), 7), T_(g["pool3_out"]))
    loss, pred = O.validation_step(sd, cfg, a, T_(g["val_text"]), T_(g["val_t"]), T_(g["val_noise"]), int(g["val_max_len"][0]))
    assert O.rel_l2(pred, T_(g["val_pred"])) <= 2e-6
    assert abs(float(loss) - float(g["val_loss"][0])) <= 2e-6 * float(g["val_loss"][0])


def test_codec_host_refuses_cpu_tensors():
    torch.manual_seed(0)
    vq = codec.VectorQuantizer(16, 8)
    assert list(vq.state_dict().keys()) == ["codebook"] and vq.codebook.shape == (16, 8)
    with pytest.raises(D.DittoError, match="CUDA tensor"):
        vq(torch.zeros(1, 2, 3, 8))
    with pytest.raises(D.DittoError, match="CUDA tensor"):
        codec.pool_latents(torch.zeros(1, 2, 3, 8), 2)
    with pytest.raises(D.DittoError, match="CUDA tensor"):
        codec.mse_loss(torch.zeros(4), torch.zeros(4))


def test_codec_init_matches_reference_distribution():
    """Same init calls in the same order as VectorQuantizer.py:19-20 => a seeded construction is reproducible and
    xavier-bounded."""
    torch.manual_seed(5)
    a = codec.VectorQuantizer(32, 16)
    torch.manual_seed(5)
    w = torch.nn.Parameter(torch.randn(32, 16))
    torch.nn.init.xavier_uniform_(w)
    assert torch.equal(a.codebook.detach(), w.detach())


# ------------------------------------------------------------------------------------------ GPU: CUDA path vs goldens / oracle
@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def make_vq(cb, dev):
    vq = codec.VectorQuantizer(cb.shape[0], cb.shape[1])
    with torch.no_grad():
        vq.codebook.copy_(cb)
    return vq.to(dev)


@pytest.mark.gpu
def test_gpu_vq_tiny_exact_with_ties(dev, golden):
    g = golden("codec.npz")
    cb, lat, want = T_(g["tiny_codebook"]), T_(g["tiny_latents"]), T_(g["tiny_indices"])
    vq = make_vq(cb, dev)
    got = vq(lat.to(dev))
    assert got.dtype == torch.int64 and got.shape == want.shape
    assert int(got[0, 0, 0]) == 3 and int(got[1, 2, 5]) == 5
    assert_indices(cb, lat.reshape(-1, cb.shape[1]), got, want, "tiny")


@pytest.mark.gpu
def test_gpu_vq_full_size_vs_reference_golden(dev, golden):
    g = golden("codec.npz")
    cb, lat, want = full_case(g)
    vq = make_vq(cb, dev)
    got = codec.latents_to_codes(vq, lat.to(dev), channels=2)
    assert got.shape == (2, 2, 750)
    assert torch.equal(got[:, 0], got[:, 1])
    assert_indices(cb, lat.reshape(-1, 768).repeat_interleave(1, 0), got[:, 0], want[:, 0], "full-size")
    # the general 4-D entry point on the materialised repeat gives the same indices
    got4 = vq(lat.to(dev).unsqueeze(1).repeat(1, 2, 1, 1))
    assert torch.equal(got4, got)


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,K,Dm", [(1, 1, 1, 4), (3, 129, 130, 20), (2, 257, 1000, 767), (1, 5, 2048, 64), (0, 7, 8, 8)])
def test_gpu_vq_ragged_shapes_vs_oracle(dev, B, T, K, Dm):
    cb = O.make_codebook(K, Dm, seed=K + Dm)
    lat = torch.randn(B, T, Dm, generator=torch.Generator().manual_seed(B * 1000 + T)) * 0.1
    vq = make_vq(cb, dev)
    got = vq.encode(lat.to(dev), repeat_channels=3)
    assert got.shape == (B, 3, T)
    if B == 0:
        return
    want = O.latents_to_codes(cb, lat, channels=3)
    assert_indices(cb, lat.reshape(-1, Dm).repeat_interleave(1, 0), got[:, 0], want[:, 0], f"{B}x{T}x{K}x{Dm}")
    assert torch.equal(got[:, 0], got[:, 2])


@pytest.mark.gpu
def test_gpu_vq_properties_at_full_c2_size(dev):
    """Size-independent properties at the C2 hand-off size (16 x 750 frames, 1024 x 768 codebook): a codebook row maps
    to itself, indices are invariant to the order of the rows, and the chosen code is within fp32 noise of the best."""
    K, Dm, B, T = 1024, 768, 16, 750
    cb = O.make_codebook(K, Dm, seed=3)
    vq = make_vq(cb, dev)
    ident = vq.encode(cb.to(dev).view(1, K, Dm))
    assert torch.equal(ident.view(-1).cpu(), torch.arange(K))
    lat = (torch.randn(B, T, Dm, generator=torch.Generator().manual_seed(9)) * 0.05).to(dev)
    got = vq.encode(lat)
    perm = torch.randperm(B * T, generator=torch.Generator().manual_seed(1)).to(dev)
    got_p = vq.encode(lat.view(1, B * T, Dm)[:, perm].contiguous())
    assert torch.equal(got.view(-1)[perm], got_p.view(-1))
    d = O.vq_distances(cb.to(dev).double(), lat.view(-1, Dm).double())
    chosen = d.gather(1, got.view(-1, 1)).squeeze(1)
    assert bool((chosen - d.min(dim=1).values <= 1e-5 * chosen.abs()).all())


@pytest.mark.gpu
def test_gpu_pool_and_mse_vs_golden(dev, golden):
    g = golden("codec.npz")
    a = T_(g["val_audio_latents"]).to(dev)
    got = codec.pool_latents(a, int(g["val_max_len"][0]))
    assert torch.equal(got.cpu(), T_(g["val_pooled"]))                       # 2 channels: (a + b) / 2 is exact
    got3 = codec.pool_latents(T_(g["pool3_in"]).to(dev), 7)
    assert O.rel_l2(got3.cpu(), T_(g["pool3_out"])) <= 1e-7
    assert codec.pool_latents(a, 10_000).shape == (2, 30, a.shape[-1])      # cap larger than the input: no truncation
    pred, noise = T_(g["val_pred"]).to(dev), T_(g["val_noise"]).to(dev)
    loss = codec.mse_loss(pred, noise)
    assert abs(float(loss) - float(g["val_loss"][0])) <= 1e-6 * float(g["val_loss"][0])
    big_a = torch.randn(3_000_001, generator=torch.Generator().manual_seed(2))
    big_b = torch.randn(3_000_001, generator=torch.Generator().manual_seed(3))
    want = float(((big_a.double() - big_b.double()) ** 2).mean())
    assert abs(float(codec.mse_loss(big_a.to(dev), big_b.to(dev))) - want) <= 1e-6 * want


@pytest.mark.gpu
@pytest.mark.parametrize("precision,bar", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_gpu_validation_step_vs_reference_golden(dev, golden, precision, bar):
    g = golden("codec.npz")
    tiny = golden("tiny_full.npz")
    cfg = O.OracleConfig(*[int(v) for v in tiny["cfg"]])
    sd = {k[4:]: T_(tiny[k]) for k in tiny.files if k.startswith("sd::")}
    m = D.DiTTO(hidden_dim=cfg.hidden_dim, num_layers=cfg.num_layers, num_heads=cfg.num_heads, time_dim=cfg.time_dim,
                text_dim=cfg.text_dim, diffusion_steps=cfg.diffusion_steps, precision=precision)
    m.load_state_dict(sd)
    m = m.to(dev)
    loss, pred = codec.validation_step(m, T_(g["val_audio_latents"]).to(dev), T_(g["val_text"]).to(dev), T_(g["val_t"]).to(dev),
                                       T_(g["val_noise"]).to(dev), int(g["val_max_len"][0]))
    assert O.rel_l2(pred.cpu(), T_(g["val_pred"])) <= bar
    assert abs(float(loss) - float(g["val_loss"][0])) <= 2 * bar * float(g["val_loss"][0])


@pytest.mark.gpu
def test_gpu_codec_bad_arguments(dev):
    lib = _lib.load()
    z = torch.zeros(8, device=dev)
    assert lib.ditto_vq_encode(None, 1, 1, 8, C.c_void_p(z.data_ptr()), 1, C.c_void_p(z.data_ptr()), 1, C.c_void_p(z.data_ptr()), None) == -1
    assert lib.ditto_pool_latents(C.c_void_p(z.data_ptr()), 1, 1, 1, 6, 1, C.c_void_p(z.data_ptr()), None) != 0   # dim % 4
    assert lib.ditto_mse_loss(C.c_void_p(z.data_ptr()), C.c_void_p(z.data_ptr()), 8, C.c_void_p(z.data_ptr()), C.c_void_p(z.data_ptr()), 8, None) != 0
