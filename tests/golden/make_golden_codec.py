"""Mint the golden fixture of the hand-off steps either side of the loop from the UNMODIFIED reference
(run in the build container only; needs /root/reference).

    python tests/golden/make_golden_codec.py          # writes tests/golden/codec.npz

  * components.VectorQuantizer.VectorQuantizer (src/components/VectorQuantizer.py) is imported as it is, its ``codebook``
    parameter overwritten with oracle.make_codebook (seeded; regenerates on the GPU box), and run on
      - a tiny case stored in full (codebook, latents, indices, incl. an exact tie and a duplicated code),
      - the full-size hand-off of SpeechGenerator.py:117-118: 2 x 750 sampled-latent-like frames, 1024 x 768 codebook,
        ``unsqueeze(1).repeat(1, 2, 1, 1)``; indices stored as int16 plus the winner's margin over the runner-up.
  * the validation iteration of src/TrainDiTTO.py:113-127 (pool -> q_sample -> DiTTO.forward -> nn.MSELoss) run with the
    reference DiTTO module and torch's own ops on the tiny model of tiny_full.npz.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ditto_oracle as O  # noqa: E402
import make_golden as G  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    M, _ = G.import_reference()
    from components.VectorQuantizer import VectorQuantizer as RefVQ
    out = {}

    # ---- tiny VQ, stored in full; row 0 sits exactly between codes 3 and 7 (tie -> lowest index), code 9 duplicates code 5
    torch.manual_seed(0)
    vq = RefVQ(16, 8).eval()
    cb = O.make_codebook(16, 8, seed=21)
    cb[9] = cb[5]
    cb[3] *= 0.1                            # the two smallest-norm codes, equal norm:
    cb[7] = -cb[3]                          # z = 0 is exactly equidistant from both
    with torch.no_grad():
        vq.codebook.copy_(cb)
    g = torch.Generator().manual_seed(22)
    lat = torch.randn(2, 3, 11, 8, generator=g) * 0.05
    lat[0, 0, 0] = 0.0
    lat[1, 2, 5] = cb[9]                    # distance 0 to codes 5 and 9
    with torch.no_grad():
        idx = vq(lat)
    assert int(idx[0, 0, 0]) == 3 and int(idx[1, 2, 5]) == 5, (idx[0, 0, 0], idx[1, 2, 5])   # ties -> lowest index
    out.update(tiny_codebook=cb.numpy(), tiny_latents=lat.numpy(), tiny_indices=idx.numpy())

    # ---- full-size hand-off (SpeechGenerator.py:117-118)
    K, D, B, T = 1024, 768, 2, 750
    vq = RefVQ(K, D).eval()
    cb = O.make_codebook(K, D, seed=31)
    with torch.no_grad():
        vq.codebook.copy_(cb)
    g = torch.Generator().manual_seed(32)
    lat = torch.randn(B, T, D, generator=g) * 0.05   # codebook-scale latents: the winner is not decided by |c|^2 alone
    with torch.no_grad():
        idx = vq(lat.unsqueeze(1).repeat(1, 2, 1, 1))
        d = O.vq_distances(cb.double(), lat.reshape(-1, D).double())
    top2 = torch.topk(d, 2, dim=1, largest=False).values
    out.update(full_meta=np.array([K, D, B, T, 31, 32]), full_indices=idx.numpy().astype(np.int16),
               full_margin=(top2[:, 1] - top2[:, 0]).float().numpy(),
               full_codebook_sum=np.array([float(cb.double().sum())]), full_latents_sum=np.array([float(lat.double().sum())]))
    print("vq: distinct codes used", int(idx.unique().numel()), " min margin", float((top2[:, 1] - top2[:, 0]).min()))

    # ---- validation iteration on the tiny model (TrainDiTTO.py:113-127), reference module + torch ops
    tiny = np.load(os.path.join(HERE, "tiny_full.npz"))
    cfgv = [int(v) for v in tiny["cfg"]]
    cfg = O.OracleConfig(*cfgv)
    sd = {k[4:]: torch.from_numpy(tiny[k]) for k in tiny.files if k.startswith("sd::")}
    ref = G.build_reference(M, cfg, sd)
    g = torch.Generator().manual_seed(41)
    Bv, Cv, Tv, max_len = 2, 2, 30, 24
    audio_latents = torch.randn(Bv, Cv, Tv, cfg.hidden_dim, generator=g)
    text = torch.randn(Bv, 40, cfg.text_dim, generator=g)
    t = torch.tensor([6, 1], dtype=torch.long)
    with torch.no_grad():
        pooled = audio_latents[:, :, :max_len].mean(dim=1)                   # TrainDiTTO.py:113-114
        text_c = text[:, :pooled.size(1)]                                     # :115
        noise = torch.randn(pooled.shape, generator=g)
        noisy = ref.q_sample(pooled, t, noise)                                # :123
        pred = ref(noisy, text_c, t)                                          # :124
        loss = nn.MSELoss()(pred, noise)                                      # :126
    out.update(val_audio_latents=audio_latents.numpy(), val_text=text.numpy(), val_t=t.numpy(), val_noise=noise.numpy(),
               val_pooled=pooled.numpy(), val_pred=pred.numpy(), val_loss=np.array([float(loss)], dtype=np.float32),
               val_max_len=np.array([max_len]))
    # a 3-channel pooling case (the sum order over channels matters beyond 2 channels)
    a3 = torch.randn(2, 3, 10, 16, generator=g)
    out.update(pool3_in=a3.numpy(), pool3_out=a3[:, :, :7].mean(dim=1).numpy())
    np.savez_compressed(os.path.join(HERE, "codec.npz"), **out)
    print("codec.npz written; loss =", float(loss))


if __name__ == "__main__":
    main()
