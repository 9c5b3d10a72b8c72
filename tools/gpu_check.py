"""Developer diagnostics run on the GPU box: each group prints rel-L2 errors of the CUDA path vs the CPU oracle.
    python tools/gpu_check.py <group> ...      groups: ops fwd32 gemmtc gemmkn fwd16 fwd16nofuse sampler
Run each group in its own process (tools/gpu_session.sh does) so that a trapped kernel cannot poison the rest."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ditto_oracle as O  # noqa: E402  (checker only)
import ditto_tts_b200 as D  # noqa: E402
from ditto_tts_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
ST = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731


def rel(a, b):
    return O.rel_l2(a.float().cpu(), b.float().cpu())


def report(name, err, bar):
    print(f"  {'PASS' if err <= bar else 'FAIL'}  {name:<58s} rel_l2={err:.3e}  (bar {bar:.0e})", flush=True)
    return err <= bar


def build_model(cfg, sd, precision, fused_rope=True):
    m = D.DiTTO(hidden_dim=cfg.hidden_dim, num_layers=cfg.num_layers, num_heads=cfg.num_heads, time_dim=cfg.time_dim,
                text_dim=cfg.text_dim, diffusion_steps=cfg.diffusion_steps, precision=precision, fused_rope=fused_rope,
                fold_cross=fused_rope)
    m.load_state_dict(sd, strict=True)
    return m.to(dev)


def g_ops():
    lib = _lib.load()
    ok = True
    g = torch.Generator().manual_seed(0)
    # layernorm
    for rows, H in ((37, 768), (5, 64), (1000, 1024)):
        x = torch.randn(rows, H, generator=g) * 3 + 1
        ga, be = torch.randn(H, generator=g), torch.randn(H, generator=g)
        ref = torch.nn.functional.layer_norm(x, (H,), ga, be, 1e-5)
        xd, gd, bd = x.to(dev), ga.to(dev), be.to(dev)
        y = torch.empty_like(xd)
        _lib.check(lib.ditto_layernorm(P(xd), P(gd), P(bd), P(y), 0, rows, H, ST()))
        ok &= report(f"layernorm f32 {rows}x{H}", rel(y, ref), 1e-6)
        yb = torch.empty(rows, H, dtype=torch.bfloat16, device=dev)
        _lib.check(lib.ditto_layernorm(P(xd), P(gd), P(bd), P(yb), 1, rows, H, ST()))
        ok &= report(f"layernorm bf16 {rows}x{H}", rel(yb, ref), 4e-3)
        _lib.check(lib.ditto_layernorm(P(xd), None, None, P(y), 0, rows, H, ST()))
        ok &= report(f"layernorm noaffine {rows}x{H}", rel(y, torch.nn.functional.layer_norm(x, (H,))), 1e-6)
    # sgemm
    for (M, N, K, batch, nk) in ((130, 70, 33, 1, 1), (257, 129, 768, 2, 1), (64, 300, 50, 3, 0), (750, 768, 750, 2, 0)):
        A = torch.randn(batch, M, K, generator=g)
        B = torch.randn(batch, N, K, generator=g) if nk else torch.randn(batch, K, N, generator=g)
        bias = torch.randn(N, generator=g)
        R = torch.randn(batch, M, N, generator=g)
        ref = 0.5 * (A @ (B.transpose(1, 2) if nk else B)) + bias + R
        Ad, Bd, bd, Rd = A.to(dev), B.to(dev), bias.to(dev), R.to(dev)
        Cd = torch.empty(batch, M, N, device=dev)
        _lib.check(lib.ditto_gemm_f32(P(Ad), K, M * K, P(Bd), K if nk else N, N * K, nk, P(Cd), N, M * N, P(bd), P(Rd), 0.5,
                                      M, N, K, batch, ST()))
        ok &= report(f"gemm_f32 M{M} N{N} K{K} b{batch} nk{nk}", rel(Cd, ref), 2e-6)
    # cfg + ddpm update, q_sample
    cfg = O.OracleConfig(64, 1, 2, 32, 64, 50)
    sd = O.make_state_dict(cfg, 0)
    m = build_model(cfg, sd, "fp32")
    s = D.DiTTOSampler(m)
    B, T, H = 3, 10, 64
    ec, eu, x, z = (torch.randn(B, T, H, generator=g) for _ in range(4))
    t = torch.tensor([49, 0, 17])
    betas, alphas, acp = O.sampler_tables(50)
    for guided in (True, False):
        w = 3.0
        eps = eu + w * (ec - eu) if guided else ec
        ref = O.p_sample_update(x, eps, z, t, betas, alphas, acp)
        out = torch.empty(B, T, H, device=dev)
        ecd, eud, xd, zd, td = ec.to(dev), eu.to(dev), x.to(dev), z.to(dev), t.to(dev)
        _lib.check(lib.ditto_cfg_ddpm_update(m.engine(), P(ecd), P(eud) if guided else None, P(xd), P(zd),
                                             P(td), w, P(out), B, T * H, ST()))
        ok &= report(f"cfg_ddpm_update guided={guided}", rel(out, ref), 1e-6)
    qs = m.q_sample(x.to(dev), t.to(dev), z.to(dev))
    ok &= report("q_sample", rel(qs, O.q_sample(sd, x, t, z)), 1e-6)
    return ok


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name))


def fwd_cases(precision, bar, fused=True):
    ok = True
    g = golden("tiny_full.npz")
    cfg = O.OracleConfig(*[int(v) for v in g["cfg"]])
    sd = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}
    m = build_model(cfg, sd, precision, fused)
    out = m(torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["text"]).to(dev), torch.from_numpy(g["t"]).to(dev))
    ok &= report(f"{precision} tiny forward vs REFERENCE golden", rel(out, torch.from_numpy(g["out"])), bar)
    del m
    f = golden("full_size.npz")
    for name in ("ragged", "c1_default", "ctor_default"):
        meta = [int(v) for v in f[f"{name}::meta"]]
        cfg = O.OracleConfig(*meta[:6])
        wseed, iseed, B, T, S, stride = meta[6:]
        sd = O.make_state_dict(cfg, wseed)
        x, text, _ = O.make_inputs(B, T, S, cfg, iseed)
        t = torch.from_numpy(f[f"{name}::t"])
        m = build_model(cfg, sd, precision, fused)
        t0 = time.time()
        out = m(x.to(dev), text.to(dev), t.to(dev))
        torch.cuda.synchronize()
        ok &= report(f"{precision} {name} forward vs REFERENCE golden ({time.time() - t0:.2f}s)",
                     rel(out[:, ::stride], torch.from_numpy(f[f"{name}::out_sub"])), bar)
        del m
        torch.cuda.empty_cache()
    return ok


def g_gemmtc(kn=False):
    lib = _lib.load()
    ok = True
    g = torch.Generator().manual_seed(0)
    shapes = ((128, 256, 64), (128, 256, 768), (300, 200, 136), (1000, 768, 768), (24000, 2304, 768), (4096, 768, 3072))
    for (M, N, K) in shapes:
        A = torch.randn(M, K, generator=g).bfloat16()
        W = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
        bias = torch.randn(N, generator=g)
        Ad, Wd, bd = A.to(dev), W.to(dev), bias.to(dev)
        ref = (Ad.float() @ Wd.float().T + bd)
        for out_bf16 in (0, 1):
            ldc = (N + 7) // 8 * 8
            Cd = torch.zeros(M, ldc, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=dev)
            t0 = time.time()
            _lib.check(lib.ditto_gemm_bf16(P(Ad), K, P(Wd), K, P(Cd), ldc, out_bf16, P(bd), None, 0, 1.0, M, N, K, ST()))
            torch.cuda.synchronize()
            ok &= report(f"gemm_bf16 M{M} N{N} K{K} out_bf16={out_bf16} ({(time.time() - t0) * 1e3:.1f} ms)",
                         rel(Cd[:, :N], ref), 4e-3 if out_bf16 else 2e-5)
    # timing of the big ones
    for (M, N, K) in ((24000, 768, 768), (24000, 2304, 768), (24000, 768, 3072)):
        Ad = torch.randn(M, K, device=dev).bfloat16()
        Wd = torch.randn(N, K, device=dev).bfloat16()
        Cd = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        for _ in range(3):
            lib.ditto_gemm_bf16(P(Ad), K, P(Wd), K, P(Cd), N, 1, None, None, 0, 1.0, M, N, K, ST())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lib.ditto_gemm_bf16(P(Ad), K, P(Wd), K, P(Cd), N, 1, None, None, 0, 1.0, M, N, K, ST())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"  TIME gemm_bf16 M{M} N{N} K{K}: {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    return ok


def g_sampler(precision="fp32", bar=1e-4):
    ok = True
    g = golden("tiny_full.npz")
    cfg = O.OracleConfig(*[int(v) for v in g["cfg"]])
    sd = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}
    m = build_model(cfg, sd, precision)
    s = D.DiTTOSampler(m)
    out = s.sample_latents(torch.from_numpy(g["text"]).to(dev), x_init=torch.from_numpy(g["x"]).to(dev),
                           noise=torch.from_numpy(g["ref_noise"]).to(dev))
    ok &= report(f"{precision} tiny 8-step sampler vs REFERENCE __p_sample loop", rel(out, torch.from_numpy(g["sampled"])), bar * 3)
    del m, s
    c = golden("cfg_traj.npz")
    B, T, S, iseed, wseed, steps = [int(v) for v in c["meta"]]
    cfg = O.OracleConfig(768, 5, 1, 256, 768, steps)
    sd = O.make_state_dict(cfg, wseed)
    x, text, noise = O.make_inputs(B, T, S, cfg, iseed, steps_noise=steps)
    m = build_model(cfg, sd, precision)
    s = D.DiTTOSampler(m, guidance_scale=float(c["w"][0]))
    rec = []
    t0 = time.time()
    out = s.sample_latents(text.to(dev), x_init=x.to(dev), noise=noise.to(dev), record=rec)
    torch.cuda.synchronize()
    print(f"  50-step CFG B={B} T={T}: {time.time() - t0:.2f}s")
    ok &= report(f"{precision} CFG step 0 eps vs REFERENCE", rel(rec[0], torch.from_numpy(c["eps_first"])), bar)
    ok &= report(f"{precision} CFG last eps vs REFERENCE trajectory", rel(rec[-1], torch.from_numpy(c["eps_last"])), bar * 3)
    ok &= report(f"{precision} CFG final latent vs REFERENCE trajectory", rel(out, torch.from_numpy(c["final"])), bar * 3)
    return ok


GROUPS = {
    "ops": g_ops,
    "fwd32": lambda: fwd_cases("fp32", 1e-4),
    "gemmtc": g_gemmtc,
    "fwd16": lambda: fwd_cases("bf16", 2e-2, True),
    "fwd16nofuse": lambda: fwd_cases("bf16", 2e-2, False),
    "sampler32": lambda: g_sampler("fp32", 1e-4),
    "sampler16": lambda: g_sampler("bf16", 2e-2),
}

if __name__ == "__main__":
    allok = True
    for name in sys.argv[1:]:
        print(f"== {name}", flush=True)
        try:
            allok &= bool(GROUPS[name]())
        except Exception as e:  # noqa: BLE001
            allok = False
            print(f"  ERROR in {name}: {type(e).__name__}: {e}", flush=True)
    sys.exit(0 if allok else 1)
