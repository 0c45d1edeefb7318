"""Developer diagnostic (run on the GPU box): per-stage comparison of the CUDA path with the
fp64 oracle.  Not part of the product; tests/ holds the asserted versions."""
import math
import sys
import time

import torch

sys.path.insert(0, ".")
from deephall_b200 import _native as nat  # noqa: E402
from oracle import jets as OJ  # noqa: E402
from oracle import mcmc as OM  # noqa: E402
from oracle import psiformer as OP  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def check_gemm():
    for (M, N, K, rpg) in [(300, 256, 256, 1), (1000, 408, 256, 4), (77, 9, 256, 3), (129, 768, 256, 1)]:
        A = torch.randn(M, K, device=dev)
        W = torch.randn(K, N, device=dev)
        b = torch.randn(N, device=dev)
        out = nat.gemm(A, W, b, rpg)
        ref = A.double() @ W.double()
        ref[::rpg] += b.double()
        print(f"gemm simt M{M} N{N} K{K} rpg{rpg}: rel {rel(out, ref):.2e}")


def check_slogdet():
    for n, K in [(3, 1), (6, 1), (12, 4), (16, 16), (32, 2)]:
        m = torch.randn(64, K, n, n, dtype=torch.complex64, device=dev)
        sign, logabs, lpsi = nat.slogdet(m)
        s_ref, l_ref = torch.linalg.slogdet(m.cpu().to(torch.complex128))
        mx = l_ref.max(-1, keepdim=True).values
        lp_ref = torch.log((s_ref * torch.exp(l_ref - mx)).sum(-1)) + mx[..., 0]
        dph = (lpsi.imag.double().cpu() - lp_ref.imag + math.pi) % (2 * math.pi) - math.pi
        print(f"slogdet n{n} K{K}: logabs {(logabs.double().cpu() - l_ref).abs().max():.2e} sign {(sign.cpu().to(torch.complex128) - s_ref).abs().max():.2e} "
              f"lpsi re {(lpsi.real.double().cpu() - lp_ref.real).abs().max():.2e} im {dph.abs().max():.2e}")


def make(cfgkw, B, seed=0, burn=True):
    cfg = OP.NetCfg(**cfgkw)
    p64 = OP.init_params(cfg, seed, torch.float64, 0.1)
    g = torch.Generator().manual_seed(1234 + seed)
    x = OM.init_guess(g, B, cfg.nelec, torch.float32)
    return cfg, p64, x


def check_forward(cfgkw, B=16, detail=False):
    cfg, p64, x = make(cfgkw, B)
    plan = nat.Plan(nspins=cfg.nspins, flux=cfg.flux, ndets=cfg.ndets, num_heads=cfg.num_heads,
                    heads_dim=cfg.heads_dim, num_layers=cfg.num_layers)
    lay = plan.param_layout()
    shapes = OP.param_shapes(cfg)
    assert list(lay.keys()) == list(shapes.keys()), (list(lay.keys()), list(shapes.keys()))
    assert plan.num_params == OP.num_params(cfg)
    flat = OP.flatten_params(p64).float().to(dev)
    xd = x.to(dev)
    lp = plan.logpsi(flat, xd)
    ref = OP.logpsi(p64, x.double(), cfg)
    dim = (lp.imag.double().cpu() - ref.imag + math.pi) % (2 * math.pi) - math.pi
    print(f"{cfgkw}: logpsi re err {(lp.real.double().cpu() - ref.real).abs().max():.2e} im err {dim.abs().max():.2e}")
    t = time.time()
    out = plan.local_energy(flat, xd)
    torch.cuda.synchronize()
    t_gpu = time.time() - t
    rj = OJ.jet_logpsi(p64, x.double(), cfg)
    refo = OJ.observables_from_jets(rj, x.double(), cfg)
    for k in ("energy", "kinetic", "potential", "angular_momentum_z", "angular_momentum_z_square", "angular_momentum_square"):
        a, b = out[k].cpu(), refo[k]
        if a.is_complex():
            a = a.to(torch.complex128)
        else:
            a = a.double()
        err = (a - b).abs() / b.abs().clamp(min=1.0)
        print(f"   {k:28s} rel err median {err.median():.2e} max {err.max():.2e}   (|ref| ~ {b.abs().mean():.3f})")
    if detail:
        rw = OJ.Rows(cfg.nelec)
        lpj = plan.debug_buffer(nat.OP_LOCAL_ENERGY, B, "lpjet").view(B, rw.R, 2)
        lpj = torch.view_as_complex(lpj.contiguous()).cpu().to(torch.complex128).T  # R,B
        d = (lpj - rj).abs()
        print("   lpjet err by row-class: J", d[rw.J].max().item(), "S", d[rw.S].max().item(), "D", d[rw.D].max().item(), "T", d[rw.T].max().item(),
              " |ref| S", rj[rw.S].abs().max().item(), "T", rj[rw.T].abs().max().item())
        hj, _ = OJ.jet_psiformer_layers(p64, x.double(), cfg, rw)  # R,B,N,D
        hg = plan.debug_buffer(nat.OP_LOCAL_ENERGY, B, "h").view(B, cfg.nelec, rw.R, cfg.dim).permute(2, 0, 1, 3).cpu().double()
        d = (hg - hj).abs()
        print("   h err by row-class: val", d[0].max().item(), "J", d[rw.J].max().item(), "S", d[rw.S].max().item(), "D", d[rw.D].max().item(), "T", d[rw.T].max().item(),
              " |ref| S", hj[rw.S].abs().max().item())
    print(f"   (local_energy B={B} took {t_gpu * 1e3:.1f} ms incl. first-call overheads)")
    return plan, flat, xd, cfg, p64


def check_mcmc(cfgkw, B=64):
    cfg, p64, x = make(cfgkw, B)
    plan = nat.Plan(nspins=cfg.nspins, flux=cfg.flux, ndets=cfg.ndets, num_heads=cfg.num_heads,
                    heads_dim=cfg.heads_dim, num_layers=cfg.num_layers)
    flat = OP.flatten_params(p64).float().to(dev)
    N = cfg.nelec
    g = torch.Generator().manual_seed(7)
    steps = 3
    rnd = OM.draw_randoms(g, steps, B, N)
    packed = torch.stack([torch.cat([n_, u_, a_[:, None]], dim=1) for (n_, u_, a_) in rnd]).contiguous().to(dev)
    # proposal parity
    x2 = plan.mcmc_propose(x.to(dev), 0.1, randoms=packed[0])
    x2_ref = OM.sph_sampling(x, rnd[0][0], rnd[0][1], 0.1)
    print(f"mcmc propose err {(x2.cpu() - x2_ref).abs().max():.2e}")
    # accept parity on identical inputs
    lp1 = torch.randn(B)
    lp2 = lp1 + torch.randn(B)
    x1d, lp1d = x.to(dev).clone(), lp1.to(dev).clone()
    nacc = plan.mcmc_accept(x1d, x2, lp1d, lp2.to(dev), randoms=packed[0])
    cond = OM.mh_accept(lp1, lp2, rnd[0][2])
    print(f"mcmc accept: naccept {nacc.item()} ref {int(cond.sum())} lp-equal {torch.equal(lp1d.cpu(), torch.where(cond, lp2, lp1))}")
    # full sweep vs fp32 oracle chain with identical randoms
    xs = x.to(dev).clone()
    nacc, lp = plan.mcmc_sweep(flat, xs, steps, 0.1, randoms=packed, want_lp=True)
    p32 = OP.cast_params(p64, torch.float32)
    xr, pm = OM.mcmc_step(lambda xx: OP.logpsi(p32, xx, cfg), x, rnd, 0.1)
    print(f"mcmc sweep: naccept {nacc.item()} ref {round(pm * steps * B)}; walkers differing {(xs.cpu() - xr).abs().amax((1, 2)).gt(1e-4).sum().item()} / {B}")
    # philox mode runs and moves walkers
    xs2 = x.to(dev).clone()
    nacc2, _ = plan.mcmc_sweep(flat, xs2, 10, 0.1, seed=42)
    print(f"mcmc philox: pmove {nacc2.item() / (10 * B):.3f}")


def check_vjp(cfgkw, B=16):
    cfg, p64, x = make(cfgkw, B)
    plan = nat.Plan(nspins=cfg.nspins, flux=cfg.flux, ndets=cfg.ndets, num_heads=cfg.num_heads,
                    heads_dim=cfg.heads_dim, num_layers=cfg.num_layers)
    flat64 = OP.flatten_params(p64)
    flat = flat64.float().to(dev)
    cot = torch.randn(B, 2)
    g = plan.logpsi_vjp(flat, x.to(dev), cot.to(dev))
    p = flat64.clone().requires_grad_(True)
    lp = OP.logpsi(OP.unflatten_params(p, cfg), x.double(), cfg)
    obj = (lp.real * cot[:, 0].double() + lp.imag * cot[:, 1].double()).sum()
    (gref,) = torch.autograd.grad(obj, p)
    gd = g.double().cpu()
    print(f"{cfgkw}: vjp rel err (global) {(gd - gref).norm() / gref.norm():.2e}")
    off = 0
    for name, shape in OP.param_shapes(cfg).items():
        n = 1
        for q in shape:
            n *= q
        a, b = gd[off:off + n], gref[off:off + n]
        e = (a - b).norm() / (b.norm() + 1e-30)
        if e > 1e-4:
            print(f"     {name}: rel {e:.2e} |ref| {b.norm():.3e}")
        off += n


def timing(cfgkw, B):
    cfg, p64, x = make(cfgkw, 64)
    plan = nat.Plan(nspins=cfg.nspins, flux=cfg.flux, ndets=cfg.ndets, num_heads=cfg.num_heads,
                    heads_dim=cfg.heads_dim, num_layers=cfg.num_layers)
    flat = OP.flatten_params(p64).float().to(dev)
    xd = plan.init_walkers(B, seed=1)
    plan.mcmc_sweep(flat, xd, 10, 0.1, seed=3)
    def t(fn, n=3):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    ms = t(lambda: plan.logpsi(flat, xd))
    print(f"{cfgkw} B={B}: logpsi {ms:.2f} ms -> {B / ms * 1e3:.3e} evals/s")
    plan.profile_begin()
    plan.logpsi(flat, xd)
    pr = plan.profile_end()
    print("   logpsi profile:", {k: (round(v['ms'], 3), v['count'], round(v['flops'] / max(v['ms'], 1e-9) / 1e9, 1)) for k, v in pr.items()})
    ms = t(lambda: plan.mcmc_sweep(flat, xd, 10, 0.1, seed=5))
    print(f"   mcmc 10 moves {ms:.2f} ms -> {B * 10 / ms * 1e3:.3e} walker-steps/s")
    ms = t(lambda: plan.local_energy(flat, xd), n=2)
    print(f"   local_energy {ms:.2f} ms -> {B / ms * 1e3:.3e} evals/s")
    plan.profile_begin()
    plan.local_energy(flat, xd)
    pr = plan.profile_end()
    print("   LE profile:", {k: (round(v['ms'], 2), v['count'], round(v['flops'] / max(v['ms'], 1e-9) / 1e9, 1)) for k, v in pr.items()})
    cot = torch.randn(B, 2, device=dev) / B
    ms = t(lambda: plan.logpsi_vjp(flat, xd, cot), n=2)
    print(f"   vjp {ms:.2f} ms")
    plan.profile_begin()
    plan.logpsi_vjp(flat, xd, cot)
    pr = plan.profile_end()
    print("   VJP profile:", {k: (round(v['ms'], 2), v['count'], round(v['flops'] / max(v['ms'], 1e-9) / 1e9, 1)) for k, v in pr.items()})


if __name__ == "__main__":
    print(nat.load().dh_version().decode())
    check_gemm()
    check_slogdet()
    check_forward(dict(nspins=(3, 0), flux=2), detail=True)
    check_forward(dict(nspins=(5, 0), flux=11, ndets=2), detail=True)
    check_forward(dict(nspins=(6, 0), flux=15))
    check_forward(dict(nspins=(12, 0), flux=33), B=8, detail=True)
    check_forward(dict(nspins=(16, 0), flux=45, ndets=4), B=4)
    check_mcmc(dict(nspins=(6, 0), flux=15))
    check_vjp(dict(nspins=(3, 0), flux=2))
    check_vjp(dict(nspins=(5, 0), flux=11, ndets=2))
    check_vjp(dict(nspins=(12, 0), flux=33), B=8)
    timing(dict(nspins=(12, 0), flux=33), 8192)
