"""Benchmark of the walker-evaluation hot path (BASELINE.json metric: walker local-energy
evals/sec and VMC steps/sec at N=12).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host CPUs

A "step" is one energy evaluation of the walker batch as the reference's inference step runs it
(optimizers/none.py:32 -> loss.py:66-94, LossMode.ENERGY_DIFF): the `local_energy` pass (E_L + the
five observables, hamiltonian.py:207-210) over the rank's walkers followed by the pmean'd energy
statistics (one packed NCCL all-reduce when N > 1).  Default workload = BASELINE.json configs[2]
(the configuration the metric is quoted on): nspins=[12,0], flux=33, default Psiformer (4x64,
2 layers, 1 det), GLOBAL batch 8192 walkers, interaction_strength 1.

  --config c2|c3|c4|c5k4   the other BASELINE configurations (global batch 4096 / 8192 / 8192 / 16384)
  --scaling strong|weak    strong (default, what BASELINE.json asks: "batch 8192 at 1/2/4/8 B200"): the global
                           batch is sharded, B/N walkers per GPU; weak: the full batch on every GPU

`extra` carries the other BASELINE metrics at the same N and scaling mode -- walker-steps/s of `mcmc_step`
and VMC steps/s (mcmc_step + loss_and_grad + optimizer + the statistics and gradient all-reduces) with
Adam and with KFAC -- and, when N > 1, the weak-scaling local-energy figure.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

_COMMON = dict(num_heads=4, heads_dim=64, num_layers=2, interaction_strength=1.0, mcmc_steps=10, mcmc_width=0.1)
WORKLOADS = {  # BASELINE.json configs[1..4]
    "c2": dict(nspins=(6, 0), flux=15, determinants=1, global_batch=4096, **_COMMON,
               name="c2: nspins=[6,0] flux=15 (1/3 Laughlin), default Psiformer, global batch 4096"),
    "c3": dict(nspins=(12, 0), flux=33, determinants=1, global_batch=8192, **_COMMON,
               name="c3: nspins=[12,0] flux=33 (1/3 filling), default Psiformer, global batch 8192"),
    "c4": dict(nspins=(10, 0), flux=21, determinants=1, global_batch=8192, **_COMMON,
               name="c4: nspins=[10,0] flux=21 (2/5 filling), default Psiformer, global batch 8192"),
    "c5k4": dict(nspins=(16, 0), flux=45, determinants=4, global_batch=16384, **_COMMON,
                 name="c5: nspins=[16,0] flux=45 (1/3 filling), Psiformer with 4 determinants, global batch 16384"),
}
WORKLOAD = WORKLOADS["c3"]
WORKLOAD_NAME = WORKLOAD["name"]
CPU_SAMPLE = 64  # walkers per CPU step: the same in `cpu_baseline` and in `--impl reference`
METRIC = "local_energy_evals_per_sec"
UNIT = "walker local-energy evals/s"


def flops_local_energy_per_walker(N, L, K, D=256, nl=2):
    """Flops of the dense contractions one local-energy evaluation executes here (R = 2N+8 jet rows;
    MHA-out folded into the Dense that follows it: 5 D^2 per layer instead of 6; layer-0 q|k|v taken
    straight from the 4 input features): 2*R*N*D*((5 nl - 3)*D + 2*L*N*K).  DESIGN.md section 5."""
    R = 2 * N + 8
    return 2.0 * R * N * D * ((5 * nl - 3) * D + 2 * L * N * K)


def gemm_traffic_per_launch():
    """dram bytes (read + write) per launch of the dominant kernel from the committed ncu --set full
    summary (profiles/r2_gemm_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_gemm_traffic.json")) as f:
            return json.load(f)["avg_dram_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.th.join(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# --------------------------------------------------------------------------------- CPU reference leg
def cpu_reference_local_energy(W, nwalkers=CPU_SAMPLE, steps=3, warmup=1):
    """Times the reference's algorithm (complex gradient + full Hessian of log psi,
    hamiltonian.py:105-133) restated in torch on the host cores, fp32, on `nwalkers` walkers per step.
    ONE protocol for both CPU legs (`cpu_baseline` and `--impl reference`): same sample, >= 1 warm-up pass, >= 3 timed
    passes, throughput from the mean pass time."""
    from oracle import hamiltonian as OH
    from oracle import mcmc as OM
    from oracle import psiformer as OP

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = OP.NetCfg(nspins=W["nspins"], flux=W["flux"], ndets=W["determinants"],
                    num_heads=W["num_heads"], heads_dim=W["heads_dim"], num_layers=W["num_layers"])
    params = OP.init_params(cfg, 0, torch.float32)
    x = OM.init_guess(torch.Generator().manual_seed(42), nwalkers, cfg.nelec, torch.float32)

    def f(xx):
        return OP.logpsi(params, xx, cfg)

    def one():
        return OH.batch_local_energy(f, x, cfg.Q, interaction_strength=W["interaction_strength"], chunk=32)

    for _ in range(max(warmup, 1)):
        one()
    steps = max(steps, 1)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = one()
    dt = time.perf_counter() - t0
    assert torch.isfinite(out["energy"].real).any()
    return nwalkers * steps / dt, dt / steps, cores


def cpu_sample_note(steps, sps):
    return (f"{CPU_SAMPLE} walkers of the same workload per step, 1 warm-up + {steps} timed steps ({sps:.2f} s each); torch-CPU "
            "restatement of the reference algorithm (grad + full Hessian, oracle/hamiltonian.py), eager fp32, all host cores; "
            "the JAX reference is not installable here (XLA-CPU would be several times faster than eager torch)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = WORKLOADS[args.config]
    steps = max(args.steps, 3)
    val, spstep, cores = cpu_reference_local_energy(W, steps=steps, warmup=max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": spstep * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (uniform walkers on the sphere, random-init parameters)",
        "config": {"workload": W["name"], "sample": f"{CPU_SAMPLE} walkers per step (throughput is per walker)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": cpu_sample_note(steps, spstep)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------- CUDA arm
def run_ours(args):
    import dataclasses

    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from deephall_b200 import loss, mcmc, networks  # noqa: F401
    from deephall_b200.config import Config, MCMC, Network, Optim, PsiformerNetwork, System
    from deephall_b200.train import VMC

    W = WORKLOADS[args.config]
    strong = args.scaling == "strong"
    if strong and W["global_batch"] % world:
        raise SystemExit(f"global batch {W['global_batch']} does not shard over {world} GPUs")
    B = W["global_batch"] // world if strong else W["global_batch"]  # walkers per GPU
    N = sum(W["nspins"])
    system = System(flux=W["flux"], nspins=W["nspins"], interaction_strength=W["interaction_strength"])
    network = Network(psiformer=PsiformerNetwork(W["num_heads"], W["heads_dim"], W["num_layers"], W["determinants"]))

    def make_vmc(per_gpu, optimizer="adam"):
        cfg = Config(batch_size=per_gpu * world, seed=42, system=system, network=network,
                     mcmc=MCMC(steps=W["mcmc_steps"], width=W["mcmc_width"]), optim=Optim(optimizer=optimizer))
        return VMC(cfg)

    vmc = make_vmc(B)
    model, params = vmc.model, vmc.state.params
    vmc.burn_in(args.burn_in)  # synthetic walkers: short equilibration, not timed
    data = vmc.state.data
    plan = model.plan(system)
    energy_step = loss.make_loss_fn(model.apply, system, loss.LossMode.ENERGY_DIFF)  # what optimizers/none.py:32 runs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local_rank)
    # ---- headline: device-resident energy evaluations (local-energy pass + pmean'd statistics)
    result = {}

    def le_step():
        result["stats"], result["diff"] = energy_step(params, data)

    for _ in range(args.warmup):
        le_step()
    barrier()
    if rank == 0:
        sampler.start()
    l0 = plan.launch_count
    ms_total = timed(le_step, args.steps, 0)
    launches = plan.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = B * world * args.steps / (ms_total * 1e-3)

    # ---- e2e: host walkers -> public API -> host energy statistics + per-walker clipped differences, every step
    x_host = data.cpu().pin_memory()
    diff_host = torch.empty((B,), dtype=torch.complex64).pin_memory()
    stat_host = torch.empty((2,), dtype=torch.float32).pin_memory()
    x_dev = torch.empty_like(data)

    def e2e_step():
        x_dev.copy_(x_host, non_blocking=True)
        stats, diff = energy_step(params, x_dev)
        diff_host.copy_(diff, non_blocking=True)
        stat_host.copy_(torch.view_as_real(stats["energy"].reshape(1)).reshape(2).float(), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    ms_e2e = timed(e2e_step, args.steps, 1)
    e2e_value = B * world * args.steps / (ms_e2e * 1e-3)

    # ---- the other two metrics of BASELINE.json (same hygiene, fewer steps)
    k2 = max(1, min(args.steps, 3))
    ms_mcmc = timed(lambda: vmc.mcmc_step(params, data, mcmc.PhiloxKey(7), W["mcmc_width"]), k2, 1) / k2
    ms_vmc = timed(lambda: vmc.step(sync_stats=False), k2, 2) / k2
    # the same iteration with the reference's default optimizer (config.py:159): KFAC, whose curvature statistics
    # cost one more forward + reverse pass (dh_kfac_factors) and ~30 small matrix inverses per step
    vmc_k = make_vmc(B, "kfac")
    vmc_k.state = vmc_k.state._replace(data=data.clone())
    ms_vmc_kfac = timed(lambda: vmc_k.step(sync_stats=False), k2, 3) / k2  # (3 warm-up steps: lazy library initialisation)
    del vmc_k
    extra_weak = None
    if world > 1 and strong:  # the weak-scaling figure next to the strong one: the full batch on every GPU
        vw = make_vmc(W["global_batch"])
        vw.burn_in(3)
        dw = vw.state.data
        ms_w = timed(lambda: energy_step(vw.state.params, dw), k2, 2) / k2
        ms_vw = timed(lambda: vw.step(sync_stats=False), k2, 1) / k2
        extra_weak = {"walkers_per_gpu": W["global_batch"], "local_energy_evals_per_sec": W["global_batch"] * world / (ms_w * 1e-3),
                      "ms_per_step": ms_w, "vmc_step_ms": ms_vw}
        del vw, dw

    # ---- roofline of the dominant kernel (dense contractions), CUDA-event timed per launch
    barrier()
    plan.profile_begin()
    plan.local_energy(params, data)
    prof = plan.profile_end()
    peaks = measured_peaks()
    g = prof["gemm"]
    achieved = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
    peak = peaks["bf16_tflops_sustained"]
    total_ms = sum(v["ms"] for v in prof.values())
    roofline = {
        "bound": "tensor", "kernel": "dense contraction (dh::tc::gemm_*)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak, "traffic": gemm_traffic_per_launch(),
        "launches": g["count"], "avg_launch_ms": g["ms"] / max(g["count"], 1),
        "share_of_step": g["ms"] / total_ms,
        "peak_source": peaks["source"] + ", dense bf16 sustained. fp32 accuracy costs 3 fp16-rate MMAs (hi*hi, lo*hi, hi*lo) per "
                       "algorithmic MAC, so the attainable frac is 1/3; `achieved` counts algorithmic flops (2MNK per launch)",
        "tensor_pipe_frac_incl_split": 3.0 * achieved / peak,
        "per_category_ms": {k: round(v["ms"], 3) for k, v in prof.items()},
        "whole_pass": {"tflops": sum(v["flops"] for v in prof.values()) / (total_ms * 1e-3) / 1e12,
                       "frac": sum(v["flops"] for v in prof.values()) / (total_ms * 1e-3) / 1e12 / peak},
    }

    if rank == 0:
        cpu = None
        if world == 1:
            v, sps, cores = cpu_reference_local_energy(W, steps=3, warmup=1)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": cpu_sample_note(3, sps)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": f"synthetic (Philox uniform walkers + {args.burn_in} burn-in sweeps, random-init parameters)",
            "config": {"workload": W["name"], "global_batch": B * world, "walkers_per_gpu": B,
                       "parallelism": f"walkers sharded x{world}; energy statistics all-reduced (NCCL) inside the timed step",
                       "l2": "per-pass working set (GBs of jet activations) exceeds the 126 MB L2; no explicit flush"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4 * world,
                    "d2h_bytes_per_step": (diff_host.numel() * 8 + 8) * world, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "extra": {
                "walker_steps_per_sec": B * world * W["mcmc_steps"] / (ms_mcmc * 1e-3),
                "mcmc_step_ms": ms_mcmc,
                "vmc_steps_per_sec": 1.0 / (ms_vmc * 1e-3),
                "vmc_step_ms": ms_vmc,
                "vmc_optimizer": "adam (SURVEY 8d M3); statistics + gradient all-reduces inside the step",
                "vmc_kfac_steps_per_sec": 1.0 / (ms_vmc_kfac * 1e-3),
                "vmc_kfac_step_ms": ms_vmc_kfac,
                "weak_scaling": extra_weak,
                "mean_energy": float(result["stats"]["energy"].real),
                "algorithmic_flops_per_walker": flops_local_energy_per_walker(N, W["flux"] + 1, W["determinants"]),
            },
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--burn-in", type=int, default=20, help="untimed Metropolis sweeps before the measurement (profiling runs use 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torchrun (one process per GPU)")
    run_ours(args)


if __name__ == "__main__":
    main()
