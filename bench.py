#!/usr/bin/env python
"""bench.py — leapfrog walker-steps/s of the ensemble-Hamiltonian hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5] [--impl ours|reference]

One "step" = one sampler iteration (two half-moves: every walker of both groups is proposed,
integrated for L leapfrog steps and accepted/rejected).  Default workload = BASELINE.json's c5:
1000-dim ill-conditioned Gaussian (kappa 1e4, full covariance), 65,536 walkers, walk move, L = 10,
fp32 — the configuration the metric is quoted on; it fits one GPU because the walk move runs in
projected form (no n x n momentum in HBM).  N > 1 shards the walkers of both halves by rows over the
ranks (strong scaling: total work fixed); after every half-move the kernels store the updated rows into the peers' copies over NVLink
(csrc/hx_dist.cuh: peer-memory stores + epoch flags, no collective on the data path).

Prints ONE JSON line (contract in the task statement) with `roofline`, `cpu_baseline`, `e2e`,
`clocks`, `gpu_launches`.  `--impl reference` times the reference algorithm on the host CPU: JAX is
not installable in this image, so that arm runs the NumPy oracle port (dense n x n momentum form,
exactly what the reference computes) on a bounded sample of the same workload.
"""
from __future__ import annotations

import os
import sys

if "reference" in sys.argv[1:]:
    # the reference arm is a host-BLAS workload: `torch.distributed.run` exports OMP_NUM_THREADS=1 for nproc > 1, which would
    # silently run it single-threaded (3x slower, measured in round 1).  Must happen before NumPy loads its BLAS.
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import argparse
import json
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "h-emcee_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # name: target, d, kappa, W, L, eps, x64, move          (BASELINE.json configs / SURVEY §8d)
    "c1": dict(target="gauss", d=10, kappa=1e3, W=32, L=10, eps=0.1, x64=True, move="walk"),
    "c2": dict(target="gauss", d=100, kappa=1e4, W=1024, L=10, eps=0.1, x64=False, move="walk"),
    "c2s": dict(target="gauss", d=100, kappa=1e4, W=1024, L=10, eps=0.1, x64=False, move="side"),
    "c3a": dict(target="rosenbrock", d=2, kappa=0, W=8192, L=10, eps=0.01, x64=False, move="walk"),
    "c3": dict(target="rosenbrock", d=50, kappa=0, W=8192, L=10, eps=0.01, x64=False, move="walk"),
    "c4": dict(target="funnel", d=25, kappa=0, W=16384, L=10, eps=0.05, x64=False, move="walk"),
    # c3 as BASELINE.json states it: "with affine-invariant step-size/trajectory-length tuning" -- run_mcmc with the step-size
    # heuristic, dual averaging and ChEES on (step_size = 0.01 as docs/tutorials/adaptation.ipynb cell 5); see bench_adaptive()
    "c3ad": dict(target="rosenbrock", d=50, kappa=0, W=8192, L=10, eps=0.01, x64=False, move="walk", adaptive=True),
    "c3aad": dict(target="rosenbrock", d=2, kappa=0, W=8192, L=10, eps=0.01, x64=False, move="walk", adaptive=True),
    # c5 with the same tuning hooks on (warm-up on the tensor-core path, adapter state on the device): `--warmup-iters 40 --steps 10`
    "c5ad": dict(target="gauss", d=1000, kappa=1e4, W=65536, L=10, eps=0.1, x64=False, move="walk", adaptive=True),
    # mid-size probes of the precision policy (not BASELINE configs)
    "m4": dict(target="gauss", d=128, kappa=1e3, W=2048, L=10, eps=0.1, x64=False, move="walk"),
    "m5": dict(target="gauss", d=64, kappa=1e3, W=8192, L=10, eps=0.1, x64=False, move="walk"),
    "m6": dict(target="gauss", d=128, kappa=1e3, W=8192, L=10, eps=0.1, x64=False, move="walk"),
    "m1": dict(target="gauss", d=256, kappa=1e3, W=4096, L=10, eps=0.1, x64=False, move="walk"),
    "m2": dict(target="gauss", d=256, kappa=1e3, W=16384, L=10, eps=0.1, x64=False, move="walk"),
    "m3": dict(target="gauss", d=512, kappa=1e3, W=16384, L=10, eps=0.1, x64=False, move="walk"),
    "c5": dict(target="gauss", d=1000, kappa=1e4, W=65536, L=10, eps=0.1, x64=False, move="walk"),
}
CPU_SAMPLE_W = {"c5ad": 4096, "c3ad": 1024, "c3aad": 1024, "m4": 1024, "m5": 1024, "m6": 1024, "m1": 1024, "m2": 1024, "m3": 1024, "c3a": 1024, "c1": 32, "c2": 1024, "c2s": 1024, "c3": 1024, "c4": 2048, "c5": 4096}


def make_cov(d, kappa, H=None):
    """Sigma = Q diag(0.1 * linspace(1, kappa, d)) Q^T with Q = qr(normal(PRNGKey(1), (d, d)))  (reference
    src/hemcee/tests/distribution.py:92-96).  H = the (d, d) float64 normal draw; by default it is drawn with
    the product's device RNG (bit-compatible with jax.random), the CPU reference arm passes the oracle's draw of the same key."""
    if H is None:
        import torch
        import hemcee_b200 as hx
        H = hx.random.normal(hx.random.PRNGKey(1), (d, d), torch.float64).cpu().numpy()      # 64-bit draws (the reference's tests run x64)
    Q, _ = np.linalg.qr(np.asarray(H, dtype=np.float64))
    eig = 0.1 * np.linspace(1.0, kappa, d)
    cov = (Q * eig) @ Q.T
    prec = (Q / eig) @ Q.T
    return 0.5 * (cov + cov.T), 0.5 * (prec + prec.T)


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's dense algorithm on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference(wl, steps, warmup, budget_s=None):
    from oracle import jaxrng as R, moves as OM, targets as OT
    cfg = WORKLOADS[wl]
    Ws = min(cfg["W"], CPU_SAMPLE_W[wl])
    d, L, eps = cfg["d"], cfg["L"], cfg["eps"]
    dt = np.float64 if cfg["x64"] else np.float32
    if cfg["target"] == "gauss":
        cov, _ = make_cov(d, cfg["kappa"], H=R.normal(R.PRNGKey(1), (d, d), np.float64))
        lp, grad = OT.make_gaussian(cov, np.zeros(d))
    elif cfg["target"] == "rosenbrock":
        lp, grad = OT.make_rosenbrock()
    else:
        lp, grad = OT.make_funnel()
    move = OM.hmc_walk_move if cfg["move"] == "walk" else OM.hmc_side_move
    k_init, k_run = R.split(R.PRNGKey(0), 2)
    x = R.normal(k_init, (Ws, d), dt)
    n = Ws // 2
    g1, g2 = x[:n].copy(), x[n:].copy()
    lp1, lp2 = lp(g1), lp(g2)
    keys = R.split(k_run, 4 * (steps + warmup)).reshape(steps + warmup, 4, 2)
    t0 = None
    done = 0
    for t in range(steps + warmup):
        if t == warmup:
            t0 = time.perf_counter()
        k = keys[t]
        q, la, _, lpq = move(g1, g2, eps, k[0], lp, grad, L, lp1)
        g1, _ = OM.accept_proposal(g1, q, la, k[2]); lp1, _ = OM.accept_proposal(lp1, lpq, la, k[2])
        q, la, _, lpq = move(g2, g1, eps, k[1], lp, grad, L, lp2)
        g2, _ = OM.accept_proposal(g2, q, la, k[3]); lp2, _ = OM.accept_proposal(lp2, lpq, la, k[3])
        if t >= warmup:
            done += 1
            if budget_s is not None and time.perf_counter() - t0 > budget_s:
                break
    el = time.perf_counter() - t0
    value = Ws * L * done / el
    sample = (f"{done} iterations of the {wl} workload at {Ws} of {cfg['W']} walkers (same dim {d}, L {L}, dtype "
              f"{'f64' if cfg['x64'] else 'f32'}); the reference's cost per walker-step grows ~linearly with the "
              f"walker count, so this sample flatters the CPU")
    return dict(value=value, unit="walker-steps/s", cores=os.cpu_count(), kind="port", sample=sample), el / max(done, 1)


def bench_adaptive(args, cfg, conf):
    """c3 with DA + ChEES tuning: one `run_mcmc` = step-size heuristic + `--warmup-iters` adaptive warm-up iterations (adapter
    state resident on the device: hx_run_warmup) + `--steps` main iterations with the tuned (step, L).  A step of the main phase
    costs W * L_found leapfrog walker-steps; the warm-up's L changes every half-move, so its throughput is reported in
    iterations/s with the trajectory lengths it ended on."""
    import torch
    import hemcee_b200 as hx
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    hx.config.update("enable_x64", cfg["x64"])
    W, d = cfg["W"], cfg["d"]
    k_init, k_run = hx.random.split(hx.random.PRNGKey(0), 2)
    if cfg["target"] == "gauss":            # c5ad: the tensor-core path with the adapter resident on the device (hx_run_warmup)
        tgt = hx.targets.GaussianFull(precision=make_cov(d, cfg["kappa"])[1])
        x0 = hx.random.normal(k_init, (W, d), torch.float32).cpu().pin_memory()
    else:
        tgt = hx.targets.Rosenbrock(d)
        x0 = (hx.random.normal(k_init, (W, d), torch.float32) * 0.3 + 1.0).cpu().pin_memory()
    Wm, K = args.warmup_iters, args.steps

    def one():
        s = hx.HamiltonianEnsembleSampler(W, d, tgt, step_size=cfg["eps"], move=hx.hmc_walk_move, verbose=False, store="thinned")
        t0 = time.perf_counter()
        # big ensembles: one iteration per batch (a c5 chain slice is 262 MB; the reference's default of >= 50 iterations per batch
        # would stage 13 GB on the device), exactly like the e2e leg of the main bench
        out = s.run_mcmc(k_run, x0, K, warmup=Wm, batch_size=1 if W * d > 10**7 else None)
        torch.cuda.synchronize(dev)
        return s, out, time.perf_counter() - t0
    one()                                   # untimed: workspaces, pinned buffers, graph capture
    import gc
    gc.collect()                            # the sampler <-> backend cycle keeps the pinned chain alive otherwise: the timed run would page-lock a new one
    clocks = ClockSampler(0, args.clock_period)
    if args.clock_period > 0:
        clocks.start()
    clocks.mark_begin()
    s, out, el = one()
    clocks.mark_end()
    clk = clocks.stop()
    step, Lf = float(s.found[0]), int(s.found[1])
    tm = s.timings
    main_ws = W * Lf * K / tm["main_s"]
    line = dict(metric="leapfrog_walker_steps_per_sec", value=main_ws, unit="walker-steps/s", n_gpus=1, steps=K, warmup=Wm,
                ms_per_step=tm["main_s"] * 1e3 / K, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32",
                data="synthetic", config=dict(conf, adaptation="find_reasonable_step_size + dual averaging + ChEES (defaults)",
                                              warmup_iterations=Wm, main_iterations=K),
                roofline=None, cpu_baseline=None, clocks=clk,
                e2e=dict(value=W * Lf * K / tm["main_s"], unit="walker-steps/s", h2d_bytes_per_step=int(W * d * 4 / max(K, 1)),
                         d2h_bytes_per_step=int(W * d * 4 + W * 4), seconds=el,
                         api="HamiltonianEnsembleSampler.run_mcmc(store='thinned') with adapt_* defaults (host initial_state -> host chain)"),
                gpu_launches=None,
                adaptive=dict(initial_step=float(s.step_size), found_step=step, found_L=Lf,
                              warmup_iterations_per_s=Wm / tm["warmup_s"], warmup_ms_per_iteration=tm["warmup_s"] * 1e3 / max(Wm, 1),
                              initial_step_size_s=tm["initial_step_size_s"], main_ms_per_iteration=tm["main_s"] * 1e3 / K,
                              acceptance_rate_main=float(np.mean(s.diagnostics_main["acceptance_rate"])),
                              note="value = main phase with the tuned (step, L); the adaptive warm-up changes L every half-move, so "
                                   "it is reported in iterations/s"))
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons during the timed region (B200_PROFILING.md's clocks line), read in-process through
    NVML every 10 ms (nvidia_ml_py).  An `nvidia-smi -lms` loop per rank was measured to stall kernel launches through the
    driver lock (2x slower steps at 8 ranks), so only rank 0 samples and it does so with direct NVML calls; the fallback is
    one `nvidia-smi -lms 100` process.  Only samples between mark_begin() and mark_end() are reported."""

    def __init__(self, gpu_index, period=0.01):
        self.gpu = gpu_index
        self.period = period
        self.rows = []          # (t, sm_mhz, power_w, reasons bitmask)
        self.max_mhz = None
        self.t0 = self.t1 = None
        self._stop = threading.Event()
        self._thread = None
        self._proc = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES if it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.gpu])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self._stop.is_set():
                    try:
                        self.rows.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                          pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                          int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))))
                    except Exception:
                        pass
                    time.sleep(self.period)
            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
            self.R = dict(hw_slowdown=pynvml.nvmlClocksEventReasonHwSlowdown, hw_thermal_slowdown=pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                          sw_thermal_slowdown=pynvml.nvmlClocksEventReasonSwThermalSlowdown, sw_power_cap=pynvml.nvmlClocksEventReasonSwPowerCap)
        except Exception:
            self._thread = None
            self._start_smi()

    def _start_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self._proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                           "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.R = dict(hw_slowdown=1, hw_thermal_slowdown=2, sw_thermal_slowdown=4, sw_power_cap=8)

            def read():
                for line in self._proc.stdout:
                    c = [x.strip() for x in line.split(",")]
                    try:
                        self.max_mhz = float(c[1])
                        mask = sum(b for b, v in zip((1, 2, 4, 8), c[3:7]) if v.lower().startswith("active"))
                        self.rows.append((time.perf_counter(), float(c[0]), float(c[2]), mask))
                    except Exception:
                        pass
            threading.Thread(target=read, daemon=True).start()
        except Exception:
            self._proc = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self._stop.set()
        if self._proc is not None:
            self._proc.terminate()
        if self._thread is None and self._proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["clock sampler unavailable"], samples=0)
        rows = [r for r in self.rows if self.t0 is None or self.t0 <= r[0] <= (self.t1 or r[0]) + 0.005]
        reasons = sorted(nm for nm, bit in self.R.items() if any(r[3] & bit for r in rows))
        return dict(sm_mhz=float(np.median([r[1] for r in rows])) if rows else None, sm_max_mhz=self.max_mhz, reasons=reasons,
                    samples=len(rows), power_w_max=max((r[2] for r in rows), default=None))


# DRAM traffic of the fused draw + projection of one c5 half-move (ncu --set full, profiles/r02_c5_fused_ncu_summary.txt)
FUSED_TRAFFIC_BYTES = 1.002e9
FUSED_TRAFFIC_SOURCE = ("profiles/r02_c5_fused_ncu_summary.txt: dram__bytes_read.sum 769.4 MB + dram__bytes_write.sum 232.6 MB of the one "
                        "k_tc_proj_fused launch of a half-move (algorithmic: C^T planes 134 MB + S0 131 MB; the two split-K slices add into "
                        "S0 with red.global.add, which reads it back once); staged path of round 1: 9.6 GB per half-move")


def algorithmic_flops_half_move(rows, n, d, L, target):
    """minimal (projected-form) flops of one half-move over `rows` walkers."""
    f = 2.0 * rows * n * d + (L + 1) * 2.0 * rows * d * d
    if target == "gauss":
        f += (L + 1) * 2.0 * rows * d * d + 2.0 * rows * d
    elif target == "rosenbrock":
        f += (L + 1) * 12.0 * rows * d
    else:
        f += (L + 1) * 8.0 * rows * d
    return f


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--precision", default="auto", choices=["auto", "exact", "bf16x3", "bf16", "tf32x3", "f16f8"])
    ap.add_argument("--leapfrog-form", default="auto", choices=["auto", "stepwise", "propagator"])
    ap.add_argument("--tune", action="append", default=[], help="hx_ctx_set_tuning knob as what=value (diagnostics)")
    ap.add_argument("--clock-period", type=float, default=0.01, help="NVML sampling period in s (0 = no clock sampling)")
    ap.add_argument("--ess-iters", type=int, default=192, help="iterations of the ESS/s leg (0 = skip)")
    ap.add_argument("--warmup-iters", type=int, default=500, help="adaptive workloads (c3ad): warm-up iterations with DA + ChEES")
    args = ap.parse_args()
    cfg = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    conf = dict(workload=f"{args.workload}: {cfg['d']}-dim {cfg['target']} (kappa {cfg['kappa']:g}), {cfg['W']} walkers, "
                         f"{cfg['move']} move, L={cfg['L']}, eps={cfg['eps']}",
                walkers=cfg["W"], dim=cfg["d"], L=cfg["L"], move=cfg["move"],
                parallelism=f"walker rows sharded over {max(world, 1)} GPU(s); updated rows stored into the peers' copies over NVLink by the kernels, one epoch barrier per half-move"
                if world > 1 else "single GPU",
                l2="state (>=262 MB at c5) and centred complement exceed the 126 MB L2; chain slice written every step")

    if args.impl == "reference":
        if rank != 0:
            return
        cb, per_iter = cpu_reference(args.workload, args.steps, max(args.warmup, 0), budget_s=150.0)   # bounded: a few minutes at most
        Ws = min(cfg["W"], CPU_SAMPLE_W[args.workload])
        if Ws != cfg["W"]:
            # the line describes what RAN: the bounded sample, with the full-size cost stated as an extrapolation
            conf = dict(conf, walkers=Ws, full_size_walkers=cfg["W"],
                        workload=conf["workload"] + f" -- CPU arm run on a {Ws}-walker sample",
                        extrapolated_full_size=dict(
                            walker_steps_per_sec=cb["value"] * Ws / cfg["W"],
                            how="dense reference form: 4 (L + 1) n^2 d flops per half-move, i.e. cost per walker-step grows linearly with "
                                "the walkers per group n; value scaled by sample walkers / full walkers (an upper bound for the CPU: the "
                                "n x n momentum of the full size, 4 GiB in f32, no longer fits any cache)"))
        line = dict(impl="reference", metric="leapfrog_walker_steps_per_sec", value=cb["value"], unit="walker-steps/s",
                    n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=per_iter * 1e3,
                    higher_is_better=True, scaling="strong", vs_baseline=None,
                    dtype="f64" if cfg["x64"] else "f32", data="synthetic", config=conf, cpu_baseline=cb,
                    e2e=dict(value=cb["value"], unit="walker-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    gpu_launches=0, note="reference = NumPy oracle port of hemcee (JAX not installable here)")
        print(json.dumps(line))
        return

    if cfg.get("adaptive"):
        if rank == 0:
            bench_adaptive(args, cfg, conf)
        return

    import torch
    import torch.distributed as dist
    import hemcee_b200 as hx
    from hemcee_b200 import _lib, _rt

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hx.config.update("enable_x64", cfg["x64"])
    hx.config.update("precision", args.precision)
    hx.config.update("leapfrog_form", args.leapfrog_form)
    tdt = torch.float64 if cfg["x64"] else torch.float32
    W, d, L, eps = cfg["W"], cfg["d"], cfg["L"], cfg["eps"]
    n = W // 2
    if n % world:
        raise SystemExit("walkers/2 must divide by the number of GPUs")
    rows = n // world
    row0 = rank * rows
    if cfg["target"] == "gauss":
        cov_t, prec = make_cov(d, cfg["kappa"])
        tgt = hx.targets.GaussianFull(precision=prec)
    elif cfg["target"] == "rosenbrock":
        tgt = hx.targets.Rosenbrock(d)
    else:
        tgt = hx.targets.Funnel(d)
    move = hx.hmc_walk_move if cfg["move"] == "walk" else hx.hmc_side_move
    k_init, k_run = hx.random.split(hx.random.PRNGKey(0), 2)
    x0 = hx.random.normal(k_init, (W, d), tdt)
    if cfg["target"] == "rosenbrock":
        x0 = x0 * 0.3 + 1.0
    lib = _lib.load()
    ctx = _rt.ctx(dev)
    for kv in args.tune:
        w_, v_ = kv.split("=")
        _lib.check(lib.hx_ctx_set_tuning(ctx, int(w_), int(v_)), "hx_ctx_set_tuning")
    K, Wu = args.steps, args.warmup
    keys_dev = hx.random.split_device(k_run, 4 * (K + Wu))
    keys_host = keys_dev.cpu().numpy().view(np.uint32).reshape(K + Wu, 4, 2)
    X = x0.clone()
    lp = tgt.log_prob(X)
    accepts = torch.zeros(W, dtype=torch.int32, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    desc = tgt.descriptor(tdt, dev)
    mk = 0 if cfg["move"] == "walk" else 1

    # the device chain buffer holds CB iterations and is reused cyclically (every iteration still writes its 262 MB slice at c5),
    # so memory does not grow with --steps
    CB = max(1, min(max(K, Wu), 16))
    if world == 1:
        chain = torch.empty((CB, W, d), dtype=tdt, device=dev)
        clp = torch.empty((CB, W), dtype=tdt, device=dev)

        def run(t0, T):
            for c0 in range(0, T, CB):
                _lib.check(lib.hx_run_iterations(ctx, _rt.dtype_code(tdt), _rt.impl_code(), mk, 0, desc, _rt.ptr(X),
                                                 _rt.ptr(lp), n, eps, L, _rt.ptr(keys_dev[4 * (t0 + c0):]), min(CB, T - c0),
                                                 _rt.ptr(chain), _rt.ptr(clp), _rt.ptr(accepts), _rt.ptr(flag), _rt.stream(dev)), "run")
    else:
        # row-sharded: every rank holds the full state in its peer-mapped exchange region and updates its own rows; the kernels
        # store the updated rows into the peers' copies over NVLink (include/hx.h section 5)
        Xf, lpf = _rt.dist_setup(dev, dist.group.WORLD, tdt, n, d)
        Xf.copy_(X); lpf.copy_(lp)
        torch.cuda.synchronize(dev)
        dist.barrier()
        accepts = torch.zeros(2 * rows, dtype=torch.int32, device=dev)
        chain = torch.empty((CB, 2 * rows, d), dtype=tdt, device=dev)
        clp = torch.empty((CB, 2 * rows), dtype=tdt, device=dev)

        def run(t0, T):
            for c0 in range(0, T, CB):
                _lib.check(lib.hx_run_iterations_sharded(ctx, _rt.dtype_code(tdt), _rt.impl_code(), mk, 0, desc, eps, L,
                                                         _rt.ptr(keys_dev[4 * (t0 + c0):]), min(CB, T - c0), _rt.ptr(chain),
                                                         _rt.ptr(clp), _rt.ptr(accepts), _rt.ptr(flag), _rt.stream(dev)), "run_sharded")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing (`value`) --------------------------------------------------------------
    clocks = ClockSampler(local, args.clock_period)
    if rank == 0 and args.clock_period > 0:       # one sampler per box: several nvidia-smi loops contend for the driver lock and stall kernel launches
        clocks.start()
    run(0, Wu)
    barrier()
    lib.hx_ctx_timing_enable(ctx, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    clocks.mark_begin()
    e0.record()
    t_host0 = time.perf_counter()
    run(Wu, K)
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / K      # host time to enqueue one iteration (launches are async)
    e1.record()
    barrier()
    clocks.mark_end()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    kms, kn, alln = np.zeros(1), np.zeros(1, np.int64), np.zeros(1, np.int64)
    import ctypes as C
    c_ms, c_kn, c_all = C.c_double(), C.c_int64(), C.c_int64()
    lib.hx_ctx_timing_read(ctx, C.byref(c_ms), C.byref(c_kn), C.byref(c_all))
    ph_ms, ph_n = (C.c_double * 8)(), (C.c_int64 * 8)()
    lib.hx_ctx_timing_phases(ctx, ph_ms, ph_n)
    names = ["main", "prep", "basis", "leapfrog", "finish", "wait_peers", "push_rows", "mean"]
    phases = {nm: round(ph_ms[i] / max(ph_n[i], 1), 4) for i, nm in enumerate(names) if ph_n[i] > 0}
    lib.hx_ctx_timing_enable(ctx, 0)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    value = W * L * K / (ms * 1e-3)
    tot_acc = accepts.sum().to(torch.float64).reshape(1)
    if world > 1:
        dist.all_reduce(tot_acc)
    acc_rate = float(tot_acc.item()) / (W * (K + Wu))

    # ---- 64-bit checksum of the final ensemble, log-probs and accept counters (bit patterns, position-weighted, wrapping
    #      int64 arithmetic): equal numbers at --gpus 1 / 2 / 4 / 8 prove that the row-sharded runs reproduce the single-GPU
    #      chain bit for bit (same K, W, workload and leapfrog form)
    def _cks(t):
        v = t.contiguous().view(torch.int32).reshape(-1).to(torch.int64)
        w = torch.arange(v.numel(), device=v.device, dtype=torch.int64) % 65521 + 1
        return int((v * w).sum().item()) & 0xFFFFFFFFFFFFFFFF
    if world == 1:
        Xfin, lpfin, accfin = X, lp, accepts
    else:
        Xfin, lpfin = Xf, lpf                       # every rank holds the full exchanged state
        parts = [torch.empty_like(accepts) for _ in range(world)]
        dist.all_gather(parts, accepts)
        accfin = torch.cat([p_[:rows] for p_ in parts] + [p_[rows:] for p_ in parts])     # global order: group 1 rows, group 2 rows
    checksum = dict(ensemble=f"{_cks(Xfin):016x}", log_prob=f"{_cks(lpfin):016x}", accepts=f"{_cks(accfin):016x}",
                    note="position-weighted sum of the 32-bit words of the final state after warmup + steps iterations (mod 2^64)")

    # ---- roofline of the dominant kernel (fused half-move) -------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
    k_launches = max(int(c_kn.value), 1)
    k_ms = c_ms.value / k_launches
    path = lib.hx_ctx_last_path(ctx)
    ctx_sms = torch.cuda.get_device_properties(dev).multi_processor_count
    form = args.leapfrog_form
    if form == "auto":
        el_ = rows * d * 28.0 / 2.6e12           # same cost model as csrc/hx_tc_walk.cu (HX_LEAP_AUTO)
        form = ("propagator" if (L + 3) * (max(el_, 25e-6) + 16e-6 * (d / 1000.0)) > (L + 1) * (27e-6 + 15e-6 * (d / 1000.0) ** 2) + 60e-6
                + max(4 * el_, 80e-6) else "stepwise")
    extra = {}
    if path == 1:
        # tensor-core path: the bracketed section is the momentum projection S0 = P0 C of one half-move: tcgen05 GEMM chunks
        # (128 x 256 tiles, K = n) while the counter-based draw of the NEXT half-move's P0 runs concurrently on the ALU pipes
        tune = dict(kv.split("=") for kv in args.tune)
        # the momentum draw is generated inside the projection GEMM (k_tc_proj_fused) unless switched off / not eligible
        fused = (int(tune.get("15", 1)) != 0 and args.precision in ("auto", "bf16x3", "tf32x3") and d <= 1024
                 and _rt.impl_code() == _lib.HX_RNG_PARTITIONABLE)
        f8 = (not fused) and args.precision in ("auto", "f16f8")        # staged projection operands: fp16 + 2 x e4m3, or bf16 hi/lo
        ffmt = int(tune.get("21", 0)) or (2 if args.precision == "auto" else 1)   # fused operand form: 1 bf16 hi/lo, 2 fp16 + fp8 (hx_tc_fused.cuh)
        ekind = int(tune.get("3", 2)) if args.precision != "tf32x3" else 0   # propagator energy operands: 2 fp16 hi/lo, 0 tf32 hi/lo
        if fused:
            kname = ("tc::k_tc_proj_fused<%s RNG warps, %s, %s> S0 = normal(key,(n,n))[rows] @ C: threefry + erf_inv in registers -> "
                     "operand tiles in shared memory -> tcgen05.mma into one TMEM accumulator, split-K 2 added with red.global.add"
                     % (tune.get("16", "24" if d > 512 else "28"), "CTA pair" if d > 512 else "single CTA",
                        "fp16 + e4m3 x e5m2 operands at true scale (kind::f16 + 2 x kind::f8f6f4)" if ffmt == 2 else "bf16 hi/lo operands (3 x kind::f16)"))
        else:
            kname = ("tc::k_tc_gemm<256,EPI_STORE,%s> projection S0 = P0 C (all row chunks of a half-move), "
                     "tc::k_rng_planes of the next half-move concurrent" % ("fp16 + e4m3" if f8 else "bf16 hi/lo"))
        flops = 2.0 * rows * n * d
        nK, dpad = -(-n // 64) * 64, -(-d // 256) * 256
        # tensor work actually issued, in units of one bf16-rate product: bf16 hi/lo = 3 products; fp16 + e4m3 = one fp16 product
        # + two e4m3 products at twice the rate = 2 units; single bf16 = 1
        units = 1.0 if args.precision == "bf16" else (2.0 if (f8 or (fused and ffmt == 2)) else 3.0)
        executed = units * 2.0 * (-(-rows // 128) * 128) * nK * dpad
        dB, dK = dpad, -(-d // 64) * 64
        # executed tensor flops of the whole half-move in bf16 units (tf32 products count double: half the MMA rate)
        ef = 2.0 if ekind == 0 else 1.0
        kf = 2.0 if args.precision == "tf32x3" else 1.0
        if form == "propagator":
            lf = kf * 3 * 2.0 * rows * (2 * dK) * (2 * dB) + ef * 3 * (2.0 * rows * (2 * dK) * dB + 2.0 * rows * dK * dB)
            lf += kf * 3 * (L + 1) * 2.0 * (2 * d) * dK * dB + 2 * 3 * 2.0 * (2 * d) * dK * dB
        else:
            lf = kf * 3 * (L + 3) * 2.0 * rows * dK * dB
        gram = 3 * 2.0 * dB * dB * nK
        step_exec = executed + lf + gram
        extra = dict(executed_tflops=executed / (k_ms * 1e-3) / 1e12, executed_frac=executed / (k_ms * 1e-3) / 1e12 / peak_tf,
                     step_executed_tflops=step_exec * 2 * K / (ms * 1e-3) / 1e12,
                     step_executed_frac=step_exec * 2 * K / (ms * 1e-3) / 1e12 / peak_tf)
        note = ("achieved/frac = ALGORITHMIC flops 2*rows*n*d of the projection (SURVEY §8d first term) / section time. fp32-class "
                "accuracy on 16-bit tensor-core operands takes 3 products per algorithmic product (hi*hi + lo*hi + hi*lo: frac <= 1/3 "
                "by construction; the fp16 + fp8 forms run their two correction products at twice the rate, 2 units: frac <= 1/2); "
                "executed_* counts the bf16-rate units actually issued (incl. d padded to 256); step_executed_* is the whole "
                "iteration (projection + leapfrog/energy GEMMs + Gram; tf32 products counted double) against the same peak.  The "
                "section is bound by the momentum draw it contains (n^2 threefry + erf_inv evaluations per half-move on the alu "
                "pipe: draw_* keys -- draw_frac is the fraction of THAT roofline, the one that binds; 3.5 ms for the arithmetic alone at c5 on "
                "one GPU at 1965 MHz), not by the tensor pipe, and the "
                "whole step runs under the 1 kW power cap (clocks)")
        # the draw inside the section: achieved normals/s against the measured rate of the draw arithmetic alone (threefry2x32-20
        # + XLA erf_inv + operand split, 96 instructions per normal: 120.5 clk per warp and normal and scheduler at 1965 MHz on
        # every SM, scripts/micro/normals_micro.cu -- alu pipe (SHF / LOP3 at 2 clk) and issue slots saturate together)
        draw_peak = ctx_sms * 4 * 32 / 120.5 * 1.965e9
        extra.update(draw_gnormals_per_s=rows * float(n) / (k_ms * 1e-3) / 1e9, draw_peak_gnormals_per_s=draw_peak / 1e9,
                     draw_frac=rows * float(n) / (k_ms * 1e-3) / draw_peak,
                     draw_peak_source="measured: draw arithmetic alone on all SMs, scripts/micro/normals_micro.cu (round 2)")
        if args.workload == "c5" and world == 1:
            if fused:
                extra["traffic_source"] = FUSED_TRAFFIC_SOURCE
            else:
                extra["traffic_source"] = ("profiles/r01e_c5_ncu_full_summary.txt: dram bytes read+write of the 7 chunk launches of one "
                                           "half-move (6 x 755.0 MB + 704.7 MB read, 28 MB written) = the P0 planes + C^T planes, "
                                           "i.e. no re-reads (same plane sizes for bf16 hi/lo: profiles/r01c)")
    else:
        kname = "k_walk_half" if mk == 0 else "k_side_half"
        flops = algorithmic_flops_half_move(rows, n, d, L, cfg["target"])
        note = "algorithmic flops = projected form 2*rows*n*d + (L+1)*(2*rows*d^2 [M] + grad) (SURVEY §8d)"
    achieved = flops / (k_ms * 1e-3) / 1e12
    traffic = None
    if "traffic_source" in extra:
        traffic = FUSED_TRAFFIC_BYTES if extra["traffic_source"] is FUSED_TRAFFIC_SOURCE else 5.263e9
    if path == 1 or mk == 1:
        roofline = dict(bound="tensor", kernel=kname, achieved=achieved, peak=peak_tf,
                        unit="TFLOP/s", frac=achieved / peak_tf,
                        traffic=traffic, peak_source=peak_src,
                        flops_per_launch=flops, kernel_ms=k_ms, kernel_share_of_step=c_ms.value / ms, note=note,
                        half_move_algorithmic_tflops=algorithmic_flops_half_move(rows, n, d, L, cfg["target"]) * 2 * K
                        / (ms * 1e-3) / 1e12, **extra)
        if mk == 1:
            roofline["note"] = ("side move: RNG / selection-bound (n sort keys per walker and shuffle round), SURVEY section 8d; the tensor "
                                "fraction is reported for completeness only")
    else:
        # SIMT walk half-move (c1-c4 sizes): ALU / RNG-bound -- rows * n normals per launch (n / L normals per walker-step against
        # <= 2 d^2 + O(d) flops), SURVEY section 8d.  Roofline = normals per second against the measured rate of the draw
        # arithmetic alone on all SMs (scripts/micro/normals_micro.cu).
        draw_peak = ctx_sms * 4 * 32 / 120.5 * 1.965e9
        half_ms = ms / K / 2.0            # whole half-move (narrow targets draw in a separate projection kernel: not only k_ms)
        gn = rows * float(n) / (half_ms * 1e-3)
        roofline = dict(bound="alu", kernel=kname + " (RNG-fused projection: threefry2x32 + erf_inv per momentum element)",
                        achieved=gn / 1e9, peak=draw_peak / 1e9, unit="Gnormal/s", frac=gn / draw_peak, traffic=None,
                        peak_source="measured: draw arithmetic alone (96 instructions per normal, 120.5 clk per warp and scheduler at "
                                    "1965 MHz) on all SMs, scripts/micro/normals_micro.cu (round 2)",
                        normals_per_launch=rows * float(n), half_move_ms=half_ms, kernel_ms=k_ms, kernel_share_of_step=c_ms.value / ms,
                        tensor_equivalent=dict(achieved_tflops=achieved, frac_of_bf16_peak=achieved / peak_tf, note=note),
                        note="achieved = normals of one half-move / time of the whole half-move (draw, projection, L + 1 kicks, "
                             "log-density, accept); latency-bound below ~1e6 normals per half-move (c1, c2: a handful of CTAs)")

    # ---- end-to-end through the public API with host buffers (`e2e`) -----------------------------------
    e2e = None
    # e2e keeps the reference's store-everything semantics, so the host chain is K x W x d: cap K by ~8 GB of pinned host memory
    isz_ = 8 if cfg["x64"] else 4
    Ke = max(1, min(K, int(8e9 // (W * d * isz_ / max(world, 1)))))
    if world == 1 and not args.no_e2e:
        x0_host = x0.cpu().pin_memory()
        s = hx.HamiltonianEnsembleSampler(W, d, tgt, step_size=eps, L=L, move=move, adapt_inital_step_size=False,
                                          adapt_step_size=False, adapt_length=False, verbose=False)
        # untimed warm-up run of the same size: page-locking the host chain buffer costs ~0.64 ms/MB the first time
        # (0.84 s for 5 slices of c5); afterwards torch's caching host allocator reuses the block, which is the
        # steady state of a user who samples repeatedly
        w = s.run_mcmc(k_run, x0_host, Ke, warmup=0, batch_size=1)
        del w, s
        import gc
        gc.collect()
        torch.cuda.synchronize(dev)
        s = hx.HamiltonianEnsembleSampler(W, d, tgt, step_size=eps, L=L, move=move, adapt_inital_step_size=False,
                                          adapt_step_size=False, adapt_length=False, verbose=False)
        t0 = time.perf_counter()
        out = s.run_mcmc(k_run, x0_host, Ke, warmup=0, batch_size=1)
        torch.cuda.synchronize(dev)
        el = time.perf_counter() - t0
        isz = 8 if cfg["x64"] else 4
        e2e = dict(value=W * L * Ke / el, unit="walker-steps/s",
                   h2d_bytes_per_step=int(W * d * isz / Ke + 32), d2h_bytes_per_step=int(W * d * isz + W * isz),
                   seconds=el, steps=Ke, api="HamiltonianEnsembleSampler.run_mcmc (host initial_state -> host chain)",
                   note="h2d = initial ensemble amortised over the steps + 4 keys; d2h = the step's chain slice and log-probs; "
                        "pinned host chain buffer reused from an untimed warm-up run of the same size")
        assert out.shape == (Ke, W, d)
    elif world > 1 and not args.no_e2e:
        # every rank streams the chain of ITS rows to its own pinned host buffer (chain_store='local'); max over ranks
        x0_host = x0.cpu().pin_memory()
        mkS = lambda: hx.HamiltonianEnsembleSampler(W, d, tgt, step_size=eps, L=L, move=move, adapt_inital_step_size=False,
                                                    adapt_step_size=False, adapt_length=False, verbose=False,
                                                    group=dist.group.WORLD, chain_store="local", exchange="peer")
        s = mkS()
        w = s.run_mcmc(k_run, x0_host, Ke, warmup=0, batch_size=1)
        del w, s
        import gc
        gc.collect()
        s = mkS()
        barrier()
        t0 = time.perf_counter()
        out = s.run_mcmc(k_run, x0_host, Ke, warmup=0, batch_size=1)
        torch.cuda.synchronize(dev)
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        el = float(el[0])
        isz = 8 if cfg["x64"] else 4
        e2e = dict(value=W * L * Ke / el, unit="walker-steps/s", h2d_bytes_per_step=int(W * d * isz / Ke + 32),
                   d2h_bytes_per_step=int((W * d * isz + W * isz) / world), seconds=el, steps=Ke,
                   api="HamiltonianEnsembleSampler.run_mcmc(group=..., chain_store='local') (host initial_state -> per-rank host chain)",
                   note="per rank: h2d = the initial ensemble amortised over the steps + keys, d2h = the chain slice and log-probs of "
                        "the rank's own rows; time = max over ranks")
        assert out.shape == (Ke, 2 * rows, d)

    # ---- ESS/s leg (BASELINE metric): tau_int via the reference's estimator (autocorr.py:44-118, c=5, tol=50,
    #      quiet) on a fixed subset of <= 16 dims x <= 1024 walkers, after a burn-in; device-resident state, untimed by `value`
    ess = None
    if world == 1 and args.ess_iters > 0 and not args.no_e2e:
        # no chain leaves the kernels here: the on-device accumulators (hemcee_b200.stats.ChainStats, csrc/hx_stats.cu) are fed
        # after every iteration inside hx_run_iterations and give the reference estimator's tau for lags < max_lag
        from hemcee_b200.stats import ChainStats
        Tb, Te = 48, args.ess_iters
        Xe = x0.clone(); lpe = tgt.log_prob(Xe)
        ke = hx.random.split_device(hx.random.PRNGKey(12345), 4 * (Tb + Te))
        acc_e = torch.zeros(W, dtype=torch.int32, device=dev)
        dims = [int(v) for v in np.linspace(0, d - 1, min(16, d)).round()]
        st = ChainStats(W, d, tdt, dev, max_lag=min(96, Te), walkers=min(1024, W), dims=dims)

        def run_e(t0, T):
            _lib.check(lib.hx_run_iterations(ctx, _rt.dtype_code(tdt), _rt.impl_code(), mk, 0, desc, _rt.ptr(Xe), _rt.ptr(lpe), n,
                                             eps, L, _rt.ptr(ke[4 * t0:]), T, None, None, _rt.ptr(acc_e),
                                             _rt.ptr(flag), _rt.stream(dev)), "run_ess")
        run_e(0, Tb)
        torch.cuda.synchronize(dev)
        st.attach()
        t0 = time.perf_counter()
        try:
            run_e(Tb, Te)
            torch.cuda.synchronize(dev)
        finally:
            st.detach()
        el = time.perf_counter() - t0
        tau = np.asarray(st.integrated_time(c=5, tol=50, quiet=True))
        tmax = float(np.nanmax(tau))
        ess = dict(value=Te * W / tmax / el, unit="ESS/s", tau_max=tmax, tau_min=float(np.nanmin(tau)), iterations=Te, burn_in=Tb,
                   seconds=el, reliable=bool(50 * tmax <= Te), dims=len(dims), walkers_in_estimate=int(st.n_track),
                   mean_abs_max=float(np.abs(st.mean()).max()), var_minmax=[float(st.var().min()), float(st.var().max())],
                   var_over_target_minmax=([float((st.var() / np.diag(cov_t)).min()), float((st.var() / np.diag(cov_t)).max())]
                                           if cfg["target"] == "gauss" else None),
                   note="min over the sampled dims of (iterations * walkers / tau_int) / wall time of those iterations; tau_int = the "
                        "reference's estimator (autocorr.py:44-118, c=5, tol=50) evaluated from on-device accumulators (no chain is "
                        "stored or copied: hx_ctx_set_chain_stats); fixed step size and L (no adaptation), chain started from "
                        "N(0, I); mean / var = running moments over all walkers and those iterations; var_over_target = their ratio to the target's "
                        "coordinate variances diag(Sigma) (long-run moments at full size: ~1 once the slow directions have equilibrated)")

    # ESS/s at N > 1: the sharded loop continues from the timed state; every rank estimates tau_int on a subset of ITS rows
    if world > 1 and args.ess_iters > 0 and not args.no_e2e:
        from hemcee_b200 import autocorr as hac
        Tb, Te = 16, args.ess_iters
        Xf.copy_(x0); lpf.copy_(tgt.log_prob(Xf))
        torch.cuda.synchronize(dev)
        dist.barrier()
        ke = hx.random.split_device(hx.random.PRNGKey(12345), 4 * (Tb + Te))
        dims = torch.linspace(0, d - 1, min(16, d)).long().to(dev)
        wsel = torch.linspace(0, 2 * rows - 1, min(max(1024 // world, 64), 2 * rows)).long().to(dev)
        sub = torch.empty((Te, wsel.numel(), dims.numel()), dtype=tdt, device=dev)

        def run_s(t0, T, store):
            for c0 in range(0, T, CB):
                nb_ = min(CB, T - c0)
                _lib.check(lib.hx_run_iterations_sharded(ctx, _rt.dtype_code(tdt), _rt.impl_code(), mk, 0, desc, eps, L,
                                                         _rt.ptr(ke[4 * (t0 + c0):]), nb_, _rt.ptr(chain), _rt.ptr(clp), _rt.ptr(accepts),
                                                         _rt.ptr(flag), _rt.stream(dev)), "run_ess_sharded")
                if store:
                    sub[c0:c0 + nb_] = chain[:nb_][:, wsel][:, :, dims]
        run_s(0, Tb + 32, False)             # burn-in (the ensemble starts from N(0, I))
        barrier()
        t0 = time.perf_counter()
        run_s(Tb + 32, Te - 32, True)
        barrier()
        el = time.perf_counter() - t0
        Tm = Te - 32
        tau = np.asarray(hac.integrated_time(sub[:Tm], c=5, tol=50, quiet=True))
        tt = torch.tensor([float(np.nanmax(tau))], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tmax = float(tt[0])
        ess = dict(value=Tm * W / tmax / el, unit="ESS/s", tau_max=tmax, iterations=Tm, burn_in=Tb + 32, seconds=el,
                   reliable=bool(50 * tmax <= Tm), dims=int(dims.numel()), walkers_in_estimate=int(wsel.numel()) * world,
                   note="max tau_int over ranks (each on a subset of its own rows); iterations * walkers / tau / wall time")

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_reference(args.workload, 1000, 1, budget_s=15.0)

    if rank == 0:
        line = dict(metric="leapfrog_walker_steps_per_sec", value=value, unit="walker-steps/s", n_gpus=world, steps=K,
                    warmup=Wu, ms_per_step=ms / K, higher_is_better=True, scaling="strong", vs_baseline=None,
                    dtype="f64" if cfg["x64"] else "f32", data="synthetic", config=conf, roofline=roofline,
                    cpu_baseline=cb, clocks=clk, e2e=e2e, gpu_launches=int(c_all.value) if world == 1 else int(c_all.value),
                    acceptance_rate=acc_rate, nonfinite=int(flag.item()), precision=args.precision, leapfrog_form=args.leapfrog_form,
                    phase_ms_per_half_move=phases, ess=ess, host_enqueue_ms_per_step=host_enqueue_ms, checksum=checksum)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
