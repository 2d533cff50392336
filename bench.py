#!/usr/bin/env python
"""Benchmark of the batched RJ-MCMC hot path (BASELINE.json metric: chain-steps/sec on 2-D Voronoi ray
tomography).

    python bench.py --gpus N --steps K --warmup W            # CUDA arm (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # reference CPU arm (oracle port, host cores)

One "step" = one RJ-MCMC iteration of EVERY chain of the job (proposal -> rasterisation of the proposed
geometry -> straight-ray forward + misfit -> Metropolis-Hastings decision), i.e. ``n_chains`` chain-steps.
The engine runs its chain-local step (one thread block per chain: change-local rasterisation + change-local
forward + misfit + decision in one kernel, ``bayesbay_b200/csrc/chain_kernels.cu``); ``BBB_FULL_FORWARD=1``
selects the full-SpMM step (every proposal re-evaluates all rays, as the reference does).
Workload at N = 1: BASELINE.json configs[2] (the configuration the metric is quoted on): 100x100 grid,
10 000 rays, k in [10, 200], hierarchical noise, 1024 chains.  N > 1: the same problem with 1024 chains per
GPU (weak scaling), the 16-temperature ladder of configs[3] and one NCCL swap round per 100 steps.

Prints ONE JSON line (see the keys below); everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "chain-steps/sec, 2D Voronoi ray tomography, 1/2/4/8 B200 vs host CPU"
UNIT = "chain-steps/s"
CHAINS_PER_GPU = 1024
SWAP_EVERY = 100
WORKLOAD = ("BASELINE configs[2]: Voronoi2D straight-ray travel-time tomography, 100x100 grid, 10000 rays "
            "(nnz 828141), k in [10,200], hierarchical noise, 1024 chains per GPU")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def build_problem():
    from bayesbay_b200 import synthetic

    return synthetic.tomography_2d()


def algorithmic_bytes(ing, n_chains):
    """B = 16 G + 16 R + 12 nnz / C + 48 K per chain-step (fp64 values, int32 indices);
    the forward + likelihood kernel's share is 8 G + 16 R + (12 nnz + 4 (R + 1) + 8 R) / C + 16."""
    G = ing["points"].shape[0]
    R = ing["indptr"].size - 1
    nnz = ing["indices"].size
    K = ing["voronoi"]["n_dimensions_max"]
    total = 16 * G + 16 * R + 12 * nnz / n_chains + 48 * K
    fwd = 8 * G + 16 * R + (12 * nnz + 4 * (R + 1) + 8 * R) / n_chains + 16
    raster = 8 * G + 24 * K
    # the chain-local kernel does rasterisation + forward + likelihood of a chain-step in one launch
    return dict(total=total, forward_likelihood=fwd, raster=raster, chain_step=raster + fwd)


# --------------------------------------------------------------------------- #
# clocks
# --------------------------------------------------------------------------- #
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception as e:  # pragma: no cover
            log("clock sampler unavailable:", e)

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------------------- #
# CUDA arm
# --------------------------------------------------------------------------- #
def cuda_arm(args):
    import torch
    import torch.distributed as dist

    import bayesbay_b200 as bb
    from bayesbay_b200 import _compile, synthetic
    from bayesbay_b200._engine import Engine, pt_swap
    from bayesbay_b200.forward import RayTomographyForward

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    device = torch.device(f"cuda:{local_rank}")
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}")

    ing = build_problem()
    # the problem is described through the reference-facing API and compiled for the device
    fwd = RayTomographyForward((ing["indptr"], ing["indices"], ing["data"]), ing["points"], "voronoi", "vel")
    par, ll = synthetic.build_problem(ing, bb, lambda _: fwd)
    spec = _compile.compile_problem(par, ll)
    C = CHAINS_PER_GPU
    eng = Engine(spec, C, device=device, seed=12345, chain_id0=rank * C)
    eng.init_states()
    tempered = world > 1
    if tempered:
        # interleaved ladder: every rank holds all 16 temperatures, so the ranks stay load balanced (hot chains
        # accept more births and carry more cells)
        ladder = np.tile(np.geomspace(1.0, 5.0, 16), world * C // 16)
        eng.temperature.copy_(torch.as_tensor(ladder[rank * C:(rank + 1) * C]))

    def swap_round(r):
        m, ld = eng.misfit_logdet()
        local = torch.stack((m, ld, eng.temperature), dim=1).contiguous()
        if world > 1:
            gathered = torch.empty((world * C, 3), dtype=torch.float64, device=device)
            dist.all_gather_into_tensor(gathered, local)
        else:
            gathered = local.clone()
        t_out, _ = pt_swap(gathered, 777, r, rank * C, C)
        eng.temperature.copy_(t_out)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    def sync_all():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    # burn the chains in a little so k and sigma are in a representative regime, then W warm-up steps
    eng.step(max(args.burn, 0))
    for _ in range(args.warmup):
        eng.step(1)
    if tempered:
        swap_round(10 ** 6)  # warm the NCCL all-gather path (untimed)
    sync_all()

    # ---- timed region: exactly K steps; L2 flushed between steps, each step timed by its own events --
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    wall0 = time.perf_counter()
    n_swaps = 0
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ev[i][0].record()
        eng.step(1)
        if tempered and ((i + 1) % min(SWAP_EVERY, args.steps) == 0):
            swap_round(n_swaps)
            n_swaps += 1
        ev[i][1].record()
    sync_all()
    wall = time.perf_counter() - wall0
    clk = clocks.stop() if rank == 0 else None
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    total_ms = float(step_ms.sum())
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * C * args.steps / (total_ms * 1e-3)

    # ---- same K steps back to back without the flush (L2-warm steady state), for information ---------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    eng.step(args.steps)
    e1.record()
    sync_all()
    warm_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([warm_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        warm_ms = float(t.item())

    # ---- per-kernel breakdown + roofline of the dominant kernel (rank 0) -----------------------------
    # stages of bbb_engine_stage, full-SpMM step: 0 = k_propose, 1 = k_idx_commit_copy + k_raster2d_inc +
    # k_raster2d_rescan, 2 = k_ray_spmm, 3 = k_finish  (6 kernel launches per step);
    # chain-local step: 0 = k_propose, 2 = k_chain_step, 1 and 3 launch nothing (2 launches per step, plus the
    # from-scratch re-evaluation of residuals and misfit every 256 steps: 5 launches)
    names = ["propose", "raster2d", "chain_step" if eng.chain_local else "ray_spmm", "finish"]
    stage_ms = np.zeros(4)
    active_fwd = 0
    active_geom = 0
    if rank == 0:
        sev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(4)]
               for _ in range(args.steps)]
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
            for s in range(4):
                sev[i][s][0].record()
                eng.stage(s)
                sev[i][s][1].record()
                if s == 0:
                    fin = torch.isfinite(eng.log_prob_ratio) & (eng.move != 4)
                    active_fwd += int(fin.sum().item())
                    active_geom += int((fin & ((eng.move & 3) != 2)).sum().item())
        torch.cuda.synchronize(device)
        for i in range(args.steps):
            for s in range(4):
                stage_ms[s] += sev[i][s][0].elapsed_time(sev[i][s][1])
        stage_ms /= args.steps
    sync_all()

    # ---- end to end through the plugin call with HOST buffers ----------------------------------------
    # One call = what the reference's pool does per advance_chain: chain states (incl. their caches: misfit and
    # nearest-site index, the analogue of the pickled d_pred / KD-tree caches) go host -> device from pinned
    # memory, the chains advance one iteration, the updated chains come back device -> host.
    e2e = e2e_run(eng, args.steps, device, world, dist if world > 1 else None)

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ab = algorithmic_bytes(ing, C)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    dom = int(np.argmax(stage_ms))
    units = {"ray_spmm": active_fwd / args.steps, "chain_step": active_fwd / args.steps,
             "raster2d": active_geom / args.steps}.get(names[dom], C)
    per_unit = {"ray_spmm": ab["forward_likelihood"], "chain_step": ab["chain_step"],
                "raster2d": ab["raster"]}.get(names[dom], 48 * 200)
    achieved = per_unit * units / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
    traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(names[dom])
        except Exception:
            traffic = None
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "chains_per_gpu": C, "n_chains": world * C,
                   "l2": "flushed between timed steps (256 MiB write); each step timed by its own CUDA events",
                   "tempering": ("16-temperature ladder, NCCL all-gather swap round every %d steps" % min(SWAP_EVERY, args.steps))
                   if tempered else "none (all chains T=1)",
                   "burn_steps": args.burn},
        "value_l2_warm": world * C * args.steps / (warm_ms * 1e-3),
        "ms_per_step_l2_warm": warm_ms / args.steps,
        "wall_s_timed_region": wall,
        "step_ms_min_median_max": [float(step_ms.min()), float(np.median(step_ms)), float(step_ms.max())],
        "clocks": {k: clk[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")} if clk else None,
        "e2e": e2e,
        "gpu_launches": (2 if eng.chain_local else 6) * args.steps + (2 * n_swaps if tempered else 0),
        "step_kind": "chain-local (k_propose, k_chain_step)" if eng.chain_local
        else "full SpMM (k_propose, k_idx_commit_copy, k_raster2d_inc, k_raster2d_rescan, k_ray_spmm, k_finish)",
        "kernel_ms": {n: float(v) for n, v in zip(names, stage_ms) if not (eng.chain_local and n in ("raster2d", "finish"))},
        "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_chain_step": per_unit, "chain_steps_per_launch": units,
                     "whole_step_frac": ab["total"] * C / (warm_ms / args.steps * 1e-3) / 1e9 / peak},
    }
    if world == 1:
        out["api_run"] = api_run(par, ll, bb, C)
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(spec, args.cpu_seconds)
    emit(out)
    if world > 1:
        dist.destroy_process_group()


def api_run(par, ll, bb, n_chains, n_iterations=2000, save_every=100):
    """Wall clock of the call a user makes -- BayesianInversion.run with the reference's defaults for what is saved
    (every chain's state AND d_pred at every save point come back to host lists, as in the reference) -- for
    information; not the device-timed `value`."""
    import torch

    inv = bb.BayesianInversion(par, ll, n_chains=n_chains)
    inv.run(n_iterations=200, burnin_iterations=100, save_every=save_every, verbose=False)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    inv.run(n_iterations=n_iterations, burnin_iterations=0, save_every=save_every, verbose=False)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    n_samples = len(inv.get_results()["voronoi.n_dimensions"])
    return {"value": n_chains * n_iterations / wall, "unit": UNIT, "wall_s": wall,
            "what": f"BayesianInversion.run(n_iterations={n_iterations}, save_every={save_every}) on the same problem and "
                    f"chain count, host wall clock incl. the saved samples ({n_samples} states with d_pred) coming back"}


def e2e_run(eng, steps, device, world, dist):
    """The plugin call with HOST buffers, one iteration per call (bbb_engine_step_hosted + _wait): the chains' states
    (k, sites, values, noise sigma, temperature, iteration) go pinned host -> device, the chains advance one
    iteration, the updated states come back device -> host and the host waits for them.  The chain set travels in
    parts on the engine's internal streams, so one part's copies overlap the other parts' kernels.  The derived
    per-chain caches (nearest-site index, residuals, misfit) stay resident on the device between calls and are used
    only after the uploaded states were compared, on the device, with the states the caches belong to
    (bit-identical -> the part's kernels run; else they are no-ops and the wait re-derives the caches with
    bbb_engine_refresh and advances the part) -- the analogue of the d_pred / KD-tree caches the reference pickles
    along with every chain."""
    import torch

    K = eng.spec["spaces"][0]["kmax"]
    cap = eng.state_image_bytes(K) // 8
    host = [torch.empty((cap,), dtype=torch.int64, pin_memory=True) for _ in range(2)]  # the caller's chains, in / out
    stage = torch.empty((cap,), dtype=torch.int64, device=device)     # upload buffer
    out = torch.empty((cap,), dtype=torch.int64, device=device)       # download buffer
    refreshed = [0]
    moved = [0, 0]

    kb = min(K, int(eng.gather_space(0)[0].max().item()))
    n0 = eng.state_image_bytes(kb) // 8
    host[0][:n0].copy_(eng.pack_states_hosted(kb))
    torch.cuda.synchronize(device)
    cur = 0

    def call(kb_up, cur):
        kb_down = eng.step_hosted(kb_up, host[cur], host[cur ^ 1], stage, out)[0]
        refreshed[0] += eng.step_hosted_wait()  # the host reads the step's result before the next call
        moved[0] += eng.state_image_bytes(kb_up)
        moved[1] += eng.state_image_bytes(kb_down)
        return kb_down

    for _ in range(3):
        kb = call(kb, cur)
        cur ^= 1
    moved[0] = moved[1] = 0
    refreshed[0] = 0
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        kb = call(kb, cur)
        cur ^= 1
    e1.record()
    torch.cuda.synchronize(device)
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"value": world * eng.C * steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": moved[0] // steps,
            "d2h_bytes_per_step": moved[1] // steps, "ms_per_step": ms / steps, "host_wall_ms_per_step": wall_ms / steps,
            "cache_refreshes": refreshed[0], "parts": len(eng.hosted_parts()) - 1,
            "what": "per call (bbb_engine_step_hosted + bbb_engine_step_hosted_wait): packed image of the chains' states in "
                    "pinned host memory (k; sites and values of the cells in use; sigma, temperature, iteration) -> device, "
                    "compared on the device with the resident states the caches belong to, one iteration, packed image -> "
                    "pinned host memory, host waits; the chain set travels in parts on internal streams (copies of one part "
                    "overlap the kernels of the others)"}


# --------------------------------------------------------------------------- #
# CPU baseline / reference arm
# --------------------------------------------------------------------------- #
def cpu_baseline(spec, seconds):
    from oracle import cpu_baseline as cb

    cores = cb.host_cores()
    n = cb.steps_for_budget(spec, seconds)
    rate, procs, wall = cb.run_pool(spec, n, cores)
    return {"value": rate, "unit": UNIT, "cores": procs, "kind": "port",
            "sample": f"{procs} chains (one process per host core) x {n} iterations of the same C3 problem, "
                      f"{wall:.1f} s wall; oracle port of the reference chain step (KD-tree + scipy CSR forward)"}


def reference_arm(args):
    """Times the reference's CPU path (oracle port: the Python reference cannot travel to the GPU box) on all
    host cores, on the CUDA arm's config, metric and unit.  Under torchrun only rank 0 works."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    from tests import problems

    ing = build_problem()
    spec = problems.spec_from_ingredients(ing)
    from oracle import cpu_baseline as cb

    cores = cb.host_cores()
    per_step_budget = max(2.0, min(20.0, args.cpu_seconds))
    n = cb.steps_for_budget(spec, per_step_budget)
    rates = []
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cb.run_pool(spec, max(5, n // 10), cores)
    t_steps = []
    for _ in range(max(1, min(args.steps, 3))):
        t0 = time.perf_counter()
        rate, procs, wall = cb.run_pool(spec, n, cores)
        rates.append(rate)
        t_steps.append(time.perf_counter() - t0)
    value = float(np.mean(rates))
    sample = (f"each step = {cores} chains (one process per host core) x {n} iterations of the C3 problem; "
              f"{len(rates)} such samples timed (bounded so the run ends within minutes)")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(rates),
        "warmup": 1, "ms_per_step": float(np.mean(t_steps)) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "chains": cores, "iterations_per_sample": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


def _claim_stdout():
    """The contract is ONE JSON line on stdout, but libraries (NCCL's version banner) write to file descriptor 1 too:
    point fd 1 at stderr for the whole run and keep the real stdout for the JSON line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


_JSON_OUT = sys.stdout


def emit(obj):
    """The one JSON line of the contract, on the real stdout."""
    _JSON_OUT.write(json.dumps(obj) + "\n")
    _JSON_OUT.flush()


def main():
    global _JSON_OUT
    _JSON_OUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--burn", type=int, default=300, help="untimed iterations before the warm-up (state in a typical regime)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline budget per process")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        cuda_arm(args)


if __name__ == "__main__":
    main()
