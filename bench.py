#!/usr/bin/env python
"""Benchmark of the batched RJ-MCMC hot path (BASELINE.json metric: chain-steps/sec on 2-D Voronoi ray
tomography).

    python bench.py --gpus N --steps K --warmup W            # CUDA arm (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the UNMODIFIED reference's chain pool on the host cores

One "step" = one RJ-MCMC iteration of EVERY chain of the job (proposal -> rasterisation of the proposed
geometry -> straight-ray forward + misfit -> Metropolis-Hastings decision), i.e. ``n_chains`` chain-steps.
The engine runs its chain-local step (``k_propose`` + ``k_chain_step``: one thread block per chain does the
change-local rasterisation + change-local forward + misfit + decision, ``bayesbay_b200/csrc/chain_kernels.cu``);
``BBB_FULL_FORWARD=1`` selects the full-SpMM step (every proposal re-evaluates all rays, as the reference does).

Workloads
  N = 1  ``value``: BASELINE.json configs[2], the configuration the metric is quoted on: 100x100 grid, 10 000 rays,
         k in [10, 200], hierarchical noise, 1024 chains, all at T = 1.  Extra key ``tempered``: the same chains under
         the 16-temperature ladder with swap rounds (the per-GPU workload of the N > 1 lines).
  N > 1  ``value``: configs[3] -- the same problem, 1024 chains per GPU (weak scaling), 16-temperature ladder, one
         NCCL all-gather + swap round per SWAP_EVERY steps.  Extra key ``replicas``: the chains of every rank at
         T = 1 without any exchange (configs[2] per GPU).
  every N, extra key ``configs``: c5 = configs[4] (256x256 grid, 100 000 rays, k in [50, 500], 1024 chains per GPU),
         c1 = configs[0] (Voronoi1D, 4 chains), c2 = configs[1] (polynomial, 4096 chains).

Prints ONE JSON line (see the keys below); everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "chain-steps/sec, 2D Voronoi ray tomography, 1/2/4/8 B200 vs host CPU"
UNIT = "chain-steps/s"
CHAINS_PER_GPU = 1024
SWAP_EVERY = 20
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


def load_json(*rel):
    try:
        return json.load(open(os.path.join(REPO, *rel)))
    except Exception:
        return {}


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
class Ctx:
    """Process-wide handles of the CUDA arm."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.rank = int(os.environ.get("RANK", 0))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{self.local_rank}"))
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device(f"cuda:{self.local_rank}")
        if self.world != args.gpus:
            log(f"warning: --gpus {args.gpus} but WORLD_SIZE {self.world}")
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.device)  # > 126 MB L2

    def sync_all(self):
        self.torch.cuda.synchronize(self.device)
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize(self.device)

    def max_over_ranks(self, ms):
        if self.world > 1:
            t = self.torch.tensor([ms], dtype=self.torch.float64, device=self.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return float(t.item())
        return float(ms)

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)


class SwapRound:
    """The parallel-tempering exchange of a sharded job: ONE kernel writes the rank's [C][3] (misfit, logdet, T) rows,
    an NCCL all-gather into a preallocated buffer (N > 1), the swap kernel replays the reference's sequential attempts
    and writes the rank's new temperatures straight into the engine's temperature buffer."""

    def __init__(self, ctx, eng, C):
        import torch

        from bayesbay_b200._engine import pt_swap

        self.ctx, self.eng, self.C, self.pt_swap = ctx, eng, C, pt_swap
        self.local = torch.empty((C, 3), dtype=torch.float64, device=ctx.device)
        self.gathered = torch.empty((ctx.world * C, 3), dtype=torch.float64, device=ctx.device) if ctx.world > 1 else self.local

    def __call__(self, r):
        eng, ctx = self.eng, self.ctx
        eng.misfit_logdet_T(self.local)
        if ctx.world > 1:
            ctx.dist.all_gather_into_tensor(self.gathered, self.local)
        self.pt_swap(self.gathered, 777, r, ctx.rank * self.C, self.C, out=eng.temperature)


def timed_steps(ctx, eng, steps, swap=None, swap_every=SWAP_EVERY):
    """K steps, one per call, the L2 flushed before every step and every step timed by its own CUDA events; returns
    (total ms = max over ranks, per-step ms array of this rank, wall seconds, swap rounds run)."""
    ev = [(ctx.event(), ctx.event()) for _ in range(steps)]
    ctx.sync_all()
    wall0 = time.perf_counter()
    n_swaps = 0
    for i in range(steps):
        ctx.flush.fill_(i & 0xFF)
        ev[i][0].record()
        eng.step(1)
        if swap is not None and ((i + 1) % min(swap_every, steps) == 0):
            swap(n_swaps)
            n_swaps += 1
        ev[i][1].record()
    ctx.sync_all()
    wall = time.perf_counter() - wall0
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    return ctx.max_over_ranks(float(step_ms.sum())), step_ms, wall, n_swaps


def timed_call(ctx, eng, steps, swap=None, swap_every=SWAP_EVERY):
    """The same number of steps back to back (one multi-iteration call per swap window, nothing flushed)."""
    e0, e1 = ctx.event(), ctx.event()
    ctx.sync_all()
    e0.record()
    done, r = 0, 10 ** 5
    while done < steps:
        n = min(swap_every, steps - done) if swap is not None else steps - done
        eng.step(n)
        done += n
        if swap is not None and n == swap_every:
            swap(r)
            r += 1
    e1.record()
    ctx.sync_all()
    return ctx.max_over_ranks(e0.elapsed_time(e1))


def set_ladder(ctx, eng, C, on):
    """16-temperature ladder of configs[3], interleaved so that every rank holds all 16 temperatures (hot chains
    accept more births and carry more cells: the ranks stay load balanced) -- or all chains back at T = 1."""
    import torch

    if on:
        ladder = np.tile(np.geomspace(1.0, 5.0, 16), ctx.world * C // 16)
        eng.temperature.copy_(torch.as_tensor(ladder[ctx.rank * C:(ctx.rank + 1) * C]))
    else:
        eng.temperature.fill_(1.0)


def cuda_arm(args):
    import bayesbay_b200 as bb
    from bayesbay_b200 import _compile, synthetic
    from bayesbay_b200._engine import Engine

    ctx = Ctx(args)
    torch, dist, world, rank, device = ctx.torch, ctx.dist, ctx.world, ctx.rank, ctx.device

    ing = build_problem()
    # the problem is described through the reference-facing API and compiled for the device
    par, ll = synthetic.build_problem(ing, bb)
    spec = _compile.compile_problem(par, ll)
    C = CHAINS_PER_GPU
    eng = Engine(spec, C, device=device, seed=12345, chain_id0=rank * C)
    eng.init_states()
    swap = SwapRound(ctx, eng, C)
    tempered_main = world > 1

    # burn the chains in a little so k and sigma are in a representative regime, then W warm-up steps
    set_ladder(ctx, eng, C, tempered_main)
    eng.step(max(args.burn, 0))
    for _ in range(args.warmup):
        eng.step(1)
    swap(10 ** 6)  # warm the all-gather / swap path (untimed)
    if not tempered_main:
        set_ladder(ctx, eng, C, False)
    ctx.sync_all()

    # ---- timed region: exactly K steps; L2 flushed between steps, each step timed by its own events ---------
    clocks = ClockSampler(ctx.local_rank)
    if rank == 0:
        clocks.start()
    total_ms, step_ms, wall, n_swaps = timed_steps(ctx, eng, args.steps, swap if tempered_main else None)
    clk = clocks.stop() if rank == 0 else None
    value = world * C * args.steps / (total_ms * 1e-3)

    # ---- the same K steps back to back without the flush (L2-warm steady state) ----------------------------
    # (every back-to-back protocol is rehearsed once untimed: graph captures, lazily loaded kernels and the first use
    # of a collective otherwise land inside a 2 ms timed region)
    timed_call(ctx, eng, args.steps, swap if tempered_main else None)
    warm_ms = timed_call(ctx, eng, args.steps, swap if tempered_main else None)

    # ---- the other per-GPU workload (see the module docstring), same protocol ------------------------------
    set_ladder(ctx, eng, C, not tempered_main)
    eng.step(40)  # let the chains settle under the changed temperatures (untimed)
    o_ms, o_step, _, o_swaps = timed_steps(ctx, eng, args.steps, None if tempered_main else swap)
    timed_call(ctx, eng, args.steps, None if tempered_main else swap)
    o_warm = timed_call(ctx, eng, args.steps, None if tempered_main else swap)
    other = {"value": world * C * args.steps / (o_ms * 1e-3), "ms_per_step": o_ms / args.steps,
             "value_l2_warm": world * C * args.steps / (o_warm * 1e-3), "swap_rounds_timed": o_swaps,
             "what": ("every rank's chains at T = 1, no exchange (configs[2] per GPU)" if tempered_main else
                      "the same chains under the 16-temperature ladder of configs[3], one swap round per %d steps "
                      "(the per-GPU workload of the N > 1 lines)" % min(SWAP_EVERY, args.steps))}
    set_ladder(ctx, eng, C, tempered_main)
    eng.step(40)

    # ---- a long run in one call: >= 1000 iterations incl. the from-scratch re-evaluations every 256 steps ---
    long_n = 1024
    timed_call(ctx, eng, 512, swap if tempered_main else None)  # (rehearsal: crosses two from-scratch re-evaluations)
    long_ms = timed_call(ctx, eng, long_n, swap if tempered_main else None)

    # ---- per-kernel breakdown + roofline of the dominant kernel (rank 0) -----------------------------------
    # stages of bbb_engine_stage, full-SpMM step: 0 = k_propose, 1 = k_idx_commit_copy + k_raster2d_inc +
    # k_raster2d_rescan, 2 = k_ray_spmm, 3 = k_finish  (6 kernel launches per step);
    # chain-local step: 0 = k_propose, 2 = k_chain_step, 1 and 3 launch nothing (2 launches per step, plus the
    # from-scratch re-evaluation of residuals and misfit every 256 steps: 5 launches)
    names = ["propose", "raster2d", "chain_step" if eng.chain_local else "ray_spmm", "finish"]
    stage_ms = np.zeros(4)
    active_fwd = active_geom = 0
    resync_ms = None
    if rank == 0:
        sev = [[(ctx.event(), ctx.event()) for _ in range(4)] for _ in range(args.steps)]
        for i in range(args.steps):
            ctx.flush.fill_(i & 0xFF)
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
        if eng.chain_local:  # the periodic from-scratch re-evaluation (full SpMM + transposes), timed on its own
            e0, e1 = ctx.event(), ctx.event()
            ctx.flush.fill_(1)
            e0.record()
            eng.resync()
            e1.record()
            torch.cuda.synchronize(device)
            resync_ms = e0.elapsed_time(e1)
    ctx.sync_all()

    # ---- end to end through the plugin call with HOST buffers ----------------------------------------------
    e2e = e2e_run(eng, args.steps, device, world, dist if world > 1 else None)
    e2e_copy = e2e_copy_engine_run(eng, args.steps, device) if world == 1 else None

    # ---- the other BASELINE configurations (every N; weak scaling: the per-GPU work is fixed) ---------------
    configs = other_configs(ctx, args)

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ab = algorithmic_bytes(ing, C)
    peaks = load_json("MEASURED_PEAKS.json")
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    dom = int(np.argmax(stage_ms))
    units = {"ray_spmm": active_fwd / args.steps, "chain_step": active_fwd / args.steps,
             "raster2d": active_geom / args.steps}.get(names[dom], C)
    per_unit = {"ray_spmm": ab["forward_likelihood"], "chain_step": ab["chain_step"],
                "raster2d": ab["raster"]}.get(names[dom], 48 * 200)
    achieved = per_unit * units / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
    traffic = load_json("profiles", "traffic.json").get(names[dom])
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "chains_per_gpu": C, "n_chains": world * C,
                   "l2": "flushed between timed steps (256 MiB write); each step timed by its own CUDA events",
                   "tempering": ("16-temperature ladder, NCCL all-gather swap round every %d steps" % min(SWAP_EVERY, args.steps))
                   if tempered_main else "none (all chains T=1)",
                   "burn_steps": args.burn,
                   "resync": "chain-local engines re-evaluate residuals and misfit from scratch every 256 steps; the %d timed "
                             "steps do not contain one -- `long_run` (1024 steps in one call) contains four, `resync_ms` is "
                             "one on its own" % args.steps},
        "value_l2_warm": world * C * args.steps / (warm_ms * 1e-3),
        "ms_per_step_l2_warm": warm_ms / args.steps,
        "long_run": {"value": world * C * long_n / (long_ms * 1e-3), "ms_per_step": long_ms / long_n, "steps": long_n,
                     "what": "1024 iterations in one call per swap window, back to back, incl. 4 from-scratch re-evaluations"},
        "resync_ms": resync_ms,
        "tempered" if not tempered_main else "replicas": other,
        "wall_s_timed_region": wall,
        "step_ms_min_median_max": [float(step_ms.min()), float(np.median(step_ms)), float(step_ms.max())],
        "clocks": {k: clk[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")} if clk else None,
        "e2e": e2e,
        **({"e2e_copy_engine": e2e_copy} if e2e_copy is not None else {}),
        "gpu_launches": (2 if eng.chain_local else 6) * args.steps + (3 * n_swaps if tempered_main else 0),
        "step_kind": "chain-local (k_propose, k_chain_step)" if eng.chain_local
        else "full SpMM (k_propose, k_idx_commit_copy, k_raster2d_inc, k_raster2d_rescan, k_ray_spmm, k_finish)",
        "kernel_ms": {n: float(v) for n, v in zip(names, stage_ms) if not (eng.chain_local and n in ("raster2d", "finish"))},
        "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_chain_step": per_unit, "chain_steps_per_launch": units,
                     "whole_step_frac": ab["total"] * C / (warm_ms / args.steps * 1e-3) / 1e9 / peak,
                     "alt": alt_rooflines(names[dom], stage_ms[dom])},
        "configs": configs,
    }
    if world == 1:
        out["api_run"] = api_run(par, ll, bb, C)
        plug = plugin_run(ing, C)
        if plug is not None:
            out["plugin_run"] = plug
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(ing, spec, args.cpu_seconds)
    emit(out)
    if world > 1:
        dist.destroy_process_group()


def alt_rooflines(kernel, live_ms):
    """The dominant kernel against the bounds that are NOT HBM bandwidth: per-launch counters of one ncu --set full
    capture (profiles/r2_counters.json) over the launch time measured live here, against the peaks measured on this
    pool's B200 with tools/peaks.cu (profiles/r2_peaks.json).  The kernel's real DRAM traffic is an eighth of its
    algorithmic bytes, so the HBM figure above says how much work is avoided, not what binds the kernel."""
    peaks = load_json("profiles", "r2_peaks.json")
    cnt = load_json("profiles", "r2_counters.json").get(kernel)
    if not peaks or not cnt or live_ms <= 0:
        return None
    s = live_ms * 1e-3
    rows = []

    def add(bound, achieved, peak, unit):
        rows.append({"bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak})

    if "dram_bytes" in cnt:
        add("hbm (bytes moved, ncu)", cnt["dram_bytes"] / s / 1e9, float(load_json("MEASURED_PEAKS.json").get("hbm_gbs", 6650.0)), "GB/s")
    if "l2_bytes" in cnt:
        add("l2", cnt["l2_bytes"] / s / 1e9, peaks["l2_read_gbs"], "GB/s")
    if cnt.get("fp64_pipe_pct_of_peak_active") and cnt.get("ncu_duration_us"):
        # ncu's fp64-pipe utilisation of the captured launch, rescaled to the live launch time
        frac = cnt["fp64_pipe_pct_of_peak_active"] / 100.0 * (cnt["ncu_duration_us"] * 1e-6) / s
        add("fp64 pipe", frac * peaks["fp64_fma_tflops"], peaks["fp64_fma_tflops"], "TFLOP/s")
    if "smem_atomic_lane_ops" in cnt:
        add("shared-memory atomics", cnt["smem_atomic_lane_ops"] / s, peaks["smem_atoms_lane_ops_per_s"], "lane-ops/s")
    if "smem_bytes" in cnt:
        add("shared-memory loads", cnt["smem_bytes"] / s / 1e9, peaks["smem_load_gbs"], "GB/s")
    if "warp_instructions" in cnt:
        # 4 schedulers per SM, one warp instruction per scheduler and clock
        add("issue slots", cnt["warp_instructions"] / s, 4 * peaks["n_sm"] * peaks["sm_clock_mhz_nominal"] * 1e6, "warp-instr/s")
    gov = max(rows, key=lambda r: r["frac"]) if rows else None
    return {"source": cnt.get("source"), "bounds": rows, "governing": gov["bound"] if gov else None,
            "governing_frac": gov["frac"] if gov else None, "note": cnt.get("note")}


def other_configs(ctx, args):
    """configs[4] (C5), configs[0] (C1), configs[1] (C2) at this N: per-GPU work fixed (weak scaling), chains of every
    rank independent.  Multi-iteration calls, CUDA events, max over ranks."""
    from bayesbay_b200 import _compile, synthetic
    from bayesbay_b200._engine import Engine
    import bayesbay_b200 as bb

    torch, world, rank, device = ctx.torch, ctx.world, ctx.rank, ctx.device
    out = {}

    def engine_for(ing, chains, seed):
        par, ll = synthetic.build_problem(ing, bb)
        spec = _compile.compile_problem(par, ll)
        eng = Engine(spec, chains, device=device, seed=seed, chain_id0=rank * chains)
        eng.init_states()
        return eng

    def rate(eng, chains, steps, reps=2):
        best = None
        for _ in range(reps):
            ms = timed_call(ctx, eng, steps)
            best = ms if best is None or ms < best else best
        return world * chains * steps / (best * 1e-3), best / steps

    if not args.no_c5:
        t0 = time.perf_counter()
        ing = synthetic.tomography_2d(nx=256, ny=256, n_sources_per_side=100, n_receivers=250, kmin=50, kmax=500)
        eng = engine_for(ing, CHAINS_PER_GPU, 4321)
        eng.step(args.burn_c5)
        v, ms = rate(eng, CHAINS_PER_GPU, args.steps)
        ev = [(ctx.event(), ctx.event()) for _ in range(args.steps)]
        ctx.sync_all()
        for i in range(args.steps):  # one iteration per call, L2 flushed (the protocol of `value`)
            ctx.flush.fill_(i & 0xFF)
            ev[i][0].record()
            eng.step(1)
            ev[i][1].record()
        ctx.sync_all()
        fl_ms = ctx.max_over_ranks(float(sum(a.elapsed_time(b) for a, b in ev)))
        ab = algorithmic_bytes(ing, CHAINS_PER_GPU)
        peak = float(load_json("MEASURED_PEAKS.json").get("hbm_gbs", 6650.0))
        k_mean = float(eng.gather_space(0)[0].float().mean().item())
        out["c5"] = {"workload": "BASELINE configs[4]: 256x256 grid, 100000 rays (nnz %d), k in [50,500], hierarchical noise, "
                                 "%d chains per GPU, chain-local step over %d ray tiles" % (ing["indices"].size, CHAINS_PER_GPU, eng.ray_tiles(0)),
                     "value": world * CHAINS_PER_GPU * args.steps / (fl_ms * 1e-3), "ms_per_step": fl_ms / args.steps,
                     "value_l2_warm": v, "ms_per_step_l2_warm": ms, "unit": UNIT, "k_mean": k_mean,
                     "roofline": {"bound": "hbm", "achieved": ab["total"] * v / world / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": ab["total"] * v / world / 1e9 / peak, "algorithmic_bytes_per_chain_step": ab["total"],
                                  "traffic": load_json("profiles", "traffic.json").get("chain_step_c5")},
                     "setup_s": time.perf_counter() - t0}
        del eng
        torch.cuda.empty_cache()
    for key, ing, chains, steps, what in (
            ("c1", synthetic.partition_1d(), 4, 20000, "BASELINE configs[0]: Voronoi1D piecewise-constant regression, 100 points, 4 chains per GPU"),
            ("c2", synthetic.polyfit(), 4096, 20000, "BASELINE configs[1]: polynomial regression, 4 Gaussian priors, hierarchical noise, 4096 chains per GPU")):
        eng = engine_for(ing, chains, 99)
        eng.step(500)
        v, ms = rate(eng, chains, steps)
        out[key] = {"workload": what, "value": v, "ms_per_step": ms, "unit": UNIT,
                    "what": "k_fused_steps: %d iterations of every chain in one launch (proposal + forward + decision per thread)" % steps}
        del eng
    return out


def api_run(par, ll, bb, n_chains, n_iterations=2000, save_every=100):
    """Wall clock of the call a user makes -- BayesianInversion.run with the reference's defaults for what is saved
    (every chain's state AND d_pred at every save point come back to host lists, as in the reference) -- for
    information; not the device-timed `value`."""
    import torch

    inv = bb.BayesianInversion(par, ll, n_chains=n_chains)
    inv.run(n_iterations=200, burnin_iterations=100, save_every=save_every, verbose=False)  # warm-up
    n_before = len(inv.get_results()["voronoi.n_dimensions"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    inv.run(n_iterations=n_iterations, burnin_iterations=0, save_every=save_every, verbose=False)
    torch.cuda.synchronize()
    wall_run = time.perf_counter() - t0
    res = inv.get_results()
    wall = time.perf_counter() - t0
    n_samples = len(res["voronoi.n_dimensions"]) - n_before
    return {"value": n_chains * n_iterations / wall, "unit": UNIT, "wall_s": wall, "wall_s_run_only": wall_run,
            "what": f"BayesianInversion.run(n_iterations={n_iterations}, save_every={save_every}) + get_results() on the same "
                    f"problem and chain count, host wall clock: the saved samples ({n_samples} states with d_pred, the "
                    "reference's defaults) are in host memory when run() returns and in the reference's result lists at the end"}


def plugin_run(ing, n_chains, n_iterations=2000, save_every=100):
    """The drop-in seam itself: the UNMODIFIED reference's `BayesianInversion.run(sampler=...)` (baseline/_ref, when it is
    installed) with the plug-in sampler of bayesbay_b200/plugin.py, which runs the reference's own chain objects on the CUDA
    engine.  Host wall clock of the second run (the first one also creates the engine); for information."""
    import torch

    try:
        import importlib

        from bayesbay_b200 import synthetic
        from bayesbay_b200.plugin import cuda_samplers

        ref_dir = os.path.join(REPO, "baseline", "_ref")
        if not os.path.isdir(os.path.join(ref_dir, "bayesbay")):
            return None
        for mod in ("matplotlib", "shapely"):  # hard imports of the reference used for plotting / polygons only
            try:
                importlib.import_module(mod)
            except ImportError:
                stubs = os.path.join(ref_dir, "_stubs")
                if stubs not in sys.path:
                    sys.path.append(stubs)
        if ref_dir not in sys.path:
            sys.path.insert(0, ref_dir)
        bayesbay = importlib.import_module("bayesbay")
        par, ll = synthetic.build_problem(ing, bayesbay, synthetic.forward_operator)
        inv = bayesbay.BayesianInversion(par, ll, n_chains=n_chains)
        CudaVanillaSampler, _ = cuda_samplers(bayesbay, seed=3)
        walls = []
        for _ in range(2):
            t0 = time.perf_counter()
            inv.run(sampler=CudaVanillaSampler(), n_iterations=n_iterations, burnin_iterations=0, save_every=save_every, verbose=False)
            torch.cuda.synchronize()
            res = inv.get_results()
            walls.append(time.perf_counter() - t0)
        return {"value": n_chains * n_iterations / walls[1], "unit": UNIT, "wall_s": walls[1], "wall_s_first_run": walls[0],
                "n_samples": len(res["voronoi.n_dimensions"]),
                "what": f"bayesbay (unmodified, baseline/_ref) BayesianInversion.run(sampler=CudaVanillaSampler(), n_iterations="
                        f"{n_iterations}, save_every={save_every}) + get_results() on {n_chains} chains the reference initialised itself: "
                        "host wall clock of the second run (the first also builds the engine)"}
    except Exception as e:  # pragma: no cover  (informational leg: never fails the bench)
        log("plugin_run skipped:", repr(e))
        return None


def e2e_run(eng, steps, device, world, dist):
    """The plugin call with HOST buffers, one iteration per call (bbb_engine_step_hosted_mapped + _wait): the chains'
    states (k, sites, values, noise sigma, temperature, iteration) live in pinned host memory as fixed-stride records;
    every chain's block of the chain kernel reads its record straight from host memory (host -> device over the bus),
    compares it with the resident state its caches belong to, advances the chain one iteration and writes the new
    record back to host memory (device -> host); the host waits for the step before the next call.  The derived
    per-chain caches (nearest-site index, residuals, misfit) stay resident on the device between calls and are only
    used for a chain whose uploaded record is bit-identical to the state they belong to (else the chain is left
    untouched and the wait re-derives the caches with bbb_engine_refresh and advances it) -- the analogue of the
    d_pred / KD-tree caches the reference pickles along with every chain.  Bytes per step are counted by the kernel
    (header + the k cells in use of every chain, each way)."""
    import torch

    rw = eng.hosted_record_words()
    host = [torch.zeros((eng.C, rw), dtype=torch.int64).pin_memory() for _ in range(2)]  # the caller's chains, in / out
    refreshed = [0]

    eng.pack_states_records(host[0])
    torch.cuda.synchronize(device)
    cur = 0
    kb = None

    def call(_, cur):
        eng.step_hosted_mapped(host[cur], host[cur ^ 1])
        refreshed[0] += eng.step_hosted_mapped_wait()  # the host has the step's result before the next call
        return None

    for _ in range(3):
        kb = call(kb, cur)
        cur ^= 1
    refreshed[0] = 0
    moved0 = eng.hosted_mapped_bytes()
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
    moved = [b - a for a, b in zip(moved0, eng.hosted_mapped_bytes())]
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"value": world * eng.C * steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": moved[0] // steps,
            "d2h_bytes_per_step": moved[1] // steps, "ms_per_step": ms / steps, "host_wall_ms_per_step": wall_ms / steps,
            "cache_refreshes": refreshed[0],
            "k_mean": float(eng.gather_space(0)[0].double().mean().item()),
            "what": "per call (bbb_engine_step_hosted_mapped + _wait): the chains' records (k; sites and values of the k cells "
                    "in use; sigma, temperature, iteration) in pinned host memory are read by the chain kernel over the bus, "
                    "compared with the resident states the caches belong to, advanced one iteration and written back to "
                    "pinned host memory by the same kernel; the host waits for the step's result before the next call"}


def e2e_copy_engine_run(eng, steps, device):
    """The same round trip with copy-engine transfers (bbb_engine_step_hosted_ragged): ragged packed image pinned host ->
    staging buffer, compared, proposal + chain kernel, packed, staging buffer -> pinned host, in parts on internal
    streams.  For information (`e2e_copy_engine`)."""
    import torch

    cap = eng.hosted_ragged_words()
    host = [torch.zeros((cap,), dtype=torch.int64, pin_memory=True) for _ in range(2)]
    stage = torch.zeros((cap,), dtype=torch.int64, device=device)
    out = torch.zeros((cap,), dtype=torch.int64, device=device)
    host[0].copy_(eng.pack_states_ragged())
    torch.cuda.synchronize(device)
    moved, redone, cur = [0, 0], 0, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(steps + 3):
        if it == 3:
            moved = [0, 0]
            e0.record()
        up, down = eng.step_hosted_ragged(host[cur], host[cur ^ 1], stage, out)
        redone += eng.step_hosted_wait()
        moved[0] += up
        moved[1] += down
        cur ^= 1
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1)
    return {"value": eng.C * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "h2d_bytes_per_step": moved[0] // steps,
            "d2h_bytes_per_step": moved[1] // steps, "cache_refreshes": redone, "parts": len(eng.hosted_parts()) - 1,
            "what": "bbb_engine_step_hosted_ragged + _wait: copy-engine transfers of a ragged packed image through staging "
                    "buffers, compare / proposal / chain / pack kernels per part on internal streams, replayed from CUDA graphs"}


# --------------------------------------------------------------------------- #
# CPU baseline / reference arm
# --------------------------------------------------------------------------- #
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def reference_sample(ing, cores, seconds, repeats):
    """The unmodified reference's chain pool (oracle/reference_runner.py: baseline/_ref, installed by
    __graft_entry__.build()); None when it cannot be imported here."""
    try:
        from oracle import reference_runner as RR

        return RR.timed_pool(ing, cores, seconds, repeats=repeats)
    except Exception as e:  # ImportError, a build made for another CPU, ...
        log("reference arm: the unmodified reference is not usable here (%s: %s); falling back to the oracle port" % (type(e).__name__, e))
        return None


def cpu_baseline(ing, spec, seconds):
    cores = host_cores()
    ref = reference_sample(ing, cores, seconds, 1)
    if ref is not None:
        return {"value": ref["rates"][0], "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": f"the unmodified reference (bayesbay 0.3.11, baseline/_ref): BayesianInversion.run on {cores} chains "
                          f"with parallel_config n_jobs={cores} (its joblib/loky chain pool, one worker per chain and core) x "
                          f"{ref['n_iterations']} iterations of the same C3 problem, {ref['walls'][0]:.1f} s wall, workers "
                          "already started; the forward is the same operator object evaluated on the host (KD-tree + scipy CSR)"}
    from oracle import cpu_baseline as cb

    n = cb.steps_for_budget(spec, seconds)
    rate, procs, wall = cb.run_pool(spec, n, cores)
    return {"value": rate, "unit": UNIT, "cores": procs, "kind": "port",
            "sample": f"{procs} chains (one process per host core) x {n} iterations of the same C3 problem, "
                      f"{wall:.1f} s wall; oracle port of the reference chain step (KD-tree + scipy CSR forward) -- the "
                      "unmodified reference could not be imported here"}


def reference_arm(args):
    """Times the reference's own CPU implementation of the path -- the unmodified package's joblib/loky chain pool --
    on all host cores, on the CUDA arm's config, metric and unit.  Under torchrun only rank 0 works."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    ing = build_problem()
    cores = host_cores()
    per_step = max(2.0, min(20.0, args.cpu_seconds))
    n_samples = max(1, min(args.steps, 3))
    ref = reference_sample(ing, cores, per_step, n_samples)
    if ref is not None:
        rates, walls, n, kind = ref["rates"], ref["walls"], ref["n_iterations"], "reference"
        sample = (f"each step = BayesianInversion.run of the unmodified reference (baseline/_ref) on {cores} chains, "
                  f"parallel_config n_jobs={cores} (joblib/loky pool, one worker per chain and host core), {n} iterations of "
                  f"the C3 problem; {len(rates)} such samples timed after a warm-up call that starts the workers")
    else:
        from oracle import cpu_baseline as cb
        from tests import problems

        spec = problems.spec_from_ingredients(ing)
        n = cb.steps_for_budget(spec, per_step)
        cb.run_pool(spec, max(5, n // 10), cores)
        rates, walls = [], []
        for _ in range(n_samples):
            t0 = time.perf_counter()
            rate, procs, wall = cb.run_pool(spec, n, cores)
            rates.append(rate)
            walls.append(time.perf_counter() - t0)
        kind = "port"
        sample = (f"each step = {cores} chains (one process per host core) x {n} iterations of the C3 problem, oracle port "
                  f"(the unmodified reference could not be imported here); {len(rates)} such samples")
    value = float(np.mean(rates))
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(rates),
        "warmup": 1, "ms_per_step": float(np.mean(walls)) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "chains": cores, "iterations_per_sample": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
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
    os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")  # (no kernel is loaded for the first time inside a timed region)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--burn", type=int, default=300, help="untimed iterations before the warm-up (state in a typical regime)")
    ap.add_argument("--burn-c5", type=int, default=150, help="untimed iterations of the C5 configuration")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline budget per process")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the configs[4] line (its set-up takes about a minute)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        cuda_arm(args)


if __name__ == "__main__":
    main()
