"""Host-side runtime of a batch of chains: owns one :class:`Engine`, converts between the
reference's per-chain objects (``State``, ``saved_states``, ``statistics``) and the device SoA
buffers, and implements the batched ``advance_chain``.

This replaces the reference's chain pool -- ``Sampler.advance_chain``'s
``joblib.Parallel(n_jobs=len(chains), backend="loky")`` over pickled ``MarkovChain`` objects
(src/bayesbay/samplers/_samplers.py:191-239) -- by launches on one CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import itertools
import os
from collections import defaultdict
from typing import List, Optional

import numpy as np

from . import _compile
from ._state import DataNoiseState, ParameterSpaceState, State

_seed_counter = itertools.count()
_base_seed: Optional[int] = None


def seed(value: Optional[int]):
    """Set the base seed of the Philox streams used by engines created afterwards (``None`` = draw
    from OS entropy, the default -- the reference's chains are unseeded too)."""
    global _base_seed, _seed_counter
    _base_seed = value
    _seed_counter = itertools.count()


def next_seed() -> int:
    n = next(_seed_counter)
    if _base_seed is None:
        return int.from_bytes(os.urandom(8), "little")
    return (int(_base_seed) * 0x9E3779B97F4A7C15 + n * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF


# --------------------------------------------------------------------------- #
# state conversion
# --------------------------------------------------------------------------- #
def state_to_host_dict(spec: dict, state: State) -> dict:
    spaces = []
    for sp in spec["spaces"]:
        ps = state[sp["name"]]
        if ps is None:
            raise ValueError(f"starting state has no parameter space {sp['name']!r}")
        k = int(ps.n_dimensions)
        sites = ps["discretization"] if sp["ndim"] else None
        if sp["ndim"] and sites is None:
            raise ValueError(f"starting state of {sp['name']!r} has no 'discretization'")
        values = []
        for p in sp["params"]:
            v = ps[p["name"]]
            if v is None or len(v) != k:
                raise ValueError(f"starting state: parameter {sp['name']}.{p['name']} should have {k} values")
            values.append(np.asarray(v, dtype=np.float64))
        if sp["ndim"] == 1:
            sites = np.asarray(sites, dtype=np.float64).reshape(k)
            if np.any(np.diff(sites) < 0):
                raise ValueError("Voronoi1D starting sites must be sorted")
        elif sp["ndim"] == 2:
            sites = np.asarray(sites, dtype=np.float64).reshape(k, 2)
        spaces.append(dict(k=k, sites=sites, values=values))
    sigma, corr = [], []
    for t in spec["targets"]:
        if t["hierarchical"]:
            noise = state[t["name"]]
            if noise is None:
                raise ValueError(f"starting state has no data-noise entry for target {t['name']!r}")
            sigma.append(float(noise.std))
            if t.get("correlated"):
                if noise.correlation is None:
                    raise ValueError(f"starting state of target {t['name']!r} has no noise correlation")
                corr.append(float(noise.correlation))
            else:
                corr.append(None)
        else:
            sigma.append(None)
            corr.append(None)
    return dict(spaces=spaces, sigma=sigma, corr=corr)


def host_dict_to_state(spec: dict, d: dict, temperature=1.0, state_cls=State, ps_cls=ParameterSpaceState,
                       noise_cls=DataNoiseState) -> State:
    pv = {}
    for sp, ss in zip(spec["spaces"], d["spaces"]):
        vals = {}
        if sp["ndim"]:
            vals["discretization"] = ss["sites"]
        for p, v in zip(sp["params"], ss["values"]):
            vals[p["name"]] = v
        pv[sp["name"]] = ps_cls(int(ss["k"]), vals)
    for i, (t, sg) in enumerate(zip(spec["targets"], d["sigma"])):
        if t["hierarchical"]:
            pv[t["name"]] = noise_cls(std=sg, correlation=d["corr"][i] if d.get("corr") else None)
    return state_cls(pv, temperature=float(temperature))


def save_schedule(iteration: int, n_iterations: int, burnin_iterations: int, save_every: int):
    """Split ``n_iterations`` (starting after ``iteration`` completed ones) into device launches
    ``[(n_steps, save_after), ...]`` that stop exactly where the reference saves: iteration ``i`` (1-based,
    counted over the chain's whole life) is saved iff ``i > burnin_iterations`` and
    ``(i - burnin_iterations) % save_every == 0`` (src/bayesbay/_markov_chain.py:284-289)."""
    if save_every < 1:
        raise ValueError("save_every must be at least 1")
    out = []
    done = 0
    while done < n_iterations:
        i_next = iteration + done + 1
        if i_next <= burnin_iterations:
            to_save = burnin_iterations + save_every - i_next
        else:
            r = (i_next - burnin_iterations) % save_every
            to_save = 0 if r == 0 else save_every - r
        n = min(n_iterations - done, to_save + 1)
        out.append((n, n == to_save + 1))
        done += n
    return out


# --------------------------------------------------------------------------- #
# the batch
# --------------------------------------------------------------------------- #
class ChainBatch:
    """All chains of one inversion resident on one GPU."""

    def __init__(self, spec: dict, n_chains: int, device=None, seed=None, chain_id0: int = 0,
                 starting_states: Optional[List[State]] = None, save_dpred: bool = True):
        import torch

        from ._engine import Engine

        if device is None:
            device = f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"
        self.spec = spec
        self.n_chains = int(n_chains)
        self.chain_id0 = int(chain_id0)
        self.save_dpred = save_dpred
        self.seed = next_seed() if seed is None else int(seed)
        self.engine = Engine(spec, self.n_chains, device=device, seed=self.seed, chain_id0=chain_id0)
        self.names = _compile.perturbation_names(spec)
        self.iteration = 0
        self.swap_round = 0
        self.resumed = False  # set by BayesianInversion.load_checkpoint: the next run keeps the restored temperatures
        self.saved: List[dict] = []  # one entry per save point: host arrays + mask of chains saved
        self._torch = torch
        # posterior field statistics of ray targets that ask for them (per chain: sum, sum of squares, count)
        self.field_stats = {}
        for t, tg in enumerate(spec["targets"]):
            f = tg["forward"]
            if f["kind"] == "ray2d" and f.get("save_field_as"):
                G = int(np.asarray(f["points"]).shape[0])
                z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=self.engine.device)  # noqa: E731
                self.field_stats[t] = dict(name=f["save_field_as"], store=bool(f.get("store_field_samples", True)),
                                           sum=z(G, self.n_chains), sumsq=z(G, self.n_chains), count=z(self.n_chains))
        if starting_states is None:
            self.engine.init_states()
        else:
            self.engine.load_states([state_to_host_dict(spec, s) for s in starting_states])

    # ---- temperatures ----------------------------------------------------
    def set_temperatures(self, temps):
        t = self._torch.as_tensor(np.asarray(temps, dtype=np.float64))
        self.engine.temperature.copy_(t)

    def temperatures(self) -> np.ndarray:
        return self.engine.temperature.cpu().numpy()

    # ---- checkpoint / resume ---------------------------------------------
    def state_dict(self) -> dict:
        fields = {t: {k: (v.cpu() if hasattr(v, "cpu") else v) for k, v in fs.items()} for t, fs in self.field_stats.items()}
        self.flush_saves()
        # (a copy of the list: a snapshot kept in memory must not grow with the run; the entries' arrays are never mutated)
        saved = [{k: v for k, v in e.items() if k != "views"} for e in self.saved]
        return dict(engine=self.engine.state_dict(), iteration=self.iteration, saved=saved, field_stats=fields,
                    swap_round=self.swap_round, seed_global=getattr(self, "seed_global", self.seed))

    def load_state_dict(self, state: dict):
        self.engine.load_state_dict(state["engine"])
        self.iteration = int(state["iteration"])
        self.swap_round = int(state.get("swap_round", 0))
        self.saved = list(state["saved"])
        self.seed_global = state["seed_global"]
        for t, fs in state.get("field_stats", {}).items():
            for key in ("sum", "sumsq", "count"):
                self.field_stats[t][key].copy_(fs[key])

    # ---- advancing -------------------------------------------------------
    def advance(self, n_iterations: int, burnin_iterations: int = 0, save_every: int = 100):
        """``BaseMarkovChain.advance_chain`` for all chains (src/bayesbay/_markov_chain.py:249-297):
        iteration ``i`` (1-based, global) is saved iff ``i > burnin`` and ``(i - burnin) % save_every == 0``
        and the chain's temperature is 1 (``_next_iteration`` :241-242)."""
        for n, save in save_schedule(self.iteration, n_iterations, burnin_iterations, save_every):
            self.engine.step(n)
            self.iteration += n
            if save:
                self._save()

    # ---- save points: device image -> pinned ring -> host, off the stepping stream -----------------------------
    N_SLOTS = 4  # images in flight (device buffer + pinned host buffer each), one worker thread per slot

    def _save_setup(self):
        """Layout of a save-point image (bbb_engine_save_point, include/bbb200.h) and the transfer machinery: a ring of
        device images written by the save kernels on the stepping stream, copied to a ring of pinned host buffers on a
        copy stream, and handed to the sample list by a worker thread -- the chains keep stepping meanwhile."""
        import queue
        import threading

        t, eng, C = self._torch, self.engine, self.n_chains
        self._kb_save = [sp["kmax"] for sp in self.spec["spaces"]]  # (all cell rows travel: no sync to learn the largest k)
        off = 0
        lay = dict(spaces=[], noise=[], dpred=[])
        for sp, kb in zip(self.spec["spaces"], self._kb_save):
            nd, P = sp["ndim"], len(sp["params"])
            lay["spaces"].append(dict(k=off, sites=off + C, values=off + C + C * kb * nd, kb=kb, nd=nd, P=P))
            off += C * (1 + kb * (nd + P))
        for tg in self.spec["targets"]:
            ent = dict(sigma=None, corr=None)
            if tg["hierarchical"]:
                ent["sigma"] = off
                off += C
                if tg.get("correlated"):
                    ent["corr"] = off
                    off += C
            lay["noise"].append(ent)
        lay["temperature"] = off
        off += C
        for tg in self.spec["targets"]:
            R = int(np.asarray(tg["dobs"]).size)
            lay["dpred"].append((off, R) if self.save_dpred else None)
            if self.save_dpred:
                off += C * R
        lay["words"] = off
        if 8 * off != eng.save_point_bytes(self._kb_save, self.save_dpred):
            raise RuntimeError("save-point layout mismatch between the host and libbbb200.so")
        self._lay = lay
        dev = eng.device
        self._save_dev = [t.empty((off,), dtype=t.int64, device=dev) for _ in range(self.N_SLOTS)]
        self._save_pin = [t.empty((off,), dtype=t.int64, pin_memory=True) for _ in range(self.N_SLOTS)]
        self._save_packed = [t.cuda.Event() for _ in range(self.N_SLOTS)]
        self._save_copied = [t.cuda.Event() for _ in range(self.N_SLOTS)]
        self._save_used = [False] * self.N_SLOTS
        self._copy_stream = t.cuda.Stream(device=dev)
        needs_scratch = self.save_dpred and any(not (eng.chain_local and eng.res[i] is not None)
                                                for i in range(len(self.spec["targets"])))
        r_max = max((int(np.asarray(tg["dobs"]).size) for tg in self.spec["targets"]), default=0)
        self._save_scratch = t.empty((r_max, C), dtype=t.float64, device=dev) if needs_scratch and r_max else None
        self._free = queue.Queue()
        for i in range(self.N_SLOTS):
            self._free.put(i)
        self._todo = queue.Queue()
        self._save_errors = []
        # (the worker holds the queues, events and pinned buffers, not the batch: the batch stays collectable)
        # (several workers: copying an image out of the pinned ring first-touches ~90 MB of fresh host memory, which one
        # thread does at 3-4 GB/s -- slower than the chains produce save points)
        self._workers = [threading.Thread(target=self._save_worker, daemon=True,
                                          args=(self._todo, self._free, self._save_copied, self._save_pin, self._save_errors))
                         for _ in range(self.N_SLOTS)]
        for w in self._workers:
            w.start()

    @staticmethod
    def _save_worker(todo, free, copied, pinned, errors):
        while True:
            item = todo.get()
            if item is None:
                todo.task_done()
                return
            slot, entry = item
            try:
                copied[slot].synchronize()
                entry["image"] = pinned[slot].numpy().copy()  # (the GIL is released for the copy)
            except Exception as exc:  # pragma: no cover
                errors.append(exc)
            finally:
                free.put(slot)
                todo.task_done()

    def __del__(self):
        todo = getattr(self, "_todo", None)
        if todo is not None:
            for _ in getattr(self, "_workers", [None]):
                todo.put(None)

    def flush_saves(self):
        """Wait until every save point submitted so far has arrived in ``self.saved``."""
        if getattr(self, "_todo", None) is not None:
            self._todo.join()
            if self._save_errors:
                raise self._save_errors[0]

    def _save(self):
        """One save point: the save kernels write the chain-major image into a free device slot (stepping stream), the
        copy stream brings it to pinned host memory, the worker thread appends it to the sample list.  Nothing here
        waits for the device unless all slots are in flight."""
        t, eng = self._torch, self.engine
        if getattr(self, "_lay", None) is None:
            self._save_setup()
        entry = dict(image=None, layout=self._lay, fields={}, chain_major=True)
        for ti, fs in self.field_stats.items():
            # rasterised field of the current states: cell value looked up through the resident nearest-site index
            f = self.spec["targets"][ti]["forward"]
            _, _, vals = eng.gather_space(f["space"])
            field = t.gather(vals[f["param"]], 0, eng.current_cell_idx(ti).long() & 0xFFFF)  # [G][C]
            m = (eng.temperature == 1.0).to(field.dtype)
            fs["sum"] += field * m
            fs["sumsq"] += field * field * m
            fs["count"] += m
            if fs["store"]:
                entry["fields"][fs["name"]] = field.t().contiguous().cpu().numpy()  # [C][G]
        slot = self._free.get()  # (blocks only when N_SLOTS images are in flight)
        main = t.cuda.current_stream(eng.device)
        if self._save_used[slot]:
            main.wait_event(self._save_copied[slot])  # the slot's previous image has left the device buffer
        eng.save_point(self._kb_save, self.save_dpred, self._save_dev[slot], self._save_scratch)
        self._save_packed[slot].record(main)
        with t.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._save_packed[slot])
            self._save_pin[slot].copy_(self._save_dev[slot], non_blocking=True)
            self._save_copied[slot].record(self._copy_stream)
        self._save_used[slot] = True
        self.saved.append(entry)
        self._todo.put((slot, entry))

    @staticmethod
    def _entry_views(entry, spec, C):
        """Numpy views into a save point's host image: mask of saved chains, per space (k, sites, values), noise, d_pred."""
        if "views" in entry:
            return entry["views"]
        img, lay = entry["image"], entry["layout"]
        f64 = img.view(np.float64)
        v = dict(mask=f64[lay["temperature"]: lay["temperature"] + C] == 1.0, spaces=[], sigma=[], corr=[], dpred=[])
        for sl in lay["spaces"]:
            kb, nd, P = sl["kb"], sl["nd"], sl["P"]
            v["spaces"].append((img[sl["k"]: sl["k"] + C],
                                f64[sl["sites"]: sl["sites"] + C * kb * nd].reshape(C, kb, nd) if nd else None,
                                f64[sl["values"]: sl["values"] + C * P * kb].reshape(C, P, kb) if P else None))
        for nl in lay["noise"]:
            v["sigma"].append(None if nl["sigma"] is None else f64[nl["sigma"]: nl["sigma"] + C])
            v["corr"].append(None if nl["corr"] is None else f64[nl["corr"]: nl["corr"] + C])
        for dl in lay["dpred"]:
            v["dpred"].append(None if dl is None else f64[dl[0]: dl[0] + C * dl[1]].reshape(C, dl[1]))
        entry["views"] = v
        return v

    # ---- results in the reference's shapes -------------------------------
    def saved_states_of(self, c: int, start: int = 0):
        """``chain.saved_states`` (src/bayesbay/_markov_chain.py:115-134): ``defaultdict(list)`` keyed
        ``"{ps}.n_dimensions"`` (trans-d only), ``"{ps}.discretization"``, ``"{ps}.{param}"``,
        ``"{target}.std"``, ``"{target}.dpred"``."""
        self.flush_saves()
        out = defaultdict(list)
        C = self.n_chains
        for entry in self.saved[start:]:
            if entry.get("image") is None and "spaces" in entry:
                self._legacy_entry(entry, c, out)  # (checkpoints written before the image layout)
                continue
            v = self._entry_views(entry, self.spec, C)
            if not v["mask"][c]:
                continue
            for sp, (k, sites, vals) in zip(self.spec["spaces"], v["spaces"]):
                kc = int(k[c])
                if sp["trans_d"]:
                    out[f"{sp['name']}.n_dimensions"].append(kc)
                if sp["ndim"] == 1:
                    out[f"{sp['name']}.discretization"].append(sites[c, :kc, 0].copy())
                elif sp["ndim"] == 2:
                    out[f"{sp['name']}.discretization"].append(sites[c, :kc].copy())
                for j, p in enumerate(sp["params"]):
                    out[f"{sp['name']}.{p['name']}"].append(vals[c, j, :kc].copy())
            for tg, sg, rr in zip(self.spec["targets"], v["sigma"], v["corr"]):
                if sg is not None:
                    out[f"{tg['name']}.std"].append(float(sg[c]))
                if rr is not None:  # DataNoiseState.todict: "{target}.correlation" (reference _state.py:26-46)
                    out[f"{tg['name']}.correlation"].append(float(rr[c]))
            for tg, dp in zip(self.spec["targets"], v["dpred"]):
                if dp is not None:
                    out[f"{tg['name']}.dpred"].append(dp[c])  # (the chain's row of the image, handed out as a view)
            for name, field in entry.get("fields", {}).items():
                out[name].append(field[c])
        return out

    def results(self, keys=None):
        """All saved samples of all chains of this batch in the order of the reference's ``get_results`` with
        ``concatenate_chains=True`` (chain by chain, ``_bayes_inversion.py:310-318``), built per save point from
        views of the host images instead of per chain."""
        self.flush_saves()
        if any(e.get("image") is None for e in self.saved):
            return None  # (entries of an old checkpoint layout: the per-chain path handles them)
        C, E = self.n_chains, len(self.saved)
        want = (lambda k: True) if keys is None else (lambda k: k in keys)
        cols = defaultdict(list)
        masks = []
        for entry in self.saved:
            v = self._entry_views(entry, self.spec, C)
            masks.append(v["mask"])
            for sp, (k, sites, vals) in zip(self.spec["spaces"], v["spaces"]):
                kl = k.tolist()
                name = sp["name"]
                if sp["trans_d"] and want(f"{name}.n_dimensions"):
                    cols[f"{name}.n_dimensions"].append(kl)
                if sp["ndim"] and want(f"{name}.discretization"):
                    cols[f"{name}.discretization"].append([sites[c, :kc, 0] for c, kc in enumerate(kl)] if sp["ndim"] == 1
                                                          else [sites[c, :kc] for c, kc in enumerate(kl)])
                for j, p in enumerate(sp["params"]):
                    if want(f"{name}.{p['name']}"):
                        cols[f"{name}.{p['name']}"].append([vals[c, j, :kc] for c, kc in enumerate(kl)])
            for tg, sg, rr, dp in zip(self.spec["targets"], v["sigma"], v["corr"], v["dpred"]):
                if sg is not None and want(f"{tg['name']}.std"):
                    cols[f"{tg['name']}.std"].append(sg.tolist())
                if rr is not None and want(f"{tg['name']}.correlation"):
                    cols[f"{tg['name']}.correlation"].append(rr.tolist())
                if dp is not None and want(f"{tg['name']}.dpred"):
                    cols[f"{tg['name']}.dpred"].append(list(dp))
            for name, field in entry.get("fields", {}).items():
                if want(name):
                    cols[name].append(list(field))
        if E == 0:
            return {}
        M = np.array(masks)
        full = bool(M.all())
        out = {}
        for key, per_entry in cols.items():
            if len(per_entry) != E:
                return None  # (a key that is not present at every save point: leave it to the per-chain path)
            if full:
                out[key] = [per_entry[e][c] for c in range(C) for e in range(E)]
            else:
                out[key] = [per_entry[e][c] for c in range(C) for e in range(E) if M[e, c]]
        return out

    def _legacy_entry(self, entry, c, out):
        if not entry["mask"][c]:
            return
        cm = entry.get("chain_major", False)
        for sp, (k, sites, vals) in zip(self.spec["spaces"], entry["spaces"]):
            kc = int(k[c])
            if sp["trans_d"]:
                out[f"{sp['name']}.n_dimensions"].append(kc)
            if sp["ndim"] == 1:
                out[f"{sp['name']}.discretization"].append((sites[c, :kc, 0] if cm else sites[0, :kc, c]).copy())
            elif sp["ndim"] == 2:
                out[f"{sp['name']}.discretization"].append((sites[c, :kc] if cm else sites[:, :kc, c].T).copy())
            for j, p in enumerate(sp["params"]):
                out[f"{sp['name']}.{p['name']}"].append((vals[c, j, :kc] if cm else vals[j, :kc, c]).copy())
        for i, (tg, sg) in enumerate(zip(self.spec["targets"], entry["sigma"])):
            if sg is not None:
                out[f"{tg['name']}.std"].append(float(sg[c]))
            rr = entry.get("corr", [None] * len(entry["sigma"]))[i]
            if rr is not None:
                out[f"{tg['name']}.correlation"].append(float(rr[c]))
        for tg, dp in zip(self.spec["targets"], entry["dpred"]):
            if dp is not None:
                out[f"{tg['name']}.dpred"].append(dp[c] if cm else dp[:, c].copy())
        for name, field in entry.get("fields", {}).items():
            out[name].append(field[c] if cm else field[:, c].copy())

    def statistics(self):
        """Per-chain statistics dictionaries in the reference's shape (``_init_statistics`` /
        ``_save_statistics``, src/bayesbay/_markov_chain.py:106-157)."""
        prop = self.engine.n_proposed.cpu().numpy()
        acc = self.engine.n_accepted.cpu().numpy()
        exc = self.engine.n_exceptions.cpu().numpy()
        out = []
        for c in range(self.n_chains):
            n_prop = defaultdict(lambda: defaultdict(float))
            n_acc = defaultdict(lambda: defaultdict(float))
            for m, nm in enumerate(self.names):
                if nm is None or prop[m, c] == 0:
                    continue
                outer, inner = nm
                if inner is None:
                    n_prop[outer] = int(prop[m, c])
                    n_acc[outer] = int(acc[m, c])
                else:
                    n_prop[outer][inner] += float(prop[m, c])
                    n_acc[outer][inner] += float(acc[m, c])
            out.append({
                "n_proposed_models": n_prop,
                "n_accepted_models": n_acc,
                "n_proposed_models_total": int(prop[:, c].sum()),
                "n_accepted_models_total": int(acc[:, c].sum()),
                # device analogue of the reference's exception counts (_markov_chain.py:205-247): proposals with a
                # non-finite likelihood ratio (rejected), rejection-sampling loops that gave up (old value kept)
                "exceptions": defaultdict(int, {k: int(v) for k, v in (("NonFiniteLikelihoodRatio", exc[0, c]),
                                                                       ("ResamplingLimitReached", exc[1, c])) if v}),
            })
        return out

    def field_statistics(self, name: str) -> dict:
        """Posterior statistics of a saved field accumulated on the device over the saved (T = 1) samples:
        ``mean`` / ``std`` [G] over all samples, ``chain_mean`` [C][G], ``count`` [C]."""
        for fs in self.field_stats.values():
            if fs["name"] == name:
                n = fs["count"].sum().clamp(min=1)
                mean = fs["sum"].sum(1) / n
                var = (fs["sumsq"].sum(1) / n - mean * mean).clamp(min=0)
                chain_mean = (fs["sum"] / fs["count"].clamp(min=1)).T
                return dict(mean=mean.cpu().numpy(), std=var.sqrt().cpu().numpy(), chain_mean=chain_mean.cpu().numpy(),
                            count=fs["count"].cpu().numpy())
        raise KeyError(f"no forward operator saves a field named {name!r}")

    def current_states(self, **classes) -> List[State]:
        temps = self.temperatures()
        states = [host_dict_to_state(self.spec, d, temps[c], **classes) for c, d in enumerate(self.engine.fetch_states())]
        return states

    def misfit_logdet_T(self):
        """[C][3] (misfit, logdet, T) of the current states: the only data a PT swap round needs."""
        return self.engine.misfit_logdet_T()


# --------------------------------------------------------------------------- #
# small device utilities behind the descriptor API
# --------------------------------------------------------------------------- #
def _cuda():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("bayesbay_b200 needs a CUDA device (there is no CPU fallback)")
    return torch


def sample_space_states(space, n: int) -> List[ParameterSpaceState]:
    """Starting states of one parameter space drawn by the device initialisation kernel."""
    from .parameterization import Parameterization

    spec = _compile.compile_problem(Parameterization(space))
    batch = ChainBatch(spec, n, save_dpred=False)
    return [s[space.name] for s in batch.current_states()]


def device_log_prior(space, ps_states: List[ParameterSpaceState]) -> List[float]:
    """``ParameterSpace.log_prior`` through ``bbb_logprior``."""
    torch = _cuda()
    from . import _lib as L
    from ._engine import _PRIOR_KIND

    sp = _compile.compile_space(space)
    P, K, Cn = len(sp["params"]), sp["kmax"], len(ps_states)
    if P == 0:
        return [0.0] * Cn
    cs = L.SpaceSpec()
    cs.n_params, cs.kmax = P, K
    for j, p in enumerate(sp["params"]):
        cs.prior_kind[j], cs.prior_a[j], cs.prior_b[j] = _PRIOR_KIND[p["kind"]], p["a"], p["b"]
    vals = np.zeros((P, K, Cn))
    k = np.zeros(Cn, np.int32)
    for c, st in enumerate(ps_states):
        k[c] = st.n_dimensions
        if k[c] > K:
            raise ValueError("state has more dimensions than n_dimensions_max")
        for j, p in enumerate(sp["params"]):
            vals[j, : k[c], c] = st[p["name"]]
    dv, dk = torch.as_tensor(vals).cuda(), torch.as_tensor(k).cuda()
    out = torch.empty(Cn, dtype=torch.float64, device="cuda")
    lib = L.load()
    L.check(lib.bbb_logprior(C.byref(cs), C.c_void_p(dv.data_ptr()), C.c_void_p(dk.data_ptr()), Cn,
                             C.c_void_p(out.data_ptr()), None))
    return [float(x) for x in out.cpu().numpy()]


def device_interpolate_2d(samples_sites, samples_values, interp_positions, as_tensor=False):
    """Batched ``Voronoi2D.interpolate_tessellation``: every sample is one "chain" of ``bbb_raster2d``."""
    torch = _cuda()
    from . import _lib as L

    pts = np.ascontiguousarray(interp_positions, dtype=np.float64)
    n = len(samples_sites)
    K = max(int(np.asarray(s).shape[0]) for s in samples_sites)
    sites = np.zeros((2, K, n))
    vals = np.zeros((K, n))
    k = np.zeros(n, np.int32)
    for c, (s, v) in enumerate(zip(samples_sites, samples_values)):
        s = np.asarray(s, dtype=np.float64)
        k[c] = s.shape[0]
        sites[:, : k[c], c] = s.T
        vals[: k[c], c] = v
    G = pts.shape[0]
    ds, dk = torch.as_tensor(sites).cuda(), torch.as_tensor(k).cuda()
    px, py = torch.as_tensor(pts[:, 0].copy()).cuda(), torch.as_tensor(pts[:, 1].copy()).cuda()
    idx = torch.empty((G, n), dtype=torch.uint8 if K <= 256 else torch.int16, device="cuda")
    lib = L.load()
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    L.check(lib.bbb_raster2d(p(ds), p(dk), K, n, p(px), p(py), G, p(idx), None))
    idx = idx.long()
    if K > 256:
        idx = idx & 0xFFFF
    fields = torch.gather(torch.as_tensor(vals).cuda(), 0, idx).T.contiguous()  # [n][G]
    return fields if as_tensor else fields.cpu().numpy()


def device_interpolate_1d(samples_sites, samples_values, interp_positions):
    """Batched ``Voronoi1D.interpolate_tessellation(..., 'nuclei')`` through ``bbb_raster1d``."""
    torch = _cuda()
    from . import _lib as L

    x = np.ascontiguousarray(np.atleast_1d(interp_positions), dtype=np.float64)
    n = len(samples_sites)
    K = max(int(np.asarray(s).size) for s in samples_sites)
    sites = np.zeros((K, n))
    vals = np.zeros((K, n))
    k = np.zeros(n, np.int32)
    for c, (s, v) in enumerate(zip(samples_sites, samples_values)):
        k[c] = np.asarray(s).size
        sites[: k[c], c] = s
        vals[: k[c], c] = v
    ds, dv, dk, dx = (torch.as_tensor(a).cuda() for a in (sites, vals, k, x))
    field = torch.empty((x.size, n), dtype=torch.float64, device="cuda")
    lib = L.load()
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    L.check(lib.bbb_raster1d(p(ds), p(dv), p(dk), K, n, p(dx), x.size, None, p(field), None))
    return field.T.cpu().numpy()
