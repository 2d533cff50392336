"""Engine-level parity (through the C ABI): the CUDA chain step against
(a) tape-driven runs of the UNMODIFIED reference (tests/golden/replay_*.npz) and
(b) the oracle driven by the same Philox streams, plus parallel-tempering swap rounds."""
import math
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import chain as OC  # noqa: E402
from oracle import rng as OR  # noqa: E402
from tests import problems  # noqa: E402


def _engine(spec, n_chains, **kw):
    from bayesbay_b200._engine import Engine

    assert torch.cuda.is_available()
    return Engine(spec, n_chains, **kw)


def _check_states(states, g, t, spec, rtol=1e-12):
    sp = spec["spaces"][0]
    for c, st in enumerate(states):
        k = int(g["k"][c, t])
        ss = st["spaces"][0]
        assert ss["k"] == k, (c, t)
        if sp["ndim"] == 2:
            np.testing.assert_allclose(ss["sites"], g["sites"][c, t, :k, :], rtol=rtol, atol=1e-15)
        elif sp["ndim"] == 1:
            np.testing.assert_allclose(ss["sites"], g["sites"][c, t, :k, 0], rtol=rtol, atol=1e-15)
        for j in range(len(sp["params"])):
            np.testing.assert_allclose(ss["values"][j], g["values"][c, t, j, :k], rtol=rtol, atol=1e-15)
        if not np.isnan(g["sigma"][c, t]):
            np.testing.assert_allclose(st["sigma"][0], g["sigma"][c, t], rtol=rtol)
        if "corr" in g.files and not np.isnan(g["corr"][c, t]):
            np.testing.assert_allclose(st["corr"][0], g["corr"][c, t], rtol=rtol)


@pytest.mark.parametrize("name", list(problems.REPLAYS))
def test_step_replay_against_reference(name, golden_dir):
    """Every chain consumes the tape the reference consumed: initial state, move choice, prior/proposal
    ratio, log-likelihood ratio, accept bit, resulting state and tape position must all agree."""
    g = np.load(os.path.join(golden_dir, f"replay_{name}.npz"))
    spec = problems.spec_from_ingredients(problems.REPLAYS[name]())
    n_chains, n_steps = g["lpr"].shape
    eng = _engine(spec, n_chains, tape=g["tape"])
    eng.init_states()
    assert np.array_equal(eng.draw_cursor.cpu().numpy(), g["cursor"][:, 0])
    _check_states(eng.fetch_states(), g, 0, spec)
    for t in range(n_steps):
        eng.step(1)
        move = eng.move.cpu().numpy()
        lpr = eng.log_prob_ratio.cpu().numpy()
        llr = eng.log_like_ratio.cpu().numpy()
        acc = eng.accepted.cpu().numpy()
        assert np.array_equal(move, g["move"][:, t]), t
        assert np.array_equal(eng.draw_cursor.cpu().numpy(), g["cursor"][:, t + 1]), t
        ref_lpr = g["lpr"][:, t]
        inf = np.isinf(ref_lpr)
        assert np.array_equal(np.isinf(lpr), inf)
        np.testing.assert_allclose(lpr[~inf], ref_lpr[~inf], rtol=1e-12, atol=1e-12)
        # the log-likelihood ratio is a DIFFERENCE of misfits and log-determinants,
        # (logdet_old - logdet_new + misfit_old - misfit_new) / 2 / T (likelihood/_log_likelihood.py:245-251): 1e-12
        # relative on those terms is 1e-12 of their magnitude on the ratio, whatever is left after the cancellation
        mis_now, ld_now = (a.cpu().numpy() for a in eng.misfit_logdet())
        scale = np.abs(mis_now) + np.abs(ld_now) + 2.0 * np.abs(g["llr"][:, t]) + 1.0
        assert (np.abs(llr - g["llr"][:, t])[~inf] <= 1e-12 * scale[~inf]).all(), t
        assert np.array_equal(acc, g["accepted"][:, t]), t
        _check_states(eng.fetch_states(), g, t + 1, spec)
    # statistics accumulate like _save_statistics
    prop = eng.n_proposed.cpu().numpy()
    accd = eng.n_accepted.cpu().numpy()
    for c in range(n_chains):
        assert np.array_equal(prop[:, c], np.bincount(g["move"][c], minlength=5))
        assert np.array_equal(accd[:, c], np.bincount(g["move"][c], weights=g["accepted"][c], minlength=5))
    assert (eng.iteration.cpu().numpy() == n_steps).all()


def _oracle_state(st):
    spaces = [OC.SpaceState(sp["k"], None if sp["sites"] is None else np.array(sp["sites"]),
                            [np.array(v) for v in sp["values"]]) for sp in st["spaces"]]
    n_t = len(st["sigma"])
    return OC.ChainState(spaces, list(st["sigma"]), 1.0, [None] * n_t, list(st.get("corr") or [None] * n_t))


@pytest.mark.parametrize("name,n_chains,n_steps", [("tomo_small", 96, 40), ("partition_1d", 70, 60),
                                                   ("partition_1d_nb", 33, 60), ("partition_1d_corr", 40, 60),
                                                   ("partition_1d_posdep", 50, 80), ("tomo_small_corr", 70, 40),
                                                   ("partition_1d_fixed", 40, 50), ("transd_polyfit", 90, 80),
                                                   ("polyfit", 64, 50)])
def test_philox_steps_against_oracle(name, n_chains, n_steps):
    """Counter-based streams: device and oracle draw the same Philox4x32-10 uniforms, so whole
    trajectories must agree (until a 1-ulp difference flips a decision, which the test tolerates
    only by comparing chains that are still in step)."""
    spec = problems.spec_from_ingredients(problems.ALL[name]())
    seed, id0 = 1234567890123, 1000
    eng = _engine(spec, n_chains, seed=seed, chain_id0=id0)
    eng.init_states()
    dev0 = eng.fetch_states()
    states = []
    for c in range(n_chains):
        st = OC.init_state(spec, OR.PhiloxStream(seed, id0 + c, 0, OR.STREAM_INIT))
        states.append(st)
        d = _oracle_state(dev0[c])
        assert d.spaces[0].k == st.spaces[0].k
        if st.spaces[0].sites is not None:
            np.testing.assert_allclose(d.spaces[0].sites, st.spaces[0].sites, rtol=1e-14)
        for a, b in zip(d.spaces[0].values, st.spaces[0].values):
            np.testing.assert_allclose(a, b, rtol=1e-13, atol=1e-15)
        if st.sigma[0] is not None:
            np.testing.assert_allclose(d.sigma[0], st.sigma[0], rtol=1e-14)
        if st.corr[0] is not None:
            np.testing.assert_allclose(d.corr[0], st.corr[0], rtol=1e-14)
    in_step = np.ones(n_chains, bool)
    for t in range(n_steps):
        eng.step(1)
        move = eng.move.cpu().numpy()
        lpr = eng.log_prob_ratio.cpu().numpy()
        llr = eng.log_like_ratio.cpu().numpy()
        acc = eng.accepted.cpu().numpy()
        for c in range(n_chains):
            if not in_step[c]:
                continue
            states[c], info = OC.step(spec, states[c], OR.PhiloxStream(seed, id0 + c, t))
            assert info["move"] == move[c], (c, t)
            if math.isinf(info["log_prob_ratio"]):
                assert lpr[c] == info["log_prob_ratio"]
            else:
                np.testing.assert_allclose(lpr[c], info["log_prob_ratio"], rtol=1e-10, atol=1e-11)
                np.testing.assert_allclose(llr[c], info["log_like_ratio"], rtol=1e-8, atol=1e-8)
            if bool(acc[c]) != info["accepted"]:
                in_step[c] = False  # decision on a knife edge; stop following this chain
    assert in_step.mean() > 0.95
    final = eng.fetch_states()
    for c in np.flatnonzero(in_step):
        d = _oracle_state(final[c])
        assert d.spaces[0].k == states[c].spaces[0].k
        for a, b in zip(d.spaces[0].values, states[c].spaces[0].values):
            np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-12)
    # cached misfit equals a fresh evaluation of the final states (no drift in the double-buffering)
    mis, ld = eng.misfit_logdet()
    eng.refresh()
    mis2, ld2 = eng.misfit_logdet()
    np.testing.assert_allclose(mis.cpu().numpy(), mis2.cpu().numpy(), rtol=1e-13)
    np.testing.assert_allclose(ld.cpu().numpy(), ld2.cpu().numpy(), rtol=1e-13)
    for c in np.flatnonzero(in_step)[:8]:
        m, l_ = OC.misfit_and_logdet(spec, states[c])
        np.testing.assert_allclose(mis[c].item(), m, rtol=1e-12)
        np.testing.assert_allclose(ld[c].item(), l_, rtol=1e-12)


def test_dpred_and_cell_index_of_current_states():
    spec = problems.spec_from_ingredients(problems.tomo_small())
    eng = _engine(spec, 40, seed=7)
    eng.init_states()
    eng.step(25)
    states = eng.fetch_states()
    dp = eng.dpred(0).cpu().numpy()
    idx = eng.current_cell_idx(0).cpu().numpy()
    fwd = spec["targets"][0]["forward"]
    from oracle import kernels as OK

    for c in range(40):
        ss = states[c]["spaces"][0]
        ref, ridx = OK.tomography_forward(fwd, ss["sites"], ss["values"][0])
        assert np.array_equal(idx[:, c], ridx), c
        np.testing.assert_allclose(dp[:, c], ref, rtol=1e-12)


def test_pt_swap_golden_and_oracle(golden_dir):
    from bayesbay_b200._engine import pt_swap

    g = np.load(os.path.join(golden_dir, "pt_swap.npz"))
    for r in range(g["misfit"].shape[0]):
        gathered = torch.as_tensor(np.stack((g["misfit"][r], g["logdet"][r], g["t_before"][r]), axis=1).copy()).cuda()
        tape = torch.as_tensor(g["tape"][int(g["cursor0"][r]): int(g["cursor1"][r])].copy()).cuda()
        t_out, n_acc = pt_swap(gathered, 0, 0, 0, 6, tape=tape)
        assert np.array_equal(t_out.cpu().numpy(), g["t_after"][r])
        assert np.array_equal(gathered.cpu().numpy()[:, 2], g["t_after"][r])
    # Philox mode against the oracle, sharded output, many chains
    rng = np.random.default_rng(0)
    for n in (2, 7, 64, 1000, 8192):
        mis = rng.uniform(50, 150, n)
        ld = rng.uniform(-20, 20, n)
        temps = OC.pt_ladder(n, 5.0, 0.4) if n > 2 else np.array([1.0, 5.0])
        ref, n_ref = OC.pt_swap_round(mis, ld, list(temps), OR.PhiloxStream(99, 0, 3, OR.STREAM_SWAP))
        gathered = torch.as_tensor(np.stack((mis, ld, temps), axis=1).copy()).cuda()
        id0, nl = n // 4, n - n // 4 - n // 3
        t_out, n_acc = pt_swap(gathered, 99, 3, id0, nl)
        assert np.array_equal(t_out.cpu().numpy(), np.array(ref)[id0: id0 + nl]), n
        assert int(n_acc.item()) == n_ref


def test_incremental_raster_equals_brute_force_at_scale():
    """Size-independent property at a BASELINE-like size: after many accepted births/deaths/moves the resident
    nearest-site index (maintained change-locally) is bit-identical to a from-scratch rasterisation of the
    current sites, and the cached misfit equals a fresh forward evaluation."""
    import ctypes as C

    from bayesbay_b200 import _lib as L
    from bayesbay_b200 import synthetic

    ing = synthetic.tomography_2d(nx=64, ny=48, n_sources_per_side=6, n_receivers=20, kmin=3, kmax=300)
    spec = problems.spec_from_ingredients(ing)
    eng = _engine(spec, 200, seed=31)  # K = 300 -> uint16 index path, 200 chains -> ragged last chain tile
    eng.init_states()
    for _ in range(4):
        eng.step(150)
        k, sites, _ = eng.gather_space(0)
        G = ing["points"].shape[0]
        ref = torch.zeros((G, 200), dtype=torch.int16, device="cuda")
        px = torch.as_tensor(ing["points"][:, 0].copy()).cuda()
        py = torch.as_tensor(ing["points"][:, 1].copy()).cuda()
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        L.check(L.load().bbb_raster2d(p(sites), p(k), 300, 200, p(px), p(py), G, p(ref), None))
        assert torch.equal(eng.current_cell_idx(0), ref)
        m0, _ = eng.misfit_logdet()
        eng.refresh()
        m1, _ = eng.misfit_logdet()
        np.testing.assert_allclose(m0.cpu().numpy(), m1.cpu().numpy(), rtol=1e-12)
    assert eng.n_accepted.sum().item() > 0.2 * eng.n_proposed.sum().item()
    kk = eng.gather_space(0)[0].cpu().numpy()
    assert kk.min() >= 3 and kk.max() <= 300 and kk.std() > 0


@pytest.mark.parametrize("name,n_chains,n_steps", [("joint_small", 80, 80), ("tomo_joint_small", 48, 60),
                                                   ("tomo_weighted_fixed", 60, 80), ("tomo_scalar_prior_birth", 60, 80),
                                                   ("joint_matrix_cov", 64, 80), ("tomo_matrix_cov", 48, 60)])
def test_multi_space_multi_target_against_oracle(name, n_chains, n_steps):
    """Several parameter spaces and targets (hierarchical, per-datum and scalar inverse covariance; linear,
    1-D lookup and ray forwards side by side): move ids, outer weights and the summed likelihood ratio
    (LogLikelihood._get_misfit_and_det over all targets) against the oracle on shared Philox streams."""
    spec = problems.SPECS[name]()
    seed, id0 = 424242, 7
    eng = _engine(spec, n_chains, seed=seed, chain_id0=id0)
    eng.init_states()
    states = [OC.init_state(spec, OR.PhiloxStream(seed, id0 + c, 0, OR.STREAM_INIT)) for c in range(n_chains)]
    for c, d in enumerate(eng.fetch_states()):
        o = _oracle_state(d)
        for a, b in zip(o.spaces, states[c].spaces):
            assert a.k == b.k
            for va, vb in zip(a.values, b.values):
                np.testing.assert_allclose(va, vb, rtol=1e-13, atol=1e-15)
    in_step = np.ones(n_chains, bool)
    seen = set()
    for t in range(n_steps):
        eng.step(1)
        move, lpr = eng.move.cpu().numpy(), eng.log_prob_ratio.cpu().numpy()
        llr, acc = eng.log_like_ratio.cpu().numpy(), eng.accepted.cpu().numpy()
        for c in range(n_chains):
            if not in_step[c]:
                continue
            states[c], info = OC.step(spec, states[c], OR.PhiloxStream(seed, id0 + c, t))
            assert info["move"] == move[c], (c, t)
            seen.add(int(move[c]))
            if math.isinf(info["log_prob_ratio"]):
                assert lpr[c] == info["log_prob_ratio"]
            else:
                np.testing.assert_allclose(lpr[c], info["log_prob_ratio"], rtol=1e-10, atol=1e-11)
                np.testing.assert_allclose(llr[c], info["log_like_ratio"], rtol=1e-8, atol=1e-8)
            if bool(acc[c]) != info["accepted"]:
                in_step[c] = False
    assert in_step.mean() > 0.9
    if name.endswith("joint_small"):
        assert len(seen) >= 6  # moves of both spaces and the noise move all occurred
    elif name == "joint_matrix_cov":
        assert len(seen) >= 5
    elif name == "tomo_matrix_cov":
        assert not eng.chain_local and len(seen) >= 3  # full inverse covariance: the full-SpMM step
    else:
        assert eng.chain_local and len(seen) >= 2
    mis, ld = eng.misfit_logdet()
    for c in np.flatnonzero(in_step)[:10]:
        m, l_ = OC.misfit_and_logdet(spec, states[c])
        np.testing.assert_allclose(mis[c].item(), m, rtol=1e-10)
        np.testing.assert_allclose(ld[c].item(), l_, rtol=1e-12, atol=1e-12)


def test_full_size_config3_properties():
    """BASELINE configs[2] at full size (100x100 grid, 10 000 rays, K = 200, 1024 chains): after a few hundred
    iterations the resident nearest-site index equals a from-scratch rasterisation, the cached misfit equals a fresh
    forward evaluation, d_pred of sampled chains equals the oracle (KD-tree + CSR mat-vec restatement), every chain
    respects its bounds, and the move counters add up."""
    import ctypes as C

    import scipy.sparse
    import scipy.spatial

    from bayesbay_b200 import _lib as L
    from bayesbay_b200 import synthetic

    ing = synthetic.tomography_2d()
    spec = problems.spec_from_ingredients(ing)
    n = 1024
    eng = _engine(spec, n, seed=2025)
    eng.init_states()
    eng.step(300)
    k, sites, vals = eng.gather_space(0)
    G = ing["points"].shape[0]
    ref = torch.zeros((G, n), dtype=torch.uint8, device="cuda")
    px = torch.as_tensor(ing["points"][:, 0].copy()).cuda()
    py = torch.as_tensor(ing["points"][:, 1].copy()).cuda()
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    L.check(L.load().bbb_raster2d(p(sites), p(k), 200, n, p(px), p(py), G, p(ref), None))
    assert torch.equal(eng.current_cell_idx(0), ref)
    dp = eng.dpred(0)
    m0, ld0 = eng.misfit_logdet()
    eng.refresh()
    m1, ld1 = eng.misfit_logdet()
    np.testing.assert_allclose(m0.cpu().numpy(), m1.cpu().numpy(), rtol=1e-12)
    assert torch.equal(ld0, ld1)
    jac = scipy.sparse.csr_matrix((ing["data"], ing["indices"], ing["indptr"]), shape=(10000, G))
    kk, ss, vv = k.cpu().numpy(), sites.cpu().numpy(), vals.cpu().numpy()
    sig = eng.sigma[0][0].cpu().numpy()
    for c in (0, 17, 511, 1023):
        s_c, v_c = ss[:, : kk[c], c].T, vv[0, : kk[c], c]
        nearest = scipy.spatial.KDTree(s_c).query(ing["points"])[1]
        assert np.array_equal(nearest, ref[:, c].cpu().numpy())
        want = jac @ (1.0 / v_c[nearest])
        np.testing.assert_allclose(dp[:, c].cpu().numpy(), want, rtol=1e-12)
        res = want - ing["d_obs"]
        np.testing.assert_allclose(m1[c].item(), float(res @ res) / sig[c] ** 2, rtol=1e-12)
    assert kk.min() >= 10 and kk.max() <= 200
    assert (vv[0][np.arange(200)[:, None] < kk[None, :]] >= 2).all() and (vv[0][np.arange(200)[:, None] < kk[None, :]] <= 4).all()
    assert (sig >= 0).all() and (sig <= 0.01).all()
    prop = eng.n_proposed.cpu().numpy()
    assert (prop.sum(0) == 300).all() and (eng.iteration.cpu().numpy() == 300).all()
    share = prop.sum(1) / prop.sum()
    np.testing.assert_allclose(share, [1 / 7, 1 / 7, 3 / 7, 1 / 7, 1 / 7], atol=0.005)


def test_full_size_config5_properties():
    """BASELINE configs[4] at full problem size (256x256 grid, 100 000 rays, nnz 21 M, K = 500 -> 16-bit nearest-site
    index, 10 ray tiles per chain-local step) on 96 chains: after 120 iterations the resident index equals a
    from-scratch rasterisation, the incrementally maintained misfit equals a fresh forward evaluation, d_pred of
    sampled chains equals the KD-tree + CSR mat-vec restatement, bounds hold and the counters add up."""
    import ctypes as C

    import scipy.sparse
    import scipy.spatial

    from bayesbay_b200 import _lib as L
    from bayesbay_b200 import synthetic

    ing = synthetic.tomography_2d(nx=256, ny=256, n_sources_per_side=100, n_receivers=250, kmin=50, kmax=500)
    spec = problems.spec_from_ingredients(ing)
    R, G = ing["indptr"].size - 1, ing["points"].shape[0]
    assert R == 100_000 and G == 65_536
    n, n_it = 96, 120
    eng = _engine(spec, n, seed=77)
    assert eng.chain_local
    eng.init_states()
    eng.step(n_it)
    k, sites, vals = eng.gather_space(0)
    ref = torch.zeros((G, n), dtype=torch.int16, device="cuda")
    px = torch.as_tensor(ing["points"][:, 0].copy()).cuda()
    py = torch.as_tensor(ing["points"][:, 1].copy()).cuda()
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    L.check(L.load().bbb_raster2d(p(sites), p(k), 500, n, p(px), p(py), G, p(ref), None))
    assert torch.equal(eng.current_cell_idx(0).to(torch.int64), ref.to(torch.int64))
    dp = eng.dpred(0)
    m0, ld0 = eng.misfit_logdet()
    eng.refresh()
    m1, ld1 = eng.misfit_logdet()
    np.testing.assert_allclose(m0.cpu().numpy(), m1.cpu().numpy(), rtol=1e-12)
    assert torch.equal(ld0, ld1)
    jac = scipy.sparse.csr_matrix((ing["data"], ing["indices"], ing["indptr"]), shape=(R, G))
    kk, ss, vv = k.cpu().numpy(), sites.cpu().numpy(), vals.cpu().numpy()
    for c in (0, 41, 95):
        s_c, v_c = ss[:, : kk[c], c].T, vv[0, : kk[c], c]
        nearest = scipy.spatial.KDTree(s_c).query(ing["points"])[1]
        assert np.array_equal(nearest, ref[:, c].cpu().numpy().astype(np.int64))
        np.testing.assert_allclose(dp[:, c].cpu().numpy(), jac @ (1.0 / v_c[nearest]), rtol=1e-12)
    assert kk.min() >= 50 and kk.max() <= 500
    prop = eng.n_proposed.cpu().numpy()
    assert (prop.sum(0) == n_it).all() and (eng.iteration.cpu().numpy() == n_it).all()


def _pair_of_engines(monkeypatch, spec, n_chains, seed, tile_cap=None):
    """The same problem on the full-SpMM step (reference behaviour: every proposal re-evaluates all rays) and on
    the chain-local step (change-local forward)."""
    monkeypatch.setenv("BBB_FULL_FORWARD", "1")
    full = _engine(spec, n_chains, seed=seed)
    monkeypatch.setenv("BBB_FULL_FORWARD", "0")
    if tile_cap is not None:
        monkeypatch.setenv("BBB_CHAIN_TILE_RAYS", str(tile_cap))
    local = _engine(spec, n_chains, seed=seed)
    monkeypatch.delenv("BBB_CHAIN_TILE_RAYS", raising=False)
    assert not full.chain_local and local.chain_local
    return full, local


@pytest.mark.parametrize("name,n_chains,n_steps,tile_cap", [("small", 64, 80, None), ("small", 48, 80, 16),
                                                            ("mid", 120, 200, None), ("mid", 60, 120, 200)])
def test_chain_local_step_equals_full_forward(monkeypatch, name, n_chains, n_steps, tile_cap):
    """Change-local forward (columns of the ray matrix, fixed-point ray accumulators, incremental residuals)
    against the full SpMM on the same Philox streams: same moves, same decisions, log-likelihood ratios within
    1e-12 of the misfit scale, bit-identical nearest-site index, residuals and misfit equal to a from-scratch
    evaluation.  `tile_cap` forces the multi-tile path (several passes over the ray accumulators)."""
    from bayesbay_b200 import synthetic

    ing = problems.tomo_small() if name == "small" else synthetic.tomography_2d(
        nx=64, ny=48, n_sources_per_side=6, n_receivers=20, kmin=3, kmax=300)
    spec = problems.spec_from_ingredients(ing)
    full, local = _pair_of_engines(monkeypatch, spec, n_chains, 5, tile_cap)
    full.init_states()
    local.init_states()
    alive = np.ones(n_chains, bool)
    for t in range(n_steps):
        full.step(1)
        local.step(1)
        assert np.array_equal(full.move.cpu().numpy()[alive], local.move.cpu().numpy()[alive]), t
        l0, l1 = full.log_like_ratio.cpu().numpy(), local.log_like_ratio.cpu().numpy()
        m0 = full.misfit_logdet()[0].cpu().numpy()
        fin = alive & np.isfinite(l0)
        np.testing.assert_allclose(l1[fin], l0[fin], rtol=1e-9, atol=1e-11 * np.abs(m0[fin]).max())
        flip = alive & (full.accepted.cpu().numpy() != local.accepted.cpu().numpy())
        alive &= ~flip  # a decision on a knife edge: stop following the chain
    assert alive.mean() > 0.97
    same = (full.current_cell_idx(0) == local.current_cell_idx(0)).all(0).cpu().numpy()
    assert same[alive].all()
    m_full = full.misfit_logdet()[0].cpu().numpy()
    m_loc = local.misfit_logdet()[0].cpu().numpy()
    np.testing.assert_allclose(m_loc[alive], m_full[alive], rtol=1e-12)
    res_inc = local.res[0].clone()
    local.resync()
    np.testing.assert_allclose(local.misfit_logdet()[0].cpu().numpy(), m_loc, rtol=1e-13)
    scale = local.res[0].abs().max().item()
    assert (res_inc - local.res[0]).abs().max().item() <= 1e-13 * max(scale, 1e-300)
    dp = local.dpred(0)
    dobs = torch.as_tensor(np.asarray(spec["targets"][0]["dobs"])).cuda()
    np.testing.assert_allclose((local.res[0].t() + dobs[:, None]).cpu().numpy(), dp.cpu().numpy(), rtol=1e-12)


@pytest.mark.parametrize("tile_cap", [None, 200])
def test_change_list_spill_to_global_memory_is_bit_identical(monkeypatch, tile_cap):
    """The shared-memory head of the change list only decides how many entries spill to the chain's global list (the
    launch trims it to fit the SM's shared-memory carve-out steps): a head of 8 entries (nearly everything spills) gives
    bit-identical trajectories, index and residuals to the default head, single- and multi-tile."""
    from bayesbay_b200 import synthetic

    ing = synthetic.tomography_2d(nx=64, ny=48, n_sources_per_side=6, n_receivers=20, kmin=3, kmax=300)
    spec = problems.spec_from_ingredients(ing)
    if tile_cap is not None:
        monkeypatch.setenv("BBB_CHAIN_TILE_RAYS", str(tile_cap))
    out = []
    for cap in (None, 8):
        if cap is None:
            monkeypatch.delenv("BBB_CHAIN_LIST_CAP", raising=False)
        else:
            monkeypatch.setenv("BBB_CHAIN_LIST_CAP", str(cap))
        eng = _engine(spec, 80, seed=9)
        eng.init_states()
        for _ in range(30):
            eng.step(1)
        eng.step(40)
        assert eng.chain_local and eng.ray_tiles(0) == (1 if tile_cap is None else eng.ray_tiles(0)) and (tile_cap is None or eng.ray_tiles(0) > 1)
        out.append((eng.pack_states(300).clone(), eng.res[0].clone(), eng.idx_cm[0].clone(), eng.n_accepted.clone()))
    monkeypatch.delenv("BBB_CHAIN_LIST_CAP", raising=False)
    for a, b in zip(*out):
        assert torch.equal(a, b)


def test_chain_local_step_is_bit_reproducible_and_shard_independent():
    """The ray accumulators are fixed point, so neither the thread timing of a launch nor the way chains are
    sharded over engines may change a single bit of the trajectories."""
    spec = problems.spec_from_ingredients(problems.tomo_small())

    def run(n, id0):
        eng = _engine(spec, n, seed=77, chain_id0=id0)
        assert eng.chain_local
        eng.init_states()
        eng.step(120)
        k, sites, vals = eng.gather_space(0)
        return k, sites, vals, eng.phi[0][0, 0].clone(), eng.res[0].clone(), eng.n_accepted.clone()

    a, b = run(96, 0), run(96, 0)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    lo, hi = run(40, 0), run(56, 40)
    assert torch.equal(torch.cat((lo[0], hi[0])), a[0])
    assert torch.equal(torch.cat((lo[2], hi[2]), dim=2), a[2])
    assert torch.equal(torch.cat((lo[3], hi[3])), a[3])
    assert torch.equal(torch.cat((lo[4], hi[4])), a[4])


@pytest.mark.parametrize("name", ["tomo_small", "partition_1d_corr", "polyfit"])
def test_packed_state_image_round_trip(name):
    """bbb_engine_pack/compare/unpack_states: the packed image is bit-identical to the chains' current states,
    detects any change, and restoring it reproduces states and misfit exactly."""
    spec = problems.spec_from_ingredients(problems.ALL[name]())
    n = 37
    eng = _engine(spec, n, seed=3)
    eng.init_states()
    eng.step(30)
    kmax = spec["spaces"][0]["kmax"]
    kb = min(kmax, int(eng.gather_space(0)[0].max().item()) + 1)
    img = eng.pack_states(kb).clone()
    assert img.numel() * 8 == eng.state_image_bytes(kb)
    before = eng.fetch_states()
    m0, l0 = eng.misfit_logdet()
    flag = torch.zeros((1,), dtype=torch.int32, device="cuda")
    assert int(eng.states_differ(kb, img, flag).item()) == 0
    # k row of the image = current k; padding cells are zero
    k = eng.gather_space(0)[0]
    assert torch.equal(img[:n], k.long())
    eng.step(7)
    flag.zero_()
    assert int(eng.states_differ(kb, img, flag).item()) == 1
    eng.unpack_states(kb, img)
    flag.zero_()
    assert int(eng.states_differ(kb, img, flag).item()) == 0
    after = eng.fetch_states()
    for a, b in zip(before, after):
        assert a["spaces"][0]["k"] == b["spaces"][0]["k"]
        if a["spaces"][0]["sites"] is not None:
            assert np.array_equal(a["spaces"][0]["sites"], b["spaces"][0]["sites"])
        for va, vb in zip(a["spaces"][0]["values"], b["spaces"][0]["values"]):
            assert np.array_equal(va, vb)
        assert a["sigma"] == b["sigma"] and a["corr"] == b["corr"]
    m1, l1 = eng.misfit_logdet()
    np.testing.assert_allclose(m1.cpu().numpy(), m0.cpu().numpy(), rtol=1e-12)
    assert torch.equal(l0, l1)
    eng.step(5)  # and the chains go on
    assert (eng.iteration.cpu().numpy() == 35).all()


def test_pipelined_halves_are_bit_identical(monkeypatch):
    """Multi-iteration calls run parts of the chain set on internal streams; chains never interact inside
    a step, so every bit must equal the single-stream run (including across a from-scratch re-evaluation)."""
    from bayesbay_b200 import synthetic

    ing = synthetic.tomography_2d(nx=32, ny=32, n_sources_per_side=5, n_receivers=12, kmin=4, kmax=60)
    spec = problems.spec_from_ingredients(ing)
    n = 2 * 148 + 37  # enough chains for the pipelined path, ragged halves

    def run(pipeline):
        monkeypatch.setenv("BBB_PIPELINE", str(pipeline))
        eng = _engine(spec, n, seed=11)
        eng.init_states()
        eng.step(1)
        eng.step(300)  # crosses the resync at 256
        eng.step(45)
        k, sites, vals = eng.gather_space(0)
        return k, vals, eng.phi[0][0, 0].clone(), eng.res[0].clone(), eng.idx_cm[0].clone(), eng.n_accepted.clone(), eng.iteration.clone()

    a = run(1)
    for parts in (2, 4):
        for x, y in zip(a, run(parts)):
            assert torch.equal(x, y)
    assert (a[-1] == 346).all()


@pytest.mark.parametrize("pipeline,graph", [(1, 1), (2, 1), (4, 1), (4, 0)])
def test_hosted_step_equals_resident_step(monkeypatch, pipeline, graph):
    """bbb_engine_step_hosted: chains held by the caller in pinned host memory, copies overlapped with the parts'
    kernels.  (a) Fast path: every bit of the downloaded states equals the resident engine's; no part is redone.
    (b) The caller edits a chain of ONE part: that part is detected on the device, left untouched by the guarded
    kernels, written + re-derived + advanced by the wait; the other parts are not redone; all chains advance once."""
    from bayesbay_b200 import synthetic

    monkeypatch.setenv("BBB_HOSTED_PARTS", str(pipeline))
    monkeypatch.setenv("BBB_HOSTED_GRAPH", str(graph))  # replayed from captured CUDA graphs / enqueued every call
    ing = synthetic.tomography_2d(nx=32, ny=32, n_sources_per_side=5, n_receivers=12, kmin=4, kmax=60)
    spec = problems.spec_from_ingredients(ing)
    n = 148 + 37
    K = spec["spaces"][0]["kmax"]
    ref, eng = _engine(spec, n, seed=5), _engine(spec, n, seed=5)
    for e in (ref, eng):
        e.init_states()
        e.step(20)
    first = eng.hosted_parts()
    assert first[0] == 0 and first[-1] == n and len(first) == pipeline + 1
    cap = eng.state_image_bytes(K) // 8
    host = [torch.empty((cap,), dtype=torch.int64, pin_memory=True) for _ in range(2)]
    stage = torch.empty((cap,), dtype=torch.int64, device="cuda")
    out = torch.empty((cap,), dtype=torch.int64, device="cuda")
    kb = min(K, int(eng.gather_space(0)[0].max().item()))
    nw = eng.state_image_bytes(kb) // 8
    host[0][:nw].copy_(eng.pack_states_hosted(kb))
    torch.cuda.synchronize()
    # a hosted image is the parts' images back to back
    rows = nw // n
    whole = eng.pack_states(kb).view(rows, n)
    for c0, c1 in zip(first[:-1], first[1:]):
        assert torch.equal(host[0][rows * c0:rows * c1].view(rows, c1 - c0), whole[:, c0:c1].cpu())
    n_steps = 300  # crosses the from-scratch re-evaluation at 256
    for it in range(n_steps):
        kb_new = eng.step_hosted(kb, host[it & 1], host[(it + 1) & 1], stage, out)[0]
        assert eng.step_hosted_wait() == 0
        assert kb_new <= kb + 8 and (kb_new % 8 == 0 or kb_new == K)
        kb = kb_new
        ref.step(1)
    nw = eng.state_image_bytes(kb) // 8
    assert torch.equal(host[n_steps & 1][:nw], ref.pack_states_hosted(kb).cpu())
    assert torch.equal(eng.pack_states(kb), ref.pack_states(kb))
    assert torch.equal(eng.res[0], ref.res[0]) and torch.equal(eng.idx_cm[0], ref.idx_cm[0])
    assert torch.equal(eng.n_accepted, ref.n_accepted)

    # (b) the caller edits one chain of the LAST part (value of its first cell)
    src = host[n_steps & 1]
    rows = nw // n
    c_edit = n - 3
    part = max(h for h in range(len(first) - 1) if first[h] <= c_edit)
    c0, c1 = first[part], first[part + 1]
    value_row = 1 + 2 * kb  # k row, 2 x kb site rows, then the values
    pos = rows * c0 + value_row * (c1 - c0) + (c_edit - c0)
    new_value = src[pos:pos + 1].view(torch.float64) * 1.03125
    src[pos:pos + 1] = new_value.view(torch.int64)
    dst = host[(n_steps + 1) & 1]
    kb2 = eng.step_hosted(kb, src, dst, stage, out)[0]
    assert eng.step_hosted_wait() == 1
    assert (eng.iteration.cpu().numpy() == 20 + n_steps + 1).all()
    nw2 = eng.state_image_bytes(kb2) // 8
    assert torch.equal(dst[:nw2], eng.pack_states_hosted(kb2).cpu())
    # the other parts advanced exactly like the resident engine; the edited part went through unpack + refresh + step
    ref.step(1)
    rows2 = nw2 // n
    ref_img = ref.pack_states_hosted(kb2).cpu()
    for h, (a0, a1) in enumerate(zip(first[:-1], first[1:])):
        if h != part:
            assert torch.equal(dst[rows2 * a0:rows2 * a1], ref_img[rows2 * a0:rows2 * a1])
    # edited part: same chains except the edited one (its caches were re-derived from scratch: the decision can only
    # differ where the misfit rounding decides, which these seeds do not hit)
    a = dst[rows2 * c0:rows2 * c1].view(rows2, c1 - c0)
    b = ref_img[rows2 * c0:rows2 * c1].view(rows2, c1 - c0)
    keep = [c for c in range(c1 - c0) if c != c_edit - c0]
    assert torch.equal(a[:, keep], b[:, keep])
    # the engine carries on from the edited state, caches consistent: resident steps, then a from-scratch check
    eng.step(10)
    res_inc = eng.res[0].clone()
    eng.resync()
    np.testing.assert_allclose(eng.res[0].cpu().numpy(), res_inc.cpu().numpy(), rtol=0, atol=1e-9)
    with pytest.raises(Exception):
        eng.step_hosted(kb2, dst, dst, stage, out)


@pytest.mark.parametrize("pipeline,graph", [(1, 1), (2, 1), (4, 0)])
def test_hosted_step_ragged_image(monkeypatch, pipeline, graph):
    """bbb_engine_step_hosted_ragged: the hosted step on the RAGGED wire format (only the cells in use travel).
    (a) The ragged image decodes to the chains' states; fast path bit-identical to resident stepping over 300
    iterations (incl. a from-scratch re-evaluation), never redone, and fewer bytes than the kb-padded rows.
    (b) An edited chain is detected, its part redone, every chain advances once.  (c) An invalid k is refused."""
    from bayesbay_b200 import synthetic

    monkeypatch.setenv("BBB_HOSTED_PARTS", str(pipeline))
    monkeypatch.setenv("BBB_HOSTED_GRAPH", str(graph))
    ing = synthetic.tomography_2d(nx=32, ny=32, n_sources_per_side=5, n_receivers=12, kmin=4, kmax=60)
    spec = problems.spec_from_ingredients(ing)
    n = 148 + 37
    K = spec["spaces"][0]["kmax"]
    ref, eng = _engine(spec, n, seed=5), _engine(spec, n, seed=5)
    for e in (ref, eng):
        e.init_states()
        e.step(20)
    cap = eng.hosted_ragged_words()
    host = [torch.zeros((cap,), dtype=torch.int64, pin_memory=True) for _ in range(2)]
    stage = torch.zeros((cap,), dtype=torch.int64, device="cuda")
    out = torch.zeros((cap,), dtype=torch.int64, device="cuda")
    host[0].copy_(eng.pack_states_ragged())
    torch.cuda.synchronize()

    def decode(e, image):
        """per chain (k, cells [k][3], sigma, temperature, iteration) from a ragged image"""
        chains = {}
        for c0, c1, words in e.ragged_parts(image):
            n_c = c1 - c0
            k = words[:n_c].numpy()
            sig, temp = words[n_c:2 * n_c].view(torch.float64).numpy(), words[2 * n_c:3 * n_c].view(torch.float64).numpy()
            it = words[3 * n_c:4 * n_c].numpy()
            cells = words[4 * n_c:].view(torch.float64).numpy().reshape(-1, 3)
            off = np.concatenate([[0], np.cumsum(k)])
            assert off[-1] == cells.shape[0]
            for lc in range(n_c):
                chains[c0 + lc] = (int(k[lc]), cells[off[lc]:off[lc + 1]].copy(), float(sig[lc]), float(temp[lc]), int(it[lc]))
        return chains

    def check_against_engine(e, image):
        got = decode(e, image)
        k, sites, vals = (v.cpu().numpy() for v in e.gather_space(0))
        sig, temp, it = e.sigma[0][0].cpu().numpy(), e.temperature.cpu().numpy(), e.iteration.cpu().numpy()
        assert sorted(got) == list(range(n))
        for c, (kc, cells, s_, t_, i_) in got.items():
            assert kc == k[c] and s_ == sig[c] and t_ == temp[c] and i_ == it[c]
            assert np.array_equal(cells[:, 0], sites[0, :kc, c]) and np.array_equal(cells[:, 1], sites[1, :kc, c])
            assert np.array_equal(cells[:, 2], vals[0, :kc, c])

    check_against_engine(eng, host[0])
    n_steps = 300
    moved = 0
    for it in range(n_steps):
        up, down = eng.step_hosted_ragged(host[it & 1], host[(it + 1) & 1], stage, out)
        assert eng.step_hosted_wait() == 0
        moved += up + down
        ref.step(1)
    last = host[n_steps & 1]
    check_against_engine(eng, last)
    ref_img = ref.pack_states_ragged().cpu()
    for (c0, c1, a), (_, _, b) in zip(eng.ragged_parts(last), ref.ragged_parts(ref_img)):
        assert torch.equal(a, b)
    assert torch.equal(eng.res[0], ref.res[0]) and torch.equal(eng.idx_cm[0], ref.idx_cm[0])
    assert torch.equal(eng.n_accepted, ref.n_accepted)
    kb = int(eng.gather_space(0)[0].max().item())
    assert moved / (2 * n_steps) < eng.state_image_bytes(kb)  # fewer bytes than the rows padded to the largest k

    # (b) the caller edits the value of the first cell of a chain of the LAST part
    parts = eng.ragged_parts(last)
    c0, c1, words = parts[-1]
    n_c = c1 - c0
    lc = n_c - 3
    pos = 4 * n_c + 3 * int(words[:lc].sum().item()) + 2
    words[pos:pos + 1] = (words[pos:pos + 1].view(torch.float64) * 1.03125).view(torch.int64)  # (a view: edits `last`)
    dst = host[(n_steps + 1) & 1]
    eng.step_hosted_ragged(last, dst, stage, out)
    assert eng.step_hosted_wait() == 1
    assert (eng.iteration.cpu().numpy() == 20 + n_steps + 1).all()
    check_against_engine(eng, dst)
    ref.step(1)
    got, want = decode(eng, dst), decode(ref, ref.pack_states_ragged().cpu())
    for c in range(n):
        if c != c0 + lc:
            assert got[c][0] == want[c][0] and np.array_equal(got[c][1], want[c][1]) and got[c][2:] == want[c][2:]
    eng.step(10)
    res_inc = eng.res[0].clone()
    eng.resync()
    np.testing.assert_allclose(eng.res[0].cpu().numpy(), res_inc.cpu().numpy(), rtol=0, atol=1e-9)
    # (c) a k outside [0, kmax] in the uploaded header is refused before anything is enqueued
    bad = dst.clone().pin_memory()
    bad[0] = K + 1
    with pytest.raises(Exception):
        eng.step_hosted_ragged(bad, host[n_steps & 1], stage, out)
    with pytest.raises(Exception):
        eng.step_hosted_ragged(dst, dst, stage, out)


@pytest.mark.parametrize("variant", ["one_tile_u8", "ray_tiles_u16"])
def test_hosted_step_mapped_records(monkeypatch, variant):
    """bbb_engine_step_hosted_mapped: the chains live in pinned host memory as fixed-stride records which the chain
    kernel reads and writes itself.  (a) The records decode to the chains' states; the fast path is bit-identical to
    resident stepping over 300 iterations (incl. a from-scratch re-evaluation) and never redone; only the cells in use
    cross the bus.  (b) Other engine calls between hosted calls (resident steps) are allowed.  (c) Edited chains are
    detected, exactly they are redone, every chain advances once.  (d) Bad arguments are refused.
    Variants: one ray tile with a 1-byte nearest-site index; several ray tiles with a 2-byte index (n_dimensions_max > 256)."""
    from bayesbay_b200 import synthetic

    kmax = 60
    if variant == "ray_tiles_u16":
        monkeypatch.setenv("BBB_CHAIN_TILE_RAYS", "96")
        kmax = 300
    ing = synthetic.tomography_2d(nx=32, ny=32, n_sources_per_side=5, n_receivers=12, kmin=4, kmax=kmax)
    spec = problems.spec_from_ingredients(ing)
    n = 148 + 37
    ref, eng = _engine(spec, n, seed=5), _engine(spec, n, seed=5)
    for e in (ref, eng):
        e.init_states()
        e.step(20)
    rw = eng.hosted_record_words()
    host = [torch.zeros((n, rw), dtype=torch.int64).pin_memory() for _ in range(2)]
    eng.pack_states_records(host[0])  # (the kernel writes the pinned tensor directly)
    torch.cuda.synchronize()
    assert torch.equal(host[0].cpu(), _records_in_use(eng, eng.pack_states_records().cpu(), host[0]))

    def check_against_engine(e, image):
        got = e.decode_records(image)
        k, sites, vals = (v.cpu().numpy() for v in e.gather_space(0))
        sig, temp, it = e.sigma[0][0].cpu().numpy(), e.temperature.cpu().numpy(), e.iteration.cpu().numpy()
        for c, d in enumerate(got):
            kc, cells, h = d["k"][0], d["cells"][0], d["header"]
            assert kc == k[c] and h[1:2].view(torch.float64).item() == sig[c] and h[2:3].view(torch.float64).item() == temp[c]
            assert int(h[3]) == it[c]
            assert np.array_equal(cells[:, 0], sites[0, :kc, c]) and np.array_equal(cells[:, 1], sites[1, :kc, c])
            assert np.array_equal(cells[:, 2], vals[0, :kc, c])

    check_against_engine(eng, host[0])
    assert eng.ray_tiles(0) == (1 if variant == "one_tile_u8" else 3) and eng.idx_cm[0].element_size() == (1 if kmax <= 256 else 2)
    n_steps = 300
    for it in range(n_steps):
        eng.step_hosted_mapped(host[it & 1], host[(it + 1) & 1])
        assert eng.step_hosted_mapped_wait() == 0
        ref.step(1)
    last = host[n_steps & 1]
    check_against_engine(eng, last)
    assert torch.equal(eng.res[0], ref.res[0]) and torch.equal(eng.idx_cm[0], ref.idx_cm[0])
    assert torch.equal(eng.n_accepted, ref.n_accepted) and torch.equal(eng.pack_states(kmax), ref.pack_states(kmax))
    up, down = eng.hosted_mapped_bytes()
    assert 0 < up < n_steps * n * rw * 8 and 0 < down < n_steps * n * rw * 8  # only the cells in use travel

    # (b) resident steps in between (they invalidate the proposal drawn ahead; the next hosted call draws it again)
    eng.step(3)
    ref.step(3)
    eng.pack_states_records(last)
    torch.cuda.synchronize()
    eng.step_hosted_mapped(last, host[(n_steps + 1) & 1])
    assert eng.step_hosted_mapped_wait() == 0
    ref.step(1)
    assert torch.equal(eng.pack_states(kmax), ref.pack_states(kmax)) and torch.equal(eng.res[0], ref.res[0])
    cur = host[(n_steps + 1) & 1]
    check_against_engine(eng, cur)

    # (c) the caller edits the value of the first cell of two chains
    base = 4  # header words: k, sigma, temperature, iteration
    edited = [5, n - 3]
    for c in edited:
        cur[c, base + 2:base + 3] = (cur[c, base + 2:base + 3].view(torch.float64) * 1.03125).view(torch.int64)
    nxt = host[n_steps & 1]
    eng.step_hosted_mapped(cur, nxt)
    assert eng.step_hosted_mapped_wait() == len(edited)
    assert (eng.iteration.cpu().numpy() == 20 + n_steps + 3 + 2).all()
    check_against_engine(eng, nxt)
    ref.step(1)
    got, want = eng.decode_records(nxt), ref.decode_records(ref.pack_states_records().cpu())
    for c in range(n):
        if c not in edited:
            assert got[c]["k"] == want[c]["k"] and np.array_equal(got[c]["cells"][0], want[c]["cells"][0])
            assert torch.equal(got[c]["header"], want[c]["header"])
    # the engine carries on from the edited states, caches consistent
    eng.step_hosted_mapped(nxt, cur)
    assert eng.step_hosted_mapped_wait() == 0
    eng.step(10)
    res_inc = eng.res[0].clone()
    eng.resync()
    np.testing.assert_allclose(eng.res[0].cpu().numpy(), res_inc.cpu().numpy(), rtol=0, atol=1e-9)
    # (d) refused: same buffer twice, pageable memory
    with pytest.raises(Exception):
        eng.step_hosted_mapped(cur, cur)
    with pytest.raises(Exception):
        eng.step_hosted_mapped(cur, torch.zeros((n, rw), dtype=torch.int64))


def test_hosted_step_mapped_with_several_spaces():
    """The mapped hosted step on a problem with two parameter spaces and two targets (ray tomography + a plain space
    read by a linear target): records hold both spaces' cells, moves of the plain space need no ray forward (decided in
    the block's short path), fast path bit-identical to resident stepping, an edit in the SECOND space is detected."""
    spec = problems.SPECS["tomo_joint_small"]()
    n = 96
    ref, eng = _engine(spec, n, seed=21), _engine(spec, n, seed=21)
    for e in (ref, eng):
        e.init_states()
        e.step(10)
    assert eng.chain_local and len(spec["spaces"]) == 2
    rw = eng.hosted_record_words()
    host = [torch.zeros((n, rw), dtype=torch.int64).pin_memory() for _ in range(2)]
    eng.pack_states_records(host[0])
    torch.cuda.synchronize()
    kb = [sp["kmax"] for sp in spec["spaces"]]
    n_steps = 120
    for it in range(n_steps):
        eng.step_hosted_mapped(host[it & 1], host[(it + 1) & 1])
        assert eng.step_hosted_mapped_wait() == 0
        ref.step(1)
    assert torch.equal(eng.pack_states(kb), ref.pack_states(kb)) and torch.equal(eng.res[0], ref.res[0])
    assert torch.equal(eng.n_accepted, ref.n_accepted) and torch.equal(eng.n_proposed, ref.n_proposed)
    last = host[n_steps & 1]
    got = eng.decode_records(last)
    for s_ in range(2):
        k, sites, vals = (None if v is None else v.cpu().numpy() for v in eng.gather_space(s_))
        nd = spec["spaces"][s_]["ndim"]
        for c in range(n):
            kc = got[c]["k"][s_]
            assert kc == k[c]
            cells = got[c]["cells"][s_]
            for d in range(nd):
                assert np.array_equal(cells[:, d], sites[d, :kc, c])
            for p_ in range(vals.shape[0]):
                assert np.array_equal(cells[:, nd + p_], vals[p_, :kc, c])
    # edit a value of the plain (second) space of one chain
    hrw = int(got[0]["header"].numel())
    base1 = hrw + spec["spaces"][0]["kmax"] * (spec["spaces"][0]["ndim"] + len(spec["spaces"][0]["params"]))
    c_edit = 17
    last[c_edit, base1:base1 + 1] = (last[c_edit, base1:base1 + 1].view(torch.float64) + 0.125).view(torch.int64)
    eng.step_hosted_mapped(last, host[(n_steps + 1) & 1])
    assert eng.step_hosted_mapped_wait() == 1
    assert (eng.iteration.cpu().numpy() == 10 + n_steps + 1).all()
    ref.step(1)
    a, b = eng.pack_states(kb).view(-1, n), ref.pack_states(kb).view(-1, n)
    keep = [c for c in range(n) if c != c_edit]
    assert torch.equal(a[:, keep], b[:, keep])


def _records_in_use(eng, records, like):
    """`records` with the words behind every chain's cells in use taken from `like` (a pack leaves them untouched)."""
    out = like.clone()
    rw = eng.hosted_record_words()
    k = records[:, 0]
    for c in range(records.shape[0]):
        n = 4 + 3 * int(k[c])
        out[c, :n] = records[c, :n]
    return out.view(-1, rw)


def test_k_bound_is_checked_on_the_device():
    """bbb_engine_set_k_bound refuses a bound below the largest k of the resident chains (it would let the ray kernel index
    outside its value table); bbb_engine_sync_k_bound measures that maximum on the device."""
    import ctypes as C

    from bayesbay_b200 import _lib as L

    spec = problems.spec_from_ingredients(problems.tomo_small())
    eng = _engine(spec, 64, seed=3)
    eng.init_states()
    eng.step(40)
    k_true = int(eng.gather_space(0)[0].max().item())
    out = (C.c_int32 * len(spec["spaces"]))()
    L.check(eng.lib.bbb_engine_sync_k_bound(eng._handle, out, eng._stream))
    assert out[0] == k_true
    L.check(eng.lib.bbb_engine_set_k_bound(eng._handle, 0, k_true))
    with pytest.raises(ValueError, match="below the largest k"):
        L.check(eng.lib.bbb_engine_set_k_bound(eng._handle, 0, k_true - 1))
    eng.step(5)  # (the engine goes on with the bound it has)


def test_exception_counters():
    """What the reference reports under statistics["exceptions"] (src/bayesbay/_markov_chain.py:205-247) has a device
    analogue: proposals with a non-finite likelihood ratio are rejected and counted; rejection-sampling loops that give
    up after 100 000 tries (the reference would loop forever) keep the old value and are counted."""
    from bayesbay_b200 import synthetic
    from bayesbay_b200._batch import ChainBatch

    # (a) a NaN datum: the misfit of every state is NaN, so every proposal has a NaN likelihood ratio
    ing = synthetic.polyfit()
    ing["d_obs"] = ing["d_obs"].copy()
    ing["d_obs"][3] = np.nan
    spec = problems.spec_from_ingredients(ing)
    batch = ChainBatch(spec, 32, seed=5, save_dpred=False)
    before = batch.current_states()
    batch.advance(200, 0, 100)
    prop = batch.engine.n_proposed.cpu().numpy()  # [move][C]
    exc = batch.engine.n_exceptions.cpu().numpy()
    assert (exc[1] == 0).all() and prop.sum(0).tolist() == [200] * 32
    np.testing.assert_array_equal(exc[0], 200)
    assert batch.engine.n_accepted.sum().item() == 0
    after = batch.current_states()
    for b, a in zip(before, after):  # ... and rejected: the coefficients never moved
        for name in ("m0", "m1", "m2", "m3"):
            assert a["my_param_space"][name] == b["my_param_space"][name]
    st = batch.statistics()[0]
    assert st["exceptions"]["NonFiniteLikelihoodRatio"] == exc[0, 0] and "ResamplingLimitReached" not in st["exceptions"]

    # (b) a noise range a thousand million times narrower than the proposal: the resampling loop gives up
    ing = synthetic.polyfit()
    ing["target"] = dict(ing["target"], std_min=10.0, std_max=10.0 + 1e-9, std_perturb_std=4.0)
    spec = problems.spec_from_ingredients(ing)
    batch = ChainBatch(spec, 8, seed=6, save_dpred=False)
    sig0 = batch.engine.sigma[0][0].clone()
    batch.advance(60, 0, 100)
    prop = batch.engine.n_proposed.cpu().numpy()
    exc = batch.engine.n_exceptions.cpu().numpy()
    assert (exc[1] > 0).any()
    assert (exc[1] <= prop[4]).all() and (exc[1] >= prop[4] - 1).all()  # (one in ~1e4 loops may find the sliver)
    assert (exc[0] == 0).all()
    moved = (batch.engine.sigma[0][0] != sig0).cpu().numpy()
    assert moved.sum() <= 1
    assert batch.statistics()[int(np.argmax(exc[1]))]["exceptions"]["ResamplingLimitReached"] == exc[1].max()
