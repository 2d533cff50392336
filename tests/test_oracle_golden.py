"""The oracle against fixtures recorded from the UNMODIFIED reference (tests/golden/make_golden.py).

This is what pins the oracle: every deterministic kernel and every step of tape-driven reference
chains must be reproduced.  CPU only."""
import math
import os

import numpy as np
import pytest

from oracle import chain as OC
from oracle import kernels as OK
from oracle import rng as OR
from tests import problems

RTOL = 1e-12


@pytest.fixture(scope="module")
def kg(golden_dir):
    return np.load(os.path.join(golden_dir, "kernels.npz"))


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert OR.philox4x32((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert OR.philox4x32((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert OR.philox4x32((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_raster_2d_matches_kdtree(kg):
    pts = kg["r2_points"]
    for i, k in enumerate(kg["r2_ks"]):
        idx = OK.nearest_site_2d(kg[f"r2_sites_{i}"], pts)
        assert np.array_equal(idx, kg[f"r2_idx_{i}"]), f"k={k}"


def test_raster_1d(kg):
    x = kg["r1_x"]
    for i, k in enumerate(kg["r1_ks"]):
        sites, vals = kg[f"r1_sites_{i}"], kg[f"r1_vals_{i}"]
        assert np.array_equal(OK.lookup1d_forward(x, sites, vals), kg[f"r1_field_{i}"])
        nn = np.array([OK.nearest_neighbour_1d(float(v), sites) for v in x])
        assert np.array_equal(nn, kg[f"r1_nn_{i}"])
    sites, xt = kg["r1_tie_sites"], kg["r1_tie_x"]
    assert np.array_equal([OK.nearest_neighbour_1d(float(v), sites) for v in xt], kg["r1_tie_nn"])
    assert np.array_equal(OK.lookup1d_forward(xt, sites, np.arange(4.0)), kg["r1_tie_field"])


def test_cython_helpers(kg):
    sites = kg["r1_tie_sites"]
    # docstring vectors of Voronoi1D.compute_cell_extents (discretization/_voronoi.py:657-666)
    assert np.array_equal(OK.cell_extents_1d(sites, 0.0, 15.0, np.nan), [3.75, 3.0, 2.25, 6.0])
    assert np.array_equal(OK.cell_extents_1d(sites, 0.0, 15.0, np.nan), kg["ext_c"])
    a = OK.cell_extents_1d(sites, 0.0, math.inf, np.nan)
    assert np.array_equal(a[:3], [3.75, 3.0, 2.25]) and np.isnan(a[3])
    assert np.array_equal(a[:3], kg["ext_a"][:3])  # quirk Q9: the -ffast-math build returns inf, not nan
    assert np.array_equal(OK.insert_1d(np.arange(5.0), 2, 9.5), kg["ins"])
    assert np.array_equal(OK.delete_1d(np.arange(5.0), 2), kg["del"])
    np.testing.assert_allclose(OK.inverse_covariance(0.7, 0.3, 6), kg["invcov"], rtol=1e-15)


def test_log_prior(kg):
    uni = dict(kind="uniform", a=-1.0, b=1.0)
    gau = dict(kind="gaussian", a=0.0, b=1.0)
    # tests/20_test_param_space_log_prior.py:16-30 (exact equality)
    lp = OK.space_log_prior([uni, gau], [np.array([0.0]), np.array([0.0])], 1)
    assert lp == math.log(1 / 2) + math.log(1 / math.sqrt(2 * math.pi)) == float(kg["lp_test20"])
    ps = [dict(kind="uniform", a=2.0, b=4.0), dict(kind="gaussian", a=-3.0, b=2.0), dict(kind="laplace", a=0.5, b=2.0)]
    vals = kg["lp_vals"]
    each = np.array([[OK.prior_log_pdf(p, float(v)) for v in row] for p, row in zip(ps, vals)])
    np.testing.assert_allclose(each, kg["lp_each"], rtol=1e-15)
    np.testing.assert_allclose(OK.space_log_prior(ps, list(vals), 7), float(kg["lp_joint"]), rtol=1e-14)
    # quickstart: p(5)=p(10)=0.1, p(20)=0 (tests/22_log_like_func.py:24-31)
    q = dict(kind="uniform", a=5.0, b=15.0)
    got = [math.exp(OK.prior_log_pdf(q, v)) for v in (5.0, 10.0, 20.0)]
    np.testing.assert_allclose(got, kg["lp_quickstart"], rtol=1e-15)
    np.testing.assert_allclose(got, [0.1, 0.1, 0.0], rtol=1e-15)


def test_misfit_logdet_ratio(kg):
    dobs = kg["ll_dobs"]
    mo = OK.misfit_and_logdet(kg["ll_dp_old"], dobs, std=0.21)
    mn = OK.misfit_and_logdet(kg["ll_dp_new"], dobs, std=0.19)
    np.testing.assert_allclose(mo, kg["ll_old"], rtol=RTOL)
    np.testing.assert_allclose(mn, kg["ll_new"], rtol=RTOL)
    for T, key in ((1.0, "ll_ratio_T1"), (3.0, "ll_ratio_T3")):
        np.testing.assert_allclose(OK.log_like_ratio(mo[0], mo[1], mn[0], mn[1], T), float(kg[key]), rtol=RTOL)
    np.testing.assert_allclose(OK.misfit_and_logdet(kg["ll_dp_old"], dobs, std=0.23, corr=0.37), kg["ll_corr"], rtol=RTOL)
    np.testing.assert_allclose(OK.misfit_and_logdet(kg["ll_dp_old"], dobs, cov_inv=4.0), kg["ll_fixed_s"], rtol=RTOL)
    np.testing.assert_allclose(OK.misfit_and_logdet(kg["ll_dp_old"], dobs, cov_inv=kg["ll_w"]), kg["ll_fixed_v"], rtol=RTOL)


def test_full_matrix_inverse_covariance(golden_dir):
    """`covariance_mat_inv @ residual` for a 2-D inverse covariance (likelihood/_target.py:150-151) against values
    recorded from the unmodified reference (tests/golden/make_cov_matrix.py)."""
    g = np.load(os.path.join(golden_dir, "cov_matrix.npz"))
    for case in range(int(g["n_cases"])):
        cov, dobs, dp = g[f"c{case}_cov_inv"], g[f"c{case}_dobs"], g[f"c{case}_dpred"]
        got = [OK.misfit_and_logdet(dp[i], dobs, cov_inv=cov) for i in range(2)]
        np.testing.assert_allclose([v[0] for v in got], g[f"c{case}_misfit"], rtol=RTOL)
        np.testing.assert_array_equal([v[1] for v in got], g[f"c{case}_logdet"])
        np.testing.assert_allclose(OK.log_like_ratio(got[0][0], got[0][1], got[1][0], got[1][1], 1.7),
                                   float(g[f"c{case}_ratio_T1p7"]), rtol=1e-11, atol=1e-13)


def test_tomography_forward(kg):
    spec = problems.spec_from_ingredients(problems.tomo_small())
    dpred, _ = OK.tomography_forward(spec["targets"][0]["forward"], kg["fw_sites"], kg["fw_vel"])
    np.testing.assert_allclose(dpred, kg["fw_dpred"], rtol=RTOL)


def _state_from_fixture(spec, g, c, t):
    sp = spec["spaces"][0]
    k = int(g["k"][c, t])
    ndim = sp["ndim"]
    sites = None
    if ndim == 2:
        sites = g["sites"][c, t, :k, :].copy()
    elif ndim == 1:
        sites = g["sites"][c, t, :k, 0].copy()
    vals = [g["values"][c, t, j, :k].copy() for j in range(len(sp["params"]))]
    corr = float(g["corr"][c, t]) if "corr" in g.files else float("nan")
    sig = float(g["sigma"][c, t])
    st = OC.ChainState([OC.SpaceState(k, sites, vals)], [None if math.isnan(sig) else sig], 1.0, [None],
                       [None if math.isnan(corr) else corr])
    OC.evaluate(spec, st)
    return st


@pytest.mark.parametrize("name", list(problems.REPLAYS))
def test_step_replay(name, golden_dir):
    """Feed the recorded tape to the oracle: every step must reproduce the reference's move choice,
    log_prob_ratio, log-likelihood ratio, accept decision, state and tape consumption."""
    g = np.load(os.path.join(golden_dir, f"replay_{name}.npz"))
    spec = problems.spec_from_ingredients(problems.REPLAYS[name]())
    n_chains, n_steps = g["lpr"].shape
    for c in range(n_chains):
        # initial state: replay the initialisation draws too
        s0 = OR.TapeStream(g["tape"][c])
        st = OC.init_state(spec, s0)
        assert s0.n == g["cursor"][c, 0]
        ref0 = _state_from_fixture(spec, g, c, 0)
        _assert_state_close(st, ref0)
        for t in range(n_steps):
            s = OR.TapeStream(g["tape"][c], int(g["cursor"][c, t]))
            st, info = OC.step(spec, st, s)
            assert info["move"] == g["move"][c, t], (c, t)
            assert info["n_draws"] == g["cursor"][c, t + 1], (c, t)
            ref_lpr = g["lpr"][c, t]
            if math.isinf(ref_lpr):
                assert info["log_prob_ratio"] == ref_lpr
            else:
                np.testing.assert_allclose(info["log_prob_ratio"], ref_lpr, rtol=1e-11, atol=1e-12)
                np.testing.assert_allclose(info["log_like_ratio"], g["llr"][c, t], rtol=1e-9, atol=1e-9)
            assert info["accepted"] == bool(g["accepted"][c, t]), (c, t)
            _assert_state_close(st, _state_from_fixture(spec, g, c, t + 1))


def _assert_state_close(a, b):
    sa, sb = a.spaces[0], b.spaces[0]
    assert sa.k == sb.k
    if sa.sites is not None:
        np.testing.assert_allclose(sa.sites, sb.sites, rtol=RTOL, atol=1e-15)
    for va, vb in zip(sa.values, sb.values):
        np.testing.assert_allclose(va, vb, rtol=RTOL, atol=1e-15)
    assert (a.sigma[0] is None) == (b.sigma[0] is None)
    if a.sigma[0] is not None:
        np.testing.assert_allclose(a.sigma[0], b.sigma[0], rtol=RTOL)
    assert (a.corr[0] is None) == (b.corr[0] is None)
    if a.corr[0] is not None:
        np.testing.assert_allclose(a.corr[0], b.corr[0], rtol=RTOL)


def test_pt_swap_rounds(golden_dir):
    g = np.load(os.path.join(golden_dir, "pt_swap.npz"))
    np.testing.assert_allclose(OC.pt_ladder(6, 5, 0.4), g["ladder"], rtol=1e-15)
    for r in range(g["misfit"].shape[0]):
        s = OR.TapeStream(g["tape"], int(g["cursor0"][r]))
        temps, _ = OC.pt_swap_round(g["misfit"][r], g["logdet"][r], list(g["t_before"][r]), s)
        assert s.n == g["cursor1"][r]
        assert np.array_equal(temps, g["t_after"][r])
    lad = np.load(os.path.join(golden_dir, "pt_ladder.npz"))
    for key in lad.files:
        np.testing.assert_allclose(OC.pt_ladder(int(key[1:]), 7.5, 0.4), lad[key], rtol=1e-15)
