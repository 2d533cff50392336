"""The forward operator descriptors as reference forwards, and the unmodified reference from baseline/_ref
(test infrastructure: oracle/reference_runner.py).  No GPU."""
import numpy as np
import pytest

from bayesbay_b200 import synthetic
from bayesbay_b200._state import ParameterSpaceState, State
from oracle import kernels as OK
from oracle import reference_runner as RR


def test_host_forwards_match_the_oracle():
    rng = np.random.default_rng(3)
    # 2-D ray tomography: KD-tree + CSR on the host == the oracle's brute-force restatement
    ing = synthetic.tomography_2d(nx=12, ny=10, n_sources_per_side=3, n_receivers=5, kmin=4, kmax=20)
    op = synthetic.forward_operator(ing)
    for _ in range(5):
        k = int(rng.integers(4, 20))
        sites = rng.uniform(-1, 1, (k, 2))
        vel = rng.uniform(2, 4, k)
        st = State({"voronoi": ParameterSpaceState(k, {"discretization": sites, "vel": vel})})
        idx = OK.nearest_site_2d(sites, ing["points"])
        import scipy.sparse

        jac = scipy.sparse.csr_matrix((ing["data"], ing["indices"], ing["indptr"]), shape=(op.n_data, ing["points"].shape[0]))
        np.testing.assert_allclose(op(st), jac @ (1 / vel[idx]), rtol=1e-13)
    # 1-D lookup incl. exact ties and points outside the sites
    ing1 = synthetic.partition_1d()
    op1 = synthetic.forward_operator(ing1)
    for _ in range(20):
        k = int(rng.integers(1, 12))
        sites = np.sort(rng.choice(np.linspace(0, 10, 41), k, replace=False))  # data positions hit sites and midpoints
        y = rng.normal(size=k)
        st = State({"voronoi": ParameterSpaceState(k, {"discretization": sites, "y": y})})
        np.testing.assert_array_equal(op1(st), OK.lookup1d_forward(ing1["x"], sites, y))
    ing2 = synthetic.polyfit()
    m = rng.normal(size=4)
    st = State({"my_param_space": ParameterSpaceState(1, {f"m{i}": np.array([m[i]]) for i in range(4)})})
    np.testing.assert_allclose(synthetic.forward_operator(ing2)(st), ing2["G"] @ m, rtol=1e-14)
    ing3 = synthetic.transd_polyfit()
    c = rng.normal(size=5)
    st = State({"my_param_space": ParameterSpaceState(5, {"coefficients": c})})
    np.testing.assert_allclose(synthetic.forward_operator(ing3)(st), np.vander(ing3["x"], 5, True) @ c, rtol=1e-14)


@pytest.mark.skipif(not RR.available(), reason="baseline/_ref is not installed (run __graft_entry__.build())")
def test_unmodified_reference_runs_the_descriptor_problem():
    """The same problem object -- built with the reference's own classes, the descriptor as its forward function --
    runs in the unmodified reference, through its own chain pool."""
    ing = synthetic.tomography_2d(nx=12, ny=10, n_sources_per_side=3, n_receivers=5, kmin=4, kmax=20)
    rate, wall, inv = RR.run_pool(ing, n_chains=2, n_iterations=40, n_jobs=2, save_every=10)
    res = inv.get_results()
    assert len(res["voronoi.n_dimensions"]) == 2 * 4
    dpred_keys = [k for k in res if k.endswith(".dpred")]
    assert len(dpred_keys) == 1 and len(res[dpred_keys[0]][0]) == ing["d_obs"].size
    assert all(ch.statistics["n_proposed_models_total"] == 40 for ch in inv.chains)
    assert rate > 0


@pytest.mark.skipif(not RR.available(), reason="baseline/_ref is not installed (run __graft_entry__.build())")
def test_interpolate_tessellation_extents_matches_the_reference():
    """Voronoi1D.interpolate_tessellation(..., input_type='extents') (discretization/_voronoi.py:705-732,
    interpolate_depth_profile in _utils_1d.pyx:119-140) incl. its sweep semantics."""
    import bayesbay_b200 as bb

    ref = RR.import_reference()
    rng = np.random.default_rng(0)
    for _ in range(20):
        k = int(rng.integers(1, 9))
        ext = rng.uniform(0.2, 3.0, k)
        vals = rng.normal(size=k)
        x0 = np.sort(rng.uniform(0.0, ext.sum() * 1.3, 40))
        x0[rng.integers(0, 40, 3)] = np.cumsum(ext)[rng.integers(0, k, 3)]  # exactly on interfaces (unsorted on purpose)
        want = ref.discretization.Voronoi1D.interpolate_tessellation(ext, vals, x0, input_type="extents")
        got = bb.discretization.Voronoi1D.interpolate_tessellation(ext, vals, x0, input_type="extents")
        np.testing.assert_array_equal(got, want)
    x0 = np.array([-0.5, 0.1, 0.7])  # a position below zero sends the sweep to the last layer for good
    ext, vals = np.array([1.0, 1.0, 1.0]), np.array([1.0, 2.0, 3.0])
    np.testing.assert_array_equal(bb.discretization.Voronoi1D.interpolate_tessellation(ext, vals, x0, input_type="extents"),
                                  ref.discretization.Voronoi1D.interpolate_tessellation(ext, vals, x0, input_type="extents"))
