"""Host-side mirror of the reference API: names, signatures, perturbation structure, the problem compiler
and the error behaviour for things that cannot run on the device.  No GPU needed."""
import inspect
import os
import sys

import numpy as np
import pytest

import bayesbay_b200 as bb
from bayesbay_b200 import _compile, synthetic
from bayesbay_b200.exceptions import DeviceUnsupported
from bayesbay_b200.forward import LinearForward, NearestLookup1D, RayTomographyForward
from tests import problems


def device_forward(ing):
    if ing["kind"] == "transd_polyfit":
        from bayesbay_b200.forward import PolynomialForward

        return PolynomialForward(ing["x"], "my_param_space", "coefficients")
    if ing["kind"] == "tomography_2d":
        return RayTomographyForward((ing["indptr"], ing["indices"], ing["data"]), ing["points"], "voronoi", "vel")
    if ing["kind"] == "partition_1d":
        return NearestLookup1D(ing["x"], "voronoi", "y")
    return LinearForward(ing["G"], "my_param_space", ["m0", "m1", "m2", "m3"])


def _same(a, b, path=""):
    if isinstance(a, dict):
        assert a.keys() == b.keys(), (path, a.keys(), b.keys())
        for k in a:
            _same(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}/{i}")
    elif isinstance(a, np.ndarray):
        assert np.array_equal(a, b), path
    else:
        assert a == b, (path, a, b)


def test_public_names_of_the_reference_exist():
    # src/bayesbay/__init__.py:3-55 and the sub-packages' __all__
    for name in ("BayesianInversion", "BaseBayesianInversion", "MarkovChain", "BaseMarkovChain", "State",
                 "ParameterSpaceState", "DataNoiseState", "Target", "LogLikelihood"):
        assert hasattr(bb, name)
    for mod, names in {
        "prior": ("Prior", "UniformPrior", "GaussianPrior", "LaplacePrior", "CustomPrior", "UniformParameter",
                  "GaussianParameter", "CustomParameter"),
        "likelihood": ("Target", "LogLikelihood"),
        "parameterization": ("Parameterization", "ParameterSpace"),
        "discretization": ("Discretization", "Voronoi", "Voronoi1D", "Voronoi2D"),
        "perturbations": ("Perturbation", "ParamPerturbation", "BirthPerturbation", "DeathPerturbation",
                          "NoisePerturbation", "ParamSpacePerturbation"),
        "samplers": ("Sampler", "VanillaSampler", "ParallelTempering", "SimulatedAnnealing"),
    }.items():
        for n in names:
            assert hasattr(getattr(bb, mod), n), (mod, n)


def test_signatures_match_the_reference():
    def params(f):
        return [p for p in inspect.signature(f).parameters if p != "self"]

    assert params(bb.BayesianInversion.__init__)[:5] == ["parameterization", "log_likelihood", "n_chains",
                                                         "walkers_starting_states", "save_dpred"]
    assert params(bb.BayesianInversion.run) == ["sampler", "n_iterations", "burnin_iterations", "save_every", "verbose",
                                                "print_every", "parallel_config"]
    assert params(bb.BayesianInversion.get_results) == ["keys", "concatenate_chains"]
    assert params(bb.prior.UniformPrior.__init__) == ["name", "vmin", "vmax", "perturb_std", "perturb_std_birth", "position"]
    assert params(bb.prior.GaussianPrior.__init__) == ["name", "mean", "std", "perturb_std", "perturb_std_birth", "position"]
    assert params(bb.discretization.Voronoi2D.__init__) == [
        "name", "vmin", "vmax", "polygon", "perturb_std", "n_dimensions", "n_dimensions_min", "n_dimensions_max",
        "n_dimensions_init_range", "parameters", "birth_from", "compute_kdtree"]
    assert params(bb.discretization.Voronoi1D.__init__) == [
        "name", "vmin", "vmax", "perturb_std", "n_dimensions", "n_dimensions_min", "n_dimensions_max",
        "n_dimensions_init_range", "parameters", "birth_from"]
    assert params(bb.parameterization.ParameterSpace.__init__) == [
        "name", "n_dimensions", "n_dimensions_min", "n_dimensions_max", "n_dimensions_init_range", "parameters"]
    assert params(bb.likelihood.Target.__init__) == [
        "name", "dobs", "covariance_mat_inv", "noise_is_correlated", "std_min", "std_max", "std_perturb_std",
        "correlation_min", "correlation_max", "correlation_perturb_std"]
    assert params(bb.samplers.ParallelTempering.__init__)[:3] == ["temperature_max", "chains_with_unit_temperature", "swap_every"]
    assert params(bb.samplers.SimulatedAnnealing.__init__) == ["temperature_start", "cooling_fraction"]
    d = {p.name: p.default for p in inspect.signature(bb.samplers.ParallelTempering.__init__).parameters.values()}
    assert (d["temperature_max"], d["chains_with_unit_temperature"], d["swap_every"]) == (5, 0.4, 500)


def test_perturbation_structure_and_names():
    """[ParamSpacePerturbation(voronoi), NoisePerturbation(d_obs)] weights [6, 1]; inner [Birth, Death,
    ParamPerturbation(vel), ParamPerturbation(discretization)] weights [1, 1, 3, 1]."""
    par, ll = synthetic.build_problem(problems.tomo_small(), bb, device_forward)
    funcs = par.perturbation_funcs + ll.perturbation_funcs
    weights = par.perturbation_weights + ll.perturbation_weights
    assert [f.__name__ for f in funcs] == ["ParamSpacePerturbation(param_space_name=voronoi)", "NoisePerturbation(d_obs)"]
    assert weights == [6, 1]
    inner = funcs[0]
    assert [f.__name__ for f in inner.perturbation_funcs] == [
        "BirthPerturbation(voronoi)", "DeathPerturbation(voronoi)", "ParamPerturbation(voronoi.vel)",
        "ParamPerturbation(voronoi.discretization)"]
    assert inner.perturbation_weights == [1, 1, 3, 1]
    par, ll = synthetic.build_problem(synthetic.polyfit(), bb, device_forward)
    assert par.perturbation_weights + ll.perturbation_weights == [3, 1]
    assert par.perturbation_funcs[0].perturbation_funcs[0].__name__ == (
        "ParamPerturbation(['my_param_space.m0', 'my_param_space.m1', 'my_param_space.m2', 'my_param_space.m3'])")


@pytest.mark.parametrize("name", list(problems.REPLAYS))
def test_compiler_emits_the_spec_the_oracle_runs(name):
    ing = problems.REPLAYS[name]()
    par, ll = synthetic.build_problem(ing, bb, device_forward)
    spec = _compile.compile_problem(par, ll)
    _same(spec, problems.spec_from_ingredients(ing))
    names = _compile.perturbation_names(spec)
    assert len(names) == 5
    assert (names[4] is None) if not spec["targets"][0]["hierarchical"] else names[4][0].startswith("NoisePerturbation(")


def test_compiler_accepts_objects_of_the_unmodified_reference():
    """Duck typing: a Parameterization / Target built with the reference package compiles to the same spec
    (this is what lets the batched sampler plug into the reference's own run(sampler=...) seam)."""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    if not os.path.isdir("/root/reference"):
        pytest.skip("the reference tree is not available on this machine")
    import make_golden

    ref = make_golden.import_reference()
    for name in ("tomo_small", "partition_1d", "polyfit_nb", "partition_1d_corr", "partition_1d_posdep",
                 "partition_1d_fixed", "transd_polyfit"):
        ing = problems.REPLAYS[name]()
        par, ll = synthetic.build_problem(ing, ref, lambda i: (lambda state: None))
        # the reference wraps forwards; swap in the device operator description for compilation
        ll._fwd_functions = [device_forward(ing)]
        _same(_compile.compile_problem(par, ll), problems.spec_from_ingredients(ing))


def test_things_that_cannot_run_on_the_device_raise():
    t = bb.likelihood.Target("d", np.zeros(3), std_min=0, std_max=1, std_perturb_std=0.1)
    with pytest.raises(TypeError, match="no CPU fallback"):
        bb.likelihood.LogLikelihood(t, lambda state: np.zeros(3))
    with pytest.raises(TypeError):
        bb.likelihood.LogLikelihood(log_like_func=lambda s: 0.0)
    with pytest.raises(TypeError):
        bb.BaseBayesianInversion([None], [lambda s: (s, 0)], n_chains=1)
    with pytest.raises(ValueError):
        bb.likelihood.Target("d", np.zeros(3), noise_is_correlated=True, covariance_mat_inv=2.0)
    # scalar, diagonal and full inverse covariances are all accepted (likelihood/_target.py:146-152)
    assert not bb.likelihood.Target("d", np.zeros(3), covariance_mat_inv=np.eye(3)).is_hierarchical
    with pytest.raises(ValueError):
        bb.likelihood.Target("d", np.zeros(3), covariance_mat_inv=np.zeros((3, 3, 3)))
    with pytest.raises(DeviceUnsupported):  # N-D position grids (LinearNDInterpolator) are not on the device
        bb.prior.UniformPrior("v", vmin=[0, 1], vmax=[1, 2], perturb_std=0.1, position=[[0, 0], [1, 1]])
    pd = bb.prior.UniformPrior("v", vmin=[0, 1], vmax=[1, 2], perturb_std=0.1, position=[0, 1])
    with pytest.raises(DeviceUnsupported):  # position-dependent priors need sites: Voronoi1D only
        _compile.compile_problem(bb.parameterization.Parameterization(
            bb.parameterization.ParameterSpace("flat", n_dimensions=2, parameters=[pd])))
    with pytest.raises(DeviceUnsupported):
        bb.discretization.Voronoi2D("v", polygon=[(0, 0), (1, 0), (1, 1)])
    cp = bb.prior.CustomPrior("c", lambda v: 0.0, lambda: 0.0, perturb_std=1.0)
    ps = bb.parameterization.ParameterSpace("ps", n_dimensions=1, parameters=[cp])
    with pytest.raises(DeviceUnsupported):
        _compile.compile_problem(bb.parameterization.Parameterization(ps))
    inner = bb.parameterization.ParameterSpace("inner", n_dimensions=1)
    with pytest.raises(DeviceUnsupported):
        bb.parameterization.ParameterSpace("outer", n_dimensions=1, parameters=[inner])
    with pytest.raises(ValueError):
        bb.discretization.Voronoi("v", 2, 0, 1, 0.1)
    with pytest.raises(ValueError, match="duplicate"):
        bb.parameterization.Parameterization([ps, ps])
    with pytest.raises(AssertionError):
        bb.likelihood.Target("d", np.zeros(3), std_min=2, std_max=1)


def test_state_containers():
    ps = bb.ParameterSpaceState(3, {"discretization": np.array([1.0, 2.0, 3.0]), "v": np.array([4.0, 5.0, 6.0])})
    st = bb.State({"space": ps, "data": bb.DataNoiseState(std=0.5, correlation=None)}, temperature=2.0)
    assert dict(st.items()).keys() == {"space.n_dimensions", "space.discretization", "space.v", "data.std"}
    c = st.copy()
    c["space"]["v"][0] = -1
    assert st["space"]["v"][0] == 4.0 and c.temperature == 2.0
    assert st["space"][[0, 2]].n_dimensions == 2
    with pytest.raises(ValueError):
        bb.ParameterSpaceState(2, {"v": np.zeros(3)})
    with pytest.raises(TypeError):
        bb.ParameterSpaceState(2.5, {})
    with pytest.raises(TypeError):
        bb.State({"x": np.zeros(2)})
    assert np.array_equal(bb.discretization.Voronoi1D.compute_cell_extents(np.array([2, 5.5, 8, 10]), lb=0, ub=15),
                          [3.75, 3.0, 2.25, 6.0])  # docstring vector, discretization/_voronoi.py:657-666
    a = bb.discretization.Voronoi1D.compute_cell_extents(np.array([2, 5.5, 8, 10]), lb=None, ub=None, fill_value=np.nan)
    assert np.isnan(a[0]) and np.isnan(a[3]) and a[1] == 3.0 and a[2] == 2.25


def test_synthetic_ray_matrix_is_exact():
    ing = problems.tomo_small()
    indptr, indices, data = ing["indptr"], ing["indices"], ing["data"]
    rows = np.add.reduceat(data, indptr[:-1])
    np.testing.assert_allclose(rows, 1.0, rtol=1e-13)  # G_ij = l_j / L_i
    assert (np.diff(indptr) > 0).all() and (data > 0).all()
    # un-normalised lengths sum to the ray length
    src = np.array([[-1.0, -0.3]])
    rcv = np.array([[0.7, 0.9]])
    ip, ix, dt = synthetic.straight_ray_csr(src, rcv, 12, 10, normalise=False)
    np.testing.assert_allclose(dt.sum(), np.hypot(1.7, 1.2), rtol=1e-13)


def test_save_schedule_follows_the_reference_rule():
    """Launch splitting vs a literal replay of BaseMarkovChain.advance_chain (_markov_chain.py:284-289)."""
    from bayesbay_b200._batch import save_schedule

    rng = np.random.default_rng(0)
    for _ in range(300):
        it0, n, burnin, every = int(rng.integers(0, 50)), int(rng.integers(0, 120)), int(rng.integers(0, 80)), int(rng.integers(1, 25))
        want = []
        i = it0
        for _ in range(n):
            i += 1
            if i > burnin and (i - burnin) % every == 0:
                want.append(i)
        got, i = [], it0
        sched = save_schedule(it0, n, burnin, every)
        for steps, save in sched:
            assert steps >= 1
            i += steps
            if save:
                got.append(i)
        assert i == it0 + n and got == want
        assert all(save for _, save in sched[:-1])  # only the last launch may stop short of a save point
    # ParallelTempering.run quirk Q3: burn-in 1500 handed over window-relative (500) -> T=1 chains save from 600 on
    saves, it = [], 0
    for n_it, burnin_it in ((500, 500), (500, 500), (500, 500), (500, 0), (0, 0)):
        for steps, save in save_schedule(it, n_it, burnin_it, 100):
            it += steps
            if save:
                saves.append(it)
    assert saves[0] == 600 and saves[-1] == 2000 and len(saves) == 15
    with pytest.raises(ValueError):
        save_schedule(0, 10, 0, 0)


def test_position_dependent_prior_host_api():
    """Hyper-parameters interpolated at a position (reference prior/_prior.py:257-297, _utils_1d.pyx:68-84)."""
    import math

    p = bb.prior.UniformPrior("y", vmin=[-35.0, -20.0, -35.0], vmax=[35.0, 40.0, 10.0], perturb_std=[3.5, 2.0, 4.0],
                              position=[0.0, 5.0, 10.0])
    assert p.get_vmin_vmax(2.5) == (-27.5, 37.5) and p.get_vmin_vmax(-1) == (-35.0, 35.0) and p.get_vmin_vmax(99) == (-35.0, 10.0)
    assert p.get_perturb_std(7.5) == 3.0 and p.get_perturb_std_birth(7.5) == 3.0
    assert p.log_prior(0.0, 2.5) == -math.log(65.0) and p.log_prior(39.0, 2.5) == -math.inf
    with pytest.raises(ValueError):
        p.get_vmin_vmax(None)
    g = bb.prior.GaussianPrior("g", mean=[0.0, 10.0], std=2.0, perturb_std=0.5, position=[0.0, 1.0])
    assert g.get_mean(0.25) == 2.5 and g.get_std(0.25) == 2.0
    spec = _compile.compile_prior(p)
    assert spec["table"]["a"].tolist() == [-35.0, -20.0, -35.0] and spec["table"]["perturb_std_birth"].tolist() == [3.5, 2.0, 4.0]
    ing = problems.REPLAYS["partition_1d_posdep"]()
    par, ll = synthetic.build_problem(ing, bb, device_forward)
    _same(_compile.compile_problem(par, ll), problems.spec_from_ingredients(ing))


def test_ray_matrix_csc_levels_reproduce_the_matrix():
    """The ray-tile-blocked, merged-column CSC copy the chain-local step reads (host-side construction, no device
    needed): level l must equal the matrix with every aligned group of 4**l columns summed."""
    import scipy.sparse as sp

    pytest.importorskip("torch")
    from bayesbay_b200._engine import morton_order, ray_matrix_csc_levels

    A = sp.random(70, 43, 0.25, format="csr", random_state=3)
    dense = A.toarray()
    for tile in (70, 32, 16):
        n_tiles = -(-70 // tile)
        for lvl, (colptr, rows, dat, bound) in enumerate(ray_matrix_csc_levels(A.indptr, A.indices, A.data, 43, tile)):
            s = 4 ** lvl
            n_g = -(-43 // s)
            assert colptr.shape == (n_tiles, n_g + 1) and colptr.dtype == np.int32 and rows.dtype == np.uint16
            M = np.zeros((70, n_g))
            for q in range(n_tiles):
                for g in range(n_g):
                    seg = slice(colptr[q, g], colptr[q, g + 1])
                    assert (np.diff(rows[seg].astype(int)) > 0).all()  # rows ascending inside a column
                    M[q * tile + rows[seg], g] += dat[seg]
            want = np.add.reduceat(dense, np.arange(0, 43, s), axis=1)
            np.testing.assert_allclose(M, want, rtol=1e-15, atol=0)
            assert bound == pytest.approx(np.abs(dense).sum(1).max())
    # Morton order: aligned groups of 4 / 16 consecutive points are 2x2 / 4x4 blocks of a regular grid
    xs, ys = np.meshgrid(np.linspace(-1, 1, 12), np.linspace(0, 3, 8))
    pts = np.column_stack((xs.ravel(), ys.ravel()))
    perm = morton_order(pts)
    assert sorted(perm.tolist()) == list(range(96))
    p = pts[perm]
    for b in range(0, 96, 4):
        assert len(set(p[b:b + 4, 0])) == 2 and len(set(p[b:b + 4, 1])) == 2
    for b in range(0, 96, 16):
        assert len(set(p[b:b + 16, 0])) == 4 and len(set(p[b:b + 16, 1])) == 4
