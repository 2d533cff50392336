"""The plug-in seam: the UNMODIFIED reference (baseline/_ref) driving the CUDA engine through its own
``BayesianInversion.run(sampler=...)`` (src/bayesbay/_bayes_inversion.py:236-249)."""
import numpy as np
import pytest

from bayesbay_b200 import synthetic
from oracle import reference_runner as RR

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not RR.available(), reason="baseline/_ref is not installed")]


def _inversion(ing, n_chains):
    bayesbay = RR.import_reference()
    par, ll = RR.build_problem(bayesbay, ing)
    return bayesbay, bayesbay.BayesianInversion(par, ll, n_chains=n_chains)


def test_reference_inversion_runs_on_the_cuda_sampler():
    from bayesbay_b200.plugin import cuda_samplers

    ing = synthetic.tomography_2d(nx=12, ny=10, n_sources_per_side=3, n_receivers=5, kmin=4, kmax=20)
    bayesbay, inv = _inversion(ing, 24)
    start = [ch.current_state.copy() for ch in inv.chains]
    CudaVanillaSampler, _ = cuda_samplers(bayesbay, seed=11)
    sampler = CudaVanillaSampler()
    inv.run(sampler=sampler, n_iterations=600, burnin_iterations=200, save_every=50, verbose=False)
    # the chains started from the states the reference drew for them ...
    assert sampler.batch.iteration == 600
    res = inv.get_results()  # the reference's own get_results_from_chains
    n_saved = 24 * (600 - 200) // 50
    for key in ("voronoi.n_dimensions", "voronoi.discretization", "voronoi.vel", "d_obs.std", "d_obs.dpred"):
        assert len(res[key]) == n_saved, key
    k0 = res["voronoi.n_dimensions"][0]
    assert res["voronoi.discretization"][0].shape == (k0, 2) and res["voronoi.vel"][0].shape == (k0,)
    assert res["d_obs.dpred"][0].shape == ing["d_obs"].shape
    # ... and carry the reference's statistics layout
    for ch in inv.chains:
        st = ch.statistics
        assert st["n_proposed_models_total"] == 600 == ch.ith_iteration
        outer = "ParamSpacePerturbation(param_space_name=voronoi)"
        assert set(st["n_proposed_models"][outer]) <= {"BirthPerturbation(voronoi)", "DeathPerturbation(voronoi)",
                                                       "ParamPerturbation(voronoi.vel)", "ParamPerturbation(voronoi.discretization)"}
        total = sum(st["n_proposed_models"][outer].values()) + st["n_proposed_models"].get("NoisePerturbation(d_obs)", 0)
        assert total == 600
        assert isinstance(ch.current_state, bayesbay.State)
        assert 4 <= ch.current_state["voronoi"].n_dimensions <= 20
    # the saved d_pred is the forward of the saved state, by the reference's own host evaluation of the operator
    fwd = inv.chains[0].log_likelihood.fwd_functions[0]
    for i in (0, n_saved // 2, n_saved - 1):
        ps = bayesbay.ParameterSpaceState(res["voronoi.n_dimensions"][i],
                                          {"discretization": res["voronoi.discretization"][i], "vel": res["voronoi.vel"][i]})
        st = bayesbay.State({"voronoi": ps})
        np.testing.assert_allclose(res["d_obs.dpred"][i], fwd(st), rtol=1e-12, atol=1e-15)
    # the fit improved from the reference's starting states
    def misfit(state):
        return float(np.sum((fwd(state) - ing["d_obs"]) ** 2))

    m0 = np.median([misfit(s) for s in start])
    m1 = np.median([misfit(ch.current_state) for ch in inv.chains])
    assert m1 < 0.5 * m0
    # a second run continues the same chains (iteration count, samples appended)
    inv.run(sampler=sampler, n_iterations=100, burnin_iterations=0, save_every=50, verbose=False)
    assert all(ch.ith_iteration == 700 for ch in inv.chains)
    assert len(inv.get_results()["voronoi.n_dimensions"]) == n_saved + 24 * 2


def test_reference_inversion_parallel_tempering_on_the_cuda_sampler():
    from bayesbay_b200.plugin import cuda_samplers

    ing = synthetic.partition_1d()
    bayesbay, inv = _inversion(ing, 10)
    _, CudaParallelTempering = cuda_samplers(bayesbay, seed=5)
    inv.run(sampler=CudaParallelTempering(temperature_max=4, swap_every=50), n_iterations=400, burnin_iterations=100,
            save_every=25, verbose=False)
    temps = sorted(ch.temperature for ch in inv.chains)
    ref_ladder = sorted(np.concatenate((np.ones(3), np.geomspace(1, 4, 7))))  # reference rule (samplers/_samplers.py:315-326)
    np.testing.assert_allclose(temps, ref_ladder, rtol=1e-14)  # temperatures are permuted, never changed
    res = inv.get_results()
    assert len(res["voronoi.n_dimensions"]) > 0
    assert all(ch.ith_iteration == 400 for ch in inv.chains)


def test_plugin_rejects_python_callables():
    from bayesbay_b200.plugin import cuda_samplers

    bayesbay = RR.import_reference()
    ing = synthetic.partition_1d()
    par, _ = RR.build_problem(bayesbay, ing)
    target = bayesbay.likelihood.Target("d_obs", ing["d_obs"], std_min=0, std_max=20, std_perturb_std=2)
    ll = bayesbay.likelihood.LogLikelihood(targets=target, fwd_functions=lambda state: np.zeros(ing["d_obs"].size))
    inv = bayesbay.BayesianInversion(par, ll, n_chains=2)
    CudaVanillaSampler, _ = cuda_samplers(bayesbay)
    with pytest.raises(TypeError, match="no CPU fallback"):
        inv.run(sampler=CudaVanillaSampler(), n_iterations=10, verbose=False)
    with pytest.raises(TypeError):
        CudaVanillaSampler().add_on_begin_iteration(lambda s, c: None)
