"""Spec dicts (the plain-data problem format shared by the oracle and the product compiler)
built straight from the synthetic ingredients -- independent of the product's API classes, so
``tests/test_compile.py`` can check that ``bayesbay_b200._compile`` emits the same thing."""
import numpy as np

from bayesbay_b200 import synthetic


def _prior(p):
    p = dict(p)
    kind = p.pop("kind", "uniform")
    if kind == "uniform":
        a, b = p["vmin"], p["vmax"]
    elif kind == "gaussian":
        a, b = p["mean"], p["std"]
    else:
        a, b = p["mean"], p["scale"]
    std_birth = p.get("perturb_std_birth")
    std_birth = p["perturb_std"] if std_birth is None else std_birth
    if p.get("position") is not None:
        pos = np.asarray(p["position"], dtype=np.float64)
        col = lambda v: np.broadcast_to(np.asarray(v, dtype=np.float64), pos.shape).copy()  # noqa: E731
        return dict(name=p["name"], kind=kind, a=None, b=None, perturb_std=None, perturb_std_birth=None,
                    table=dict(position=pos, a=col(a), b=col(b), perturb_std=col(p["perturb_std"]),
                               perturb_std_birth=col(std_birth)))
    return dict(name=p["name"], kind=kind, a=float(a), b=float(b), perturb_std=float(p["perturb_std"]),
                perturb_std_birth=float(std_birth))


def _target(ing, forward):
    t = ing["target"]
    cov = t.get("covariance_mat_inv")
    out = dict(name=t["name"], dobs=np.asarray(ing["d_obs"], dtype=np.float64), hierarchical=cov is None,
               std_min=float(t.get("std_min", 0.01)), std_max=float(t.get("std_max", 1)),
               std_perturb_std=float(t.get("std_perturb_std", 0.1)),
               cov_inv=cov, correlated=bool(t.get("noise_is_correlated", False)),
               corr_min=float(t.get("correlation_min", 0.01)), corr_max=float(t.get("correlation_max", 1.0)),
               corr_perturb_std=float(t.get("correlation_perturb_std", 0.1)), forward=forward)
    return out


def spec_from_ingredients(ing):
    kind = ing["kind"]
    if kind in ("tomography_2d", "partition_1d"):
        v = ing["voronoi"]
        ndim = 2 if kind == "tomography_2d" else 1
        trans_d = v.get("n_dimensions") is None
        kmin, kmax = (v["n_dimensions_min"], v["n_dimensions_max"]) if trans_d else (v["n_dimensions"],) * 2
        space = dict(
            name=v["name"], kind="voronoi2d" if ndim == 2 else "voronoi1d", ndim=ndim, trans_d=trans_d,
            kmin=kmin, kmax=kmax, kinit_max=int((kmax - kmin) * 0.3 + kmin),
            vmin=[float(x) for x in np.atleast_1d(v["vmin"])], vmax=[float(x) for x in np.atleast_1d(v["vmax"])],
            site_std=[float(v["perturb_std"])] * ndim, birth_from=v.get("birth_from", "neighbour"),
            params=[_prior(ing["prior"])],
            weights=dict(birth=1 if trans_d else 0, death=1 if trans_d else 0, values=3, sites=1),
        )
        if ndim == 2:
            fwd = dict(kind="ray2d", space=0, param=0, transform="reciprocal", indptr=ing["indptr"],
                       indices=ing["indices"], data=ing["data"], points=ing["points"])
        else:
            fwd = dict(kind="lookup1d", space=0, param=0, x=ing["x"])
    elif kind == "transd_polyfit":
        sp = ing["space"]
        kmin, kmax = sp["n_dimensions_min"], sp["n_dimensions_max"]
        space = dict(name=sp["name"], kind="plain", ndim=0, trans_d=True, kmin=kmin, kmax=kmax,
                     kinit_max=int((kmax - kmin) * 0.3 + kmin), vmin=[], vmax=[], site_std=[], birth_from="prior",
                     params=[_prior(ing["prior"])], weights=dict(birth=1, death=1, values=3, sites=0))
        fwd = dict(kind="polynomial", space=0, param=0, x=ing["x"])
    elif kind == "polyfit":
        space = dict(name=ing["space"]["name"], kind="plain", ndim=0, trans_d=False, kmin=1, kmax=1, kinit_max=1,
                     vmin=[], vmax=[], site_std=[], birth_from="prior", params=[_prior(p) for p in ing["priors"]],
                     weights=dict(birth=0, death=0, values=3, sites=0))
        fwd = dict(kind="linear", space=0, params=[0, 1, 2, 3], G=ing["G"])
    else:
        raise ValueError(kind)
    return dict(spaces=[space], targets=[_target(ing, fwd)])


def tomo_small():
    return synthetic.tomography_2d(nx=12, ny=10, n_sources_per_side=3, n_receivers=5, kmin=4, kmax=20)


def partition_small(birth_from="prior"):
    return synthetic.partition_1d(n_data=30, kmin=2, kmax=12, birth_from=birth_from)


REPLAYS = {
    "tomo_small": tomo_small,
    "partition_1d": partition_small,
    "partition_1d_nb": lambda: partition_small("neighbour"),
    "partition_1d_corr": lambda: synthetic.partition_1d(n_data=30, kmin=2, kmax=12, correlated=True),
    "partition_1d_posdep": lambda: synthetic.partition_1d(n_data=30, kmin=2, kmax=12, birth_from="neighbour",
                                                          position_dependent=True),
    "partition_1d_laplace": lambda: synthetic.partition_1d(n_data=30, kmin=2, kmax=12, birth_from="neighbour",
                                                           prior_kind="laplace"),
    "partition_1d_fixed": synthetic.partition_1d_fixed,
    "transd_polyfit": synthetic.transd_polyfit,
    "polyfit": synthetic.polyfit,
    "polyfit_nb": lambda: synthetic.polyfit(all_gaussian=False),
}


def joint_small():
    """Two parameter spaces and three targets with different noise models: a fixed-dimension plain space read by a
    linear forward (hierarchical noise) and a per-datum-weighted copy of it, and a trans-d Voronoi1D space with
    Gaussian + Laplace parameters read by a 1-D lookup (fixed scalar inverse covariance).  Spec dict only."""
    poly = synthetic.polyfit()
    part = synthetic.partition_1d(n_data=25, kmin=1, kmax=9, birth_from="neighbour")
    rng = np.random.default_rng(4)
    plain = dict(name="coeffs", kind="plain", ndim=0, trans_d=False, kmin=1, kmax=1, kinit_max=1, vmin=[], vmax=[],
                 site_std=[], birth_from="prior", params=[_prior(p) for p in poly["priors"]],
                 weights=dict(birth=0, death=0, values=3, sites=0))
    vor = dict(name="layers", kind="voronoi1d", ndim=1, trans_d=True, kmin=1, kmax=9, kinit_max=3, vmin=[0.0], vmax=[10.0],
               site_std=[0.75], birth_from="neighbour",
               params=[dict(name="y", kind="gaussian", a=0.0, b=15.0, perturb_std=3.0, perturb_std_birth=4.0),
                       dict(name="z", kind="laplace", a=1.0, b=2.0, perturb_std=0.5, perturb_std_birth=0.5)],
               weights=dict(birth=1, death=1, values=3, sites=1))

    def tgt(name, dobs, hier, cov, fwd, **kw):
        return dict(name=name, dobs=np.asarray(dobs, dtype=np.float64), hierarchical=hier, std_min=0.0, std_max=40.0,
                    std_perturb_std=4.0, cov_inv=cov, correlated=False, corr_min=0.01, corr_max=1.0,
                    corr_perturb_std=0.1, forward=fwd, **kw)

    targets = [
        tgt("poly", poly["d_obs"], True, None, dict(kind="linear", space=0, params=[0, 1, 2, 3], G=poly["G"])),
        tgt("poly_w", poly["d_obs"] + 1.0, False, rng.uniform(0.001, 0.004, poly["d_obs"].size),
            dict(kind="linear", space=0, params=[3, 2, 1, 0], G=poly["G"][:, ::-1].copy())),
        tgt("steps", part["d_obs"], False, 0.04, dict(kind="lookup1d", space=1, param=0, x=part["x"])),
    ]
    return dict(spaces=[plain, vor], targets=targets)


def tomo_joint_small():
    """Ray tomography plus a plain space with its own linear target: exercises the pipeline mode with a
    thread-evaluated target next to the ray target (spec dict only)."""
    spec = spec_from_ingredients(tomo_small())
    poly = synthetic.polyfit()
    spec["spaces"].append(dict(name="coeffs", kind="plain", ndim=0, trans_d=False, kmin=1, kmax=1, kinit_max=1, vmin=[],
                               vmax=[], site_std=[], birth_from="prior", params=[_prior(p) for p in poly["priors"]],
                               weights=dict(birth=0, death=0, values=3, sites=0)))
    spec["targets"].append(dict(name="poly", dobs=poly["d_obs"], hierarchical=True, std_min=0.0, std_max=40.0,
                                std_perturb_std=4.0, cov_inv=None, correlated=False, corr_min=0.01, corr_max=1.0,
                                corr_perturb_std=0.1,
                                forward=dict(kind="linear", space=1, params=[0, 1, 2, 3], G=poly["G"])))
    return spec


def tomo_weighted_fixed():
    """Ray tomography variant for the chain-local step: per-datum inverse covariance (no noise move), the cell value
    used as it is (slowness parameterisation, `transform="identity"`), Gaussian prior, FIXED number of cells (only value
    and site moves).  Spec dict only."""
    ing = synthetic.tomography_2d(nx=16, ny=12, n_sources_per_side=3, n_receivers=6, kmin=4, kmax=20)
    spec = spec_from_ingredients(ing)
    sp = spec["spaces"][0]
    sp.update(trans_d=False, kmin=9, kmax=9, kinit_max=9, weights=dict(birth=0, death=0, values=3, sites=2))
    sp["params"] = [dict(name="slow", kind="gaussian", a=0.33, b=0.05, perturb_std=0.01, perturb_std_birth=0.01)]
    rng = np.random.default_rng(8)
    tg = spec["targets"][0]
    tg.update(hierarchical=False, cov_inv=rng.uniform(2e4, 2e5, np.asarray(tg["dobs"]).size))
    tg["forward"] = dict(tg["forward"], transform="identity")
    return spec


def tomo_scalar_prior_birth():
    """Ray tomography variant: fixed scalar inverse covariance, cells born from the prior."""
    ing = synthetic.tomography_2d(nx=16, ny=12, n_sources_per_side=3, n_receivers=6, kmin=2, kmax=30)
    spec = spec_from_ingredients(ing)
    spec["spaces"][0]["birth_from"] = "prior"
    spec["targets"][0].update(hierarchical=False, cov_inv=2.5e5)
    return spec


def _spd_inverse_covariance(n, scale, seed):
    """A dense symmetric positive-definite inverse covariance [n][n]: exponentially decaying correlations between
    neighbouring data (full matrix, `covariance_mat_inv.ndim == 2`, likelihood/_target.py:150-151)."""
    rng = np.random.default_rng(seed)
    i = np.arange(n)
    cov = np.exp(-np.abs(i[:, None] - i[None, :]) / 2.5) * np.outer(s := rng.uniform(0.8, 1.25, n), s)
    return np.linalg.inv(cov) * scale


def joint_matrix_cov():
    """joint_small with FULL inverse covariance matrices on the weighted linear target and on the 1-D lookup target."""
    spec = joint_small()
    for t, scale, seed in ((1, 0.0025, 11), (2, 0.04, 12)):
        tg = spec["targets"][t]
        tg.update(hierarchical=False, cov_inv=_spd_inverse_covariance(tg["dobs"].size, scale, seed))
    return spec


def tomo_matrix_cov():
    """Ray tomography with a FULL inverse covariance matrix on the travel times (runs the full-SpMM step: the misfit is
    reduced from the materialised residual vector)."""
    ing = synthetic.tomography_2d(nx=16, ny=12, n_sources_per_side=3, n_receivers=6, kmin=2, kmax=30)
    spec = spec_from_ingredients(ing)
    tg = spec["targets"][0]
    tg.update(hierarchical=False, cov_inv=_spd_inverse_covariance(np.asarray(tg["dobs"]).size, 2.5e5, 13))
    return spec


# oracle-vs-device cases without a reference replay fixture
SPECS = {"joint_small": joint_small, "tomo_joint_small": tomo_joint_small, "tomo_weighted_fixed": tomo_weighted_fixed,
         "tomo_scalar_prior_birth": tomo_scalar_prior_birth, "joint_matrix_cov": joint_matrix_cov,
         "tomo_matrix_cov": tomo_matrix_cov}
EXTRA = {
    "tomo_small_corr": lambda: synthetic.tomography_2d(nx=12, ny=10, n_sources_per_side=3, n_receivers=5, kmin=4,
                                                       kmax=20, correlated=True),
}
ALL = {**REPLAYS, **EXTRA}
