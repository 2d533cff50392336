"""Seeded synthetic problem generators for the BASELINE.json configurations.

These produce *ingredients* (plain numpy arrays + hyper-parameters), never API
objects, so the very same ingredients can be fed to

* ``bayesbay_b200`` (this package, the CUDA path),
* the unmodified reference ``bayesbay`` (golden-vector generation, see
  ``tests/golden/make_golden.py``), and
* the CPU oracle under ``oracle/``.

Shapes follow SURVEY.md section 8(d):

* C1  ``partition_1d``   -- docs/source/tutorials/42_transd_partition_mod.ipynb cells 3-14
* C2  ``polyfit``        -- docs/source/tutorials/02_hierarchical_polyfit.ipynb cells 3-18
* C3/C4/C5 ``tomography_2d`` -- docs/source/tutorials/31_sw_tomography.ipynb cells 4-15
  (seislib is replaced by the exact straight-ray generator below).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "transd_polyfit",
    "partition_1d_fixed",
    "straight_ray_csr",
    "regular_grid_points",
    "checkerboard",
    "tomography_2d",
    "partition_1d",
    "polyfit",
    "build_problem",
]


# --------------------------------------------------------------------------- #
# straight-ray geometry
# --------------------------------------------------------------------------- #
def regular_grid_points(nx, ny, xmin=-1.0, xmax=1.0, ymin=-1.0, ymax=1.0):
    """Cell-centred points of an nx-by-ny grid; index g = iy * nx + ix."""
    dx = (xmax - xmin) / nx
    dy = (ymax - ymin) / ny
    xs = xmin + (np.arange(nx) + 0.5) * dx
    ys = ymin + (np.arange(ny) + 0.5) * dy
    gx, gy = np.meshgrid(xs, ys, indexing="xy")
    return np.column_stack((gx.ravel(), gy.ravel()))


def straight_ray_csr(sources, receivers, nx, ny, xmin=-1.0, xmax=1.0, ymin=-1.0, ymax=1.0,
                     normalise=True):
    """Exact segment/grid intersection lengths for every (source, receiver) pair.

    Returns CSR arrays ``(indptr int32 [R+1], indices int32 [nnz], data float64 [nnz])``
    of the ray matrix ``G[r, g] = l_rg / L_r`` (rows sum to one when
    ``normalise``), so that ``G @ (1 / v)`` is the mean slowness along the ray -- the
    quantity the tomography notebook's Jacobian produces (31_sw_tomography.ipynb cell 6-7).
    Ray order is ``r = isource * n_receivers + ireceiver`` (np.ndindex order of cell 7).
    """
    sources = np.asarray(sources, dtype=np.float64)
    receivers = np.asarray(receivers, dtype=np.float64)
    dx = (xmax - xmin) / nx
    dy = (ymax - ymin) / ny
    xlines = xmin + dx * np.arange(nx + 1)
    ylines = ymin + dy * np.arange(ny + 1)
    indptr = [0]
    indices = []
    data = []
    for s in sources:
        for rcv in receivers:
            x0, y0 = s
            x1, y1 = rcv
            ts = [np.array([0.0, 1.0])]
            if x1 != x0:
                t = (xlines - x0) / (x1 - x0)
                ts.append(t[(t > 0.0) & (t < 1.0)])
            if y1 != y0:
                t = (ylines - y0) / (y1 - y0)
                ts.append(t[(t > 0.0) & (t < 1.0)])
            t = np.unique(np.concatenate(ts))
            seg = np.diff(t)
            tm = 0.5 * (t[:-1] + t[1:])
            xm = x0 + tm * (x1 - x0)
            ym = y0 + tm * (y1 - y0)
            ix = np.floor((xm - xmin) / dx).astype(np.int64)
            iy = np.floor((ym - ymin) / dy).astype(np.int64)
            ok = (ix >= 0) & (ix < nx) & (iy >= 0) & (iy < ny) & (seg > 0)
            length = np.hypot(x1 - x0, y1 - y0)
            cols = (iy[ok] * nx + ix[ok]).astype(np.int64)
            vals = seg[ok] * length
            # merge duplicate cells (cannot happen for a straight ray, but keep CSR canonical)
            order = np.argsort(cols, kind="stable")
            cols = cols[order]
            vals = vals[order]
            if cols.size:
                keep = np.concatenate(([True], cols[1:] != cols[:-1]))
                if not keep.all():
                    grp = np.cumsum(keep) - 1
                    vals = np.bincount(grp, weights=vals)
                    cols = cols[keep]
                if normalise:
                    vals = vals / vals.sum()
            indices.append(cols.astype(np.int32))
            data.append(vals.astype(np.float64))
            indptr.append(indptr[-1] + cols.size)
    return (
        np.asarray(indptr, dtype=np.int32),
        np.concatenate(indices) if indices else np.zeros(0, np.int32),
        np.concatenate(data) if data else np.zeros(0, np.float64),
    )


def checkerboard(points, kx=3, ky=3, ref=3.0, amp=0.3, xmin=-1.0, xmax=1.0, ymin=-1.0, ymax=1.0):
    """kx-by-ky checkerboard ``ref +- amp`` evaluated at ``points`` (G, 2)."""
    ix = np.minimum((np.asarray(points)[:, 0] - xmin) / (xmax - xmin) * kx, kx - 1e-9).astype(int)
    iy = np.minimum((np.asarray(points)[:, 1] - ymin) / (ymax - ymin) * ky, ky - 1e-9).astype(int)
    return ref + amp * np.where((ix + iy) % 2 == 0, 1.0, -1.0)


def _boundary_sources(n_per_side, xmin=-1.0, xmax=1.0, ymin=-1.0, ymax=1.0):
    d = (xmax - xmin) / n_per_side
    mid_x = xmin + d / 2 + d * np.arange(n_per_side)
    d = (ymax - ymin) / n_per_side
    mid_y = ymin + d / 2 + d * np.arange(n_per_side)
    return np.concatenate(
        (
            np.column_stack((mid_x, np.full(n_per_side, ymin))),
            np.column_stack((np.full(n_per_side, xmax), mid_y)),
            np.column_stack((mid_x, np.full(n_per_side, ymax))),
            np.column_stack((np.full(n_per_side, xmin), mid_y)),
        )
    )


# --------------------------------------------------------------------------- #
# problem ingredients
# --------------------------------------------------------------------------- #
def tomography_2d(nx=100, ny=100, n_sources_per_side=25, n_receivers=100, kmin=10, kmax=200,
                  noise_std=0.002, seed_receivers=1, seed_noise=2, correlated=False):
    """2-D Voronoi straight-ray travel-time tomography (configs C3/C4/C5).

    C3: defaults (100x100 grid, 100 x 100 = 10 000 rays, k in [10, 200]).
    C5: ``nx=ny=256, n_sources_per_side=100, n_receivers=250, kmin=50, kmax=500``.
    """
    points = regular_grid_points(nx, ny)
    sources = _boundary_sources(n_sources_per_side)
    receivers = np.random.default_rng(seed_receivers).uniform(-1.0, 1.0, (n_receivers, 2))
    indptr, indices, data = straight_ray_csr(sources, receivers, nx, ny)
    vel_true = checkerboard(points)
    n_rays = indptr.size - 1
    import scipy.sparse

    jac = scipy.sparse.csr_matrix((data, indices, indptr), shape=(n_rays, nx * ny))
    d_clean = jac @ (1.0 / vel_true)
    d_obs = d_clean + np.random.default_rng(seed_noise).normal(0.0, noise_std, n_rays)
    return dict(
        kind="tomography_2d",
        nx=nx, ny=ny, points=points, indptr=indptr, indices=indices, data=data,
        vel_true=vel_true, d_obs=d_obs,
        prior=dict(name="vel", vmin=2.0, vmax=4.0, perturb_std=0.1),
        voronoi=dict(name="voronoi", vmin=[-1.0, -1.0], vmax=[1.0, 1.0], perturb_std=0.05,
                     n_dimensions_min=kmin, n_dimensions_max=kmax, birth_from="neighbour",
                     compute_kdtree=True),
        target=dict(name="d_obs", std_min=0.0, std_max=0.01, std_perturb_std=0.001,
                    **(dict(noise_is_correlated=True, correlation_min=0.01, correlation_max=0.9,
                            correlation_perturb_std=0.1) if correlated else {})),
    )


def _piecewise(x):
    edges = [1.0, 2.5, 3.0, 4.0, 6.0, 6.5, 8.0, 9.0, 10.0]
    vals = [1.0, 20.0, 0.0, -3.0, -10.0, -20.0, 25.0, 0.0, 10.0]
    out = np.empty_like(x)
    for i, xv in enumerate(x):
        for e, v in zip(edges, vals):
            if xv <= e:
                out[i] = v
                break
    return out


def partition_1d(n_data=100, noise_std=5.0, seed=30, kmin=2, kmax=40, birth_from="prior", correlated=False,
                 position_dependent=False, prior_kind="uniform"):
    """1-D Voronoi piecewise-constant regression (config C1).  ``position_dependent`` gives the value prior
    position-dependent bounds and step sizes (reference prior/_prior.py:257-297, tests/19_*.py);
    ``prior_kind="laplace"`` puts a LaplacePrior(mean 0, scale 15) on the cell values."""
    x = np.linspace(0.0, 10.0, n_data)
    rng = np.random.RandomState(seed)  # the notebook uses legacy np.random.seed(30)
    d_obs = _piecewise(x) + rng.normal(0.0, noise_std, x.size)
    return dict(
        kind="partition_1d",
        x=x, d_obs=d_obs,
        prior=(dict(name="y", vmin=[-35.0, -20.0, -35.0], vmax=[35.0, 40.0, 10.0], perturb_std=[3.5, 2.0, 4.0],
                    perturb_std_birth=[5.0, 3.0, 5.0], position=[0.0, 5.0, 10.0]) if position_dependent
               else (dict(kind="laplace", name="y", mean=0.0, scale=15.0, perturb_std=3.5) if prior_kind == "laplace"
                     else dict(name="y", vmin=-35.0, vmax=35.0, perturb_std=3.5))),
        voronoi=dict(name="voronoi", vmin=0.0, vmax=10.0, perturb_std=0.75,
                     n_dimensions_min=kmin, n_dimensions_max=kmax, birth_from=birth_from),
        target=dict(name="d_obs", std_min=0.0, std_max=20.0, std_perturb_std=2.0,
                    **(dict(noise_is_correlated=True, correlation_min=0.01, correlation_max=0.9,
                            correlation_perturb_std=0.1) if correlated else {})),
    )


def polyfit(n_data=15, noise_std=20.0, seed=30, all_gaussian=True):
    """Fixed-dimension cubic polynomial regression with hierarchical noise (config C2).

    ``all_gaussian=True`` is BASELINE.json config 2 (Gaussian priors); ``False`` is the
    notebook variant (m1, m2 uniform) whose statistics the reference publishes.
    """
    x = np.linspace(-5.0, 10.0, n_data)
    G = np.vander(x, 4, True)
    truth = np.array([20.0, -10.0, -3.0, 1.0])
    rng = np.random.RandomState(seed)
    d_obs = G @ truth + rng.normal(0.0, noise_std, n_data)
    if all_gaussian:
        priors = [
            dict(kind="gaussian", name="m0", mean=20.0, std=1.0, perturb_std=0.5),
            dict(kind="gaussian", name="m1", mean=-10.0, std=2.0, perturb_std=0.4),
            dict(kind="gaussian", name="m2", mean=-3.0, std=2.0, perturb_std=0.5),
            dict(kind="gaussian", name="m3", mean=1.0, std=0.1, perturb_std=0.1),
        ]
    else:
        priors = [
            dict(kind="gaussian", name="m0", mean=20.0, std=1.0, perturb_std=0.5),
            dict(kind="uniform", name="m1", vmin=-13.0, vmax=-7.0, perturb_std=0.4),
            dict(kind="uniform", name="m2", vmin=-10.0, vmax=4.0, perturb_std=0.5),
            dict(kind="gaussian", name="m3", mean=1.0, std=0.1, perturb_std=0.1),
        ]
    return dict(
        kind="polyfit",
        x=x, G=G, d_obs=d_obs, truth=truth, priors=priors,
        space=dict(name="my_param_space", n_dimensions=1),
        target=dict(name="my_data", std_min=0.0, std_max=40.0, std_perturb_std=4.0),
    )


def transd_polyfit(n_data=20, noise_std=20.0, seed=30, kmin=1, kmax=8):
    """Trans-dimensional polynomial regression: the number of coefficients is unknown
    (docs/source/tutorials/03_transd_polyfit.ipynb); known data noise (fixed scalar inverse covariance)."""
    x = np.linspace(-5.0, 10.0, n_data)
    truth = np.array([20.0, -10.0, -3.0, 1.0])
    rng = np.random.RandomState(seed)
    d_obs = np.vander(x, 4, True) @ truth + rng.normal(0.0, noise_std, n_data)
    return dict(
        kind="transd_polyfit", x=x, d_obs=d_obs, truth=truth,
        prior=dict(name="coefficients", vmin=-25.0, vmax=25.0, perturb_std=0.5),
        space=dict(name="my_param_space", n_dimensions_min=kmin, n_dimensions_max=kmax),
        target=dict(name="my_data", covariance_mat_inv=1.0 / noise_std ** 2),
    )


def partition_1d_fixed(n_data=30, n_cells=9, **kw):
    """Fixed number of Voronoi cells (docs/source/tutorials/41_simple_partition_mod.ipynb): only value and site
    moves."""
    ing = partition_1d(n_data=n_data, **kw)
    ing["voronoi"] = dict(name="voronoi", vmin=0.0, vmax=10.0, perturb_std=0.75, n_dimensions=n_cells)
    return ing


# --------------------------------------------------------------------------- #
# API-object builder (works with bayesbay_b200 AND the unmodified reference)
# --------------------------------------------------------------------------- #
def forward_operator(ing):
    """The device forward operator (bayesbay_b200.forward) of a synthetic problem -- also a valid forward function
    for the unmodified reference."""
    from .forward import LinearForward, NearestLookup1D, PolynomialForward, RayTomographyForward

    kind = ing["kind"]
    if kind == "tomography_2d":
        return RayTomographyForward((ing["indptr"], ing["indices"], ing["data"]), ing["points"], ing["voronoi"]["name"],
                                    ing["prior"]["name"])
    if kind == "partition_1d":
        return NearestLookup1D(ing["x"], ing["voronoi"]["name"], ing["prior"]["name"])
    if kind == "transd_polyfit":
        return PolynomialForward(ing["x"], ing["space"]["name"], ing["prior"]["name"])
    if kind == "polyfit":
        return LinearForward(ing["G"], ing["space"]["name"], [p["name"] for p in ing["priors"]])
    raise ValueError(kind)


def build_problem(ing, bb, make_forward=forward_operator):
    """Build ``(parameterization, log_likelihood)`` from ingredients using package ``bb``.

    ``bb`` is either :mod:`bayesbay_b200` or the reference :mod:`bayesbay` -- the
    constructor calls are identical, which is the drop-in claim.  ``make_forward(ing)``
    returns the forward: a device operator descriptor for ``bayesbay_b200`` or a plain
    ``state -> ndarray`` callable for the reference.
    """
    kind = ing["kind"]
    if kind == "tomography_2d":
        p = ing["prior"]
        vel = bb.prior.UniformPrior(p["name"], vmin=p["vmin"], vmax=p["vmax"], perturb_std=p["perturb_std"])
        space = bb.discretization.Voronoi2D(parameters=[vel], **ing["voronoi"])
    elif kind == "partition_1d":
        p = dict(ing["prior"])
        y = {"uniform": bb.prior.UniformPrior, "laplace": bb.prior.LaplacePrior}[p.pop("kind", "uniform")](**p)
        space = bb.discretization.Voronoi1D(parameters=[y], **ing["voronoi"])
    elif kind == "transd_polyfit":
        coeff = bb.prior.UniformPrior(**ing["prior"])
        space = bb.parameterization.ParameterSpace(parameters=[coeff], **ing["space"])
    elif kind == "polyfit":
        params = []
        for p in ing["priors"]:
            p = dict(p)
            k = p.pop("kind")
            cls = {"gaussian": bb.prior.GaussianPrior, "uniform": bb.prior.UniformPrior}[k]
            params.append(cls(**p))
        space = bb.parameterization.ParameterSpace(parameters=params, **ing["space"])
    else:
        raise ValueError(kind)
    parameterization = bb.parameterization.Parameterization(space)
    t = ing["target"]
    target = bb.likelihood.Target(t["name"], ing["d_obs"], **{k: v for k, v in t.items() if k != "name"})
    log_likelihood = bb.likelihood.LogLikelihood(targets=target, fwd_functions=make_forward(ing))
    return parameterization, log_likelihood
