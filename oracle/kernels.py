"""Deterministic kernels of the hot path, restated in numpy.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  All reference citations are relative
to ``/root/reference``.
"""
from __future__ import annotations

import math

import numpy as np

LOG_SQRT_TWO_PI = 0.5 * math.log(2.0 * math.pi)
SQRT_TWO_PI = math.sqrt(2.0 * math.pi)


# --------------------------------------------------------------------------- #
# F1: Voronoi nearest-site rasterisation
# --------------------------------------------------------------------------- #
def nearest_site_2d(sites, points):
    """Index of the nearest Voronoi site for every query point (2-D Euclidean).

    Follows ``Voronoi2D.interpolate_tessellation`` (src/bayesbay/discretization/_voronoi.py:1429-1454)
    and the tomography forward (docs/source/tutorials/31_sw_tomography.ipynb cell 12), which call
    ``scipy.spatial.KDTree(sites).query(points)[1]``.  A KD-tree query is an exact nearest
    neighbour search, so it is restated as a brute-force ``argmin`` of the squared distance
    ``dx*dx + dy*dy`` (products and sum rounded separately, no FMA), ties to the lowest index.
    """
    sites = np.asarray(sites, dtype=np.float64)
    points = np.asarray(points, dtype=np.float64)
    out = np.empty(points.shape[0], dtype=np.int64)
    step = max(1, (1 << 22) // max(1, sites.shape[0]))
    for s in range(0, points.shape[0], step):
        p = points[s:s + step]
        dx = p[:, 0:1] - sites[None, :, 0]
        dy = p[:, 1:2] - sites[None, :, 1]
        out[s:s + step] = np.argmin(dx * dx + dy * dy, axis=1)
    return out


def nearest_neighbour_2d(sites, point):
    """``Voronoi.nearest_neighbour`` (src/bayesbay/discretization/_voronoi.py:462-465).

    The reference takes ``argmin(norm(sites - p))``; the square root is monotone so the
    squared distance is used (differs only when sqrt merges two distances 1 ulp apart).
    """
    sites = np.asarray(sites, dtype=np.float64)
    dx = sites[:, 0] - point[0]
    dy = sites[:, 1] - point[1]
    return int(np.argmin(dx * dx + dy * dy))


def nearest_neighbour_1d(xp, x):
    """``nearest_neighbour_1d`` (src/bayesbay/_utils_1d.pyx:102-114): sorted sites, linear scan,
    ``xp < x[0]`` -> 0, past the end -> last, tie -> upper cell."""
    n = len(x)
    if n == 1 or xp < x[0]:
        return 0
    for i in range(n - 1):
        x0 = x[i]
        x1 = x[i + 1]
        if x0 <= xp <= x1:
            if abs(x0 - xp) < abs(x1 - xp):
                return i
            return i + 1
    return n - 1


def nearest_site_1d(sites, xs):
    """Vectorised cell index of ``interpolate_nearest_1d`` (src/bayesbay/_utils_1d.pyx:87-100,171-192):
    ``xp <= x[0]`` -> 0; ``xp >= x[-1]`` -> last; else ``nearest_neighbour_1d``."""
    sites = np.asarray(sites, dtype=np.float64)
    xs = np.asarray(xs, dtype=np.float64)
    n = sites.size
    out = np.empty(xs.size, dtype=np.int64)
    if n == 1:
        out[:] = 0
        return out
    # first i with sites[i] <= xp <= sites[i+1]  ==  first i with sites[i+1] >= xp
    i = np.searchsorted(sites[1:], xs, side="left")
    i = np.minimum(i, n - 2)
    x0 = sites[i]
    x1 = sites[i + 1]
    idx = np.where(np.abs(x0 - xs) < np.abs(x1 - xs), i, i + 1)
    idx = np.where(xs <= sites[0], 0, idx)
    idx = np.where(xs >= sites[-1], n - 1, idx)
    out[:] = idx
    return out


def cell_extents_1d(sites, lb=0.0, ub=math.inf, fill_value=0.0):
    """``compute_voronoi1d_cell_extents`` (src/bayesbay/_utils_1d.pyx:30-52).

    Quirk Q9: the shipped ``-ffast-math`` build returns ``inf`` for unbounded cells where the
    source says ``fill_value``; the source semantics are restated here (finite entries agree).
    """
    sites = np.asarray(sites, dtype=np.float64)
    n = sites.size
    ext = np.zeros(n)
    if lb != -math.inf:
        istart = 0
        d1 = lb
    else:
        istart = 1
        ext[0] = fill_value
        d1 = (sites[0] + sites[1]) / 2.0
    d2 = d1
    i = istart - 1
    for i in range(istart, n - 1):
        d2 = (sites[i] + sites[i + 1]) / 2.0
        ext[i] = d2 - d1
        d1 = d2
    ext[i + 1] = ub - d2 if ub != math.inf else fill_value
    return ext


def insert_1d(values, index, value):
    """``insert_1d`` (src/bayesbay/_utils_1d.pyx:216-230)."""
    return np.concatenate((values[:index], [value], values[index:]))


def delete_1d(values, index):
    """``delete_1d`` (src/bayesbay/_utils_1d.pyx:233-241)."""
    return np.concatenate((values[:index], values[index + 1:]))


def inverse_covariance(sigma, r, n):
    """``inverse_covariance`` (src/bayesbay/_utils_1d.pyx:198-211): tridiagonal AR(1) precision."""
    factor = 1.0 / (sigma ** 2 * (1.0 - r ** 2))
    m = np.zeros((n, n))
    for i in range(1, n - 1):
        m[i, i] = (1 + r ** 2) * factor
        m[i, i - 1] = -r * factor
        m[i - 1, i] = -r * factor
    m[0, 0] = factor
    m[n - 1, n - 1] = factor
    m[n - 1, n - 2] = -r * factor
    m[n - 2, n - 1] = -r * factor
    return m


# --------------------------------------------------------------------------- #
# F2-F4: forward operators
# --------------------------------------------------------------------------- #
def ray_forward(indptr, indices, data, field):
    """``jacobian @ field`` for a CSR ray matrix (31_sw_tomography.ipynb cell 12,
    tests/21_test_voronoi_2d.py:57-60); row sums taken left to right."""
    prod = data * field[indices]
    out = np.zeros(indptr.size - 1)
    nz = indptr[1:] > indptr[:-1]
    red = np.add.reduceat(prod, indptr[:-1][nz]) if prod.size else np.zeros(0)
    out[nz] = red
    return out


def tomography_forward(fwd, sites, values):
    """d_pred of the straight-ray forward: nearest site -> field -> reciprocal -> CSR mat-vec."""
    idx = nearest_site_2d(sites, fwd["points"])
    field = values[idx]
    if fwd.get("transform", "reciprocal") == "reciprocal":
        field = 1.0 / field
    return ray_forward(fwd["indptr"], fwd["indices"], fwd["data"], field), idx


def lookup1d_forward(x, sites, values):
    """``Voronoi1D.interpolate_tessellation(sites, values, x, 'nuclei')``
    (src/bayesbay/discretization/_voronoi.py:705-732 -> _utils_1d.pyx:171-192)."""
    return values[nearest_site_1d(sites, x)]


def linear_forward(G, m):
    """``fwd_operator @ m`` (docs/source/tutorials/02_hierarchical_polyfit.ipynb cell 12)."""
    return G @ m


# --------------------------------------------------------------------------- #
# L1-L3: likelihood and prior
# --------------------------------------------------------------------------- #
def misfit_and_logdet(dpred, dobs, std=None, cov_inv=None, corr=None):
    """One target's contribution to ``LogLikelihood._get_misfit_and_det``
    (src/bayesbay/likelihood/_log_likelihood.py:308-332) with
    ``Target.inverse_covariance_times_vector`` (likelihood/_target.py:136-170; correlated noise uses the
    tridiagonal ``inverse_covariance``, _utils_1d.pyx:198-211) and ``Target.log_determinant_covariance``
    (likelihood/_target.py:172-208)."""
    r = dpred - dobs
    if cov_inv is not None:
        # `covariance_mat_inv @ vector` for a 2-D array, `covariance_mat_inv * vector` otherwise (_target.py:146-152)
        misfit = float(r @ (cov_inv @ r)) if np.ndim(cov_inv) == 2 else float(r @ (cov_inv * r))
        return misfit, 0.0
    n = dobs.size
    if corr is None:
        misfit = float(r @ (1.0 / std ** 2 * r))
        rho = 0.0
    else:
        misfit = float(r @ (inverse_covariance(std, corr, n) @ r))
        rho = corr
    log_det = (2 * n) * math.log(std) + (n - 1) * math.log(1.0 - rho ** 2)
    return misfit, log_det


def log_like_ratio(old_misfit, old_logdet, new_misfit, new_logdet, temperature=1.0):
    """``_log_likelihood_ratio_from_targets`` (likelihood/_log_likelihood.py:245-251) then ``/ T``
    (``log_likelihood_ratio``, :204-243)."""
    return (old_logdet - new_logdet + old_misfit - new_misfit) / 2 / temperature


def interpolate_linear_1d(xp, x, y):
    """``_interpolate_linear_1d`` (src/bayesbay/_utils_1d.pyx:68-84): clamp outside, linear inside."""
    if xp <= x[0]:
        return float(y[0])
    if xp >= x[-1]:
        return float(y[-1])
    i = 0
    while not (x[i] <= xp <= x[i + 1]):
        i += 1
    return float(y[i] + (xp - x[i]) * (y[i + 1] - y[i]) / (x[i + 1] - x[i]))


def prior_at(p, position=None):
    """Hyper-parameters of a prior at a position: ``Prior.get_hyper_param`` (src/bayesbay/prior/_prior.py:257-297).
    ``p["table"] = dict(position, a, b, perturb_std, perturb_std_birth)`` holds position-dependent ones."""
    tab = p.get("table")
    if tab is None:
        return p
    x = float(position)
    return dict(kind=p["kind"], name=p["name"], **{k: interpolate_linear_1d(x, tab["position"], tab[k])
                                                   for k in ("a", "b", "perturb_std", "perturb_std_birth")})


def prior_log_pdf(p, v):
    """``UniformPrior.log_prior`` (src/bayesbay/prior/_prior.py:453-481), ``GaussianPrior.log_prior``
    (:652-682), ``LaplacePrior.log_prior`` (:853-881) with scalar hyper-parameters."""
    kind = p["kind"]
    if kind == "uniform":
        return -math.log(p["b"] - p["a"]) if p["a"] <= v <= p["b"] else -math.inf
    if kind == "gaussian":
        return -0.5 * math.log(2 * math.pi) - math.log(p["b"]) - 0.5 * ((v - p["a"]) / p["b"]) ** 2
    if kind == "laplace":
        return -math.log(2) - math.log(p["b"]) - abs(v - p["a"]) / p["b"]
    raise ValueError(kind)


def space_log_prior(params, values, k):
    """``ParameterSpace.log_prior`` (src/bayesbay/parameterization/_parameter_space.py:189-195):
    sum over parameters, sum over cells."""
    tot = 0.0
    for p, v in zip(params, values):
        tot += sum(prior_log_pdf(p, float(x)) for x in v[:k])
    return tot


def prior_value_ratio(p, old, new):
    """log prior ratio of a Gaussian random-walk value move: ``UniformPrior.perturb_value``
    (prior/_prior.py:400-451), ``GaussianPrior.perturb_value`` (:599-650),
    ``LaplacePrior.perturb_value`` (:800-851)."""
    kind = p["kind"]
    if kind == "uniform":
        return 0.0 if p["a"] <= new <= p["b"] else -math.inf
    if kind == "gaussian":
        return (old - new) * (old + new - 2 * p["a"]) / (2 * p["b"] ** 2)
    if kind == "laplace":
        return (abs(old - p["a"]) - abs(new - p["a"])) / p["b"]
    raise ValueError(kind)
