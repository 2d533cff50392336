"""Device-native forward operators ("second seam").

The reference's ``LogLikelihood(targets, fwd_functions)`` takes arbitrary Python callables
``state -> d_pred`` (src/bayesbay/likelihood/_log_likelihood.py:67-78).  A Python callable cannot
run inside a CUDA kernel and this package has no CPU fallback, so the forward functions of the
tutorials are provided as operator *descriptors* that carry the arrays the engine uploads:

* :class:`RayTomographyForward` -- ``jacobian @ (1 / vel[nearest_site(grid_points)])``
  (docs/source/tutorials/31_sw_tomography.ipynb cell 12, tests/21_test_voronoi_2d.py:57-68);
* :class:`NearestLookup1D` -- ``Voronoi1D.interpolate_tessellation(sites, values, x, 'nuclei')``
  (discretization/_voronoi.py:705-732; the lookup of 42_transd_partition_mod.ipynb cell 11);
* :class:`LinearForward` -- ``G @ m`` (02_hierarchical_polyfit.ipynb cell 12);
* :class:`PolynomialForward` -- ``np.vander(x, len(m), True) @ m`` on a trans-dimensional ``ParameterSpace``
  (03_transd_polyfit.ipynb cell 5).

They are listed where the reference lists forward callables, and they ARE valid reference forwards: called on
the host with a ``State`` (of this package or of the unmodified reference) they evaluate exactly what the
tutorials' forward functions evaluate, with scipy / numpy (``Callable[[State], np.ndarray]``,
likelihood/_log_likelihood.py:67-78, 325-331).  That is how the same problem object is run inside the reference
(``bayesbay_b200.plugin``, the parity tests, the reference arm of ``bench.py``).  The CUDA engine never calls
them -- it reads the arrays they carry -- so this is not a fallback of the device path.
"""
from __future__ import annotations

import numpy as np

__all__ = ["ForwardOperator", "RayTomographyForward", "NearestLookup1D", "LinearForward", "PolynomialForward"]


class ForwardOperator:
    kind = None

    def __call__(self, state):  # pragma: no cover - abstract
        raise NotImplementedError

    @property
    def __name__(self):
        return type(self).__name__


class RayTomographyForward(ForwardOperator):
    """Straight-ray travel times through a 2-D Voronoi field.

    Parameters
    ----------
    jacobian : scipy.sparse matrix or ``(indptr, indices, data)``
        ray matrix, rows = rays, columns = ``grid_points`` (shared by all chains)
    grid_points : (G, 2) array
        query points of the nearest-site rasterisation
    param_space_name, param_name : str
        the Voronoi2D space and the parameter holding the cell values
    transform : {"reciprocal", "identity"}
        ``"reciprocal"`` multiplies the ray matrix with ``1 / value`` (velocity -> slowness)
    save_field_as : str, optional
        the tutorial's forward stores the rasterised field with ``state.save_to_extra_storage('interp_vel', ...)``
        so that it shows up in ``get_results()``; give the key (e.g. ``"interp_vel"``) to get the same entry.
        Running sums for the posterior mean / standard deviation map are kept on the device either way
        (``BayesianInversion.get_field_statistics``).
    store_field_samples : bool
        keep every saved field (G doubles per chain per save point, as the reference does) in the results;
        switch off for large runs and use the on-device statistics instead
    """

    kind = "ray2d"

    def __init__(self, jacobian, grid_points, param_space_name="voronoi", param_name="vel", transform="reciprocal",
                 save_field_as=None, store_field_samples=True):
        if isinstance(jacobian, (tuple, list)):
            indptr, indices, data = jacobian
        else:
            csr = jacobian.tocsr()
            csr.sort_indices()
            indptr, indices, data = csr.indptr, csr.indices, csr.data
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.data = np.ascontiguousarray(data, dtype=np.float64)
        self.grid_points = np.ascontiguousarray(grid_points, dtype=np.float64)
        if self.grid_points.ndim != 2 or self.grid_points.shape[1] != 2:
            raise ValueError("grid_points should have shape (G, 2)")
        if transform not in ("reciprocal", "identity"):
            raise ValueError("transform should be 'reciprocal' or 'identity'")
        if self.indices.size and (self.indices.min() < 0 or self.indices.max() >= self.grid_points.shape[0]):
            raise ValueError("jacobian column index outside grid_points")
        if self.indptr[0] != 0 or self.indptr[-1] != self.indices.size or self.data.size != self.indices.size:
            raise ValueError("inconsistent CSR arrays")
        self.param_space_name = param_space_name
        self.param_name = param_name
        self.transform = transform
        self.save_field_as = save_field_as
        self.store_field_samples = bool(store_field_samples)

    @property
    def n_data(self):
        return self.indptr.size - 1

    def __call__(self, state):
        """Host evaluation on one ``State`` (31_sw_tomography.ipynb cell 12): nearest site of every grid point with
        the space's cached KD-tree when there is one, then ``jacobian @ field``."""
        import scipy.sparse
        import scipy.spatial

        if getattr(self, "_jacobian", None) is None:
            self._jacobian = scipy.sparse.csr_matrix((self.data, self.indices, self.indptr),
                                                     shape=(self.indptr.size - 1, self.grid_points.shape[0]))
        ps = state[self.param_space_name]
        tree = None
        if hasattr(ps, "saved_in_cache") and ps.saved_in_cache("kdtree"):
            tree = ps.load_from_cache("kdtree")
        if tree is None:
            tree = scipy.spatial.KDTree(ps["discretization"])
        field = ps[self.param_name][tree.query(self.grid_points)[1]]
        if self.save_field_as is not None and hasattr(state, "save_to_extra_storage"):
            state.save_to_extra_storage(self.save_field_as, field)
        return self._jacobian @ (1 / field if self.transform == "reciprocal" else field)

    def __getstate__(self):  # (the cached scipy matrix is rebuilt where the object lands)
        d = dict(self.__dict__)
        d.pop("_jacobian", None)
        return d


class NearestLookup1D(ForwardOperator):
    """``d_pred[i] = values[nearest_site(x[i])]`` on a Voronoi1D space."""

    kind = "lookup1d"

    def __init__(self, x, param_space_name, param_name):
        self.x = np.ascontiguousarray(x, dtype=np.float64).ravel()
        self.param_space_name = param_space_name
        self.param_name = param_name

    @property
    def n_data(self):
        return self.x.size

    def __call__(self, state):
        """``Voronoi1D.interpolate_tessellation(sites, values, x, 'nuclei')`` (discretization/_voronoi.py:705-732;
        nearest_neighbour_1d, _utils_1d.pyx:102-114): below the first site -> cell 0, past the last -> the last,
        between two sites the nearer one, the upper one on a tie."""
        ps = state[self.param_space_name]
        sites = np.asarray(ps["discretization"], dtype=np.float64)
        values = np.asarray(ps[self.param_name])
        k = sites.size
        if k == 1:
            return np.full(self.x.shape, values[0])
        hi = np.clip(np.searchsorted(sites, self.x, side="left"), 1, k - 1)  # first site >= x (clipped): x in [lo, hi]
        lo = hi - 1
        idx = np.where(np.abs(sites[lo] - self.x) < np.abs(sites[hi] - self.x), lo, hi)
        idx = np.where(self.x < sites[0], 0, idx)
        return values[idx]


class LinearForward(ForwardOperator):
    """``d_pred = G @ concatenate([state[ps][p] for p in param_names])`` on a fixed-dimension space."""

    kind = "linear"

    def __init__(self, G, param_space_name, param_names):
        self.G = np.ascontiguousarray(G, dtype=np.float64)
        if self.G.ndim != 2:
            raise ValueError("G should be a 2-D array")
        self.param_space_name = param_space_name
        self.param_names = list(param_names)

    @property
    def n_data(self):
        return self.G.shape[0]

    def __call__(self, state):  # 02_hierarchical_polyfit.ipynb cell 12
        ps = state[self.param_space_name]
        return self.G @ np.concatenate([np.atleast_1d(ps[n]) for n in self.param_names])


class PolynomialForward(ForwardOperator):
    """``d_pred = np.vander(x, k, increasing=True) @ state[ps][param]`` with ``k`` the current number of
    dimensions of a (trans-dimensional) ``ParameterSpace``: the forward of the trans-d polynomial regression
    tutorial."""

    kind = "polynomial"

    def __init__(self, x, param_space_name, param_name):
        self.x = np.ascontiguousarray(x, dtype=np.float64).ravel()
        self.param_space_name = param_space_name
        self.param_name = param_name

    @property
    def n_data(self):
        return self.x.size

    def __call__(self, state):  # 03_transd_polyfit.ipynb cell 5
        m = np.asarray(state[self.param_space_name][self.param_name])
        return np.vander(self.x, len(m), True) @ m
