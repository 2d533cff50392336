"""Voronoi discretisation descriptors (reference src/bayesbay/discretization/_voronoi.py,
_discretization.py).  Site sampling / moves / birth / death / nearest-site search run on the device;
the tessellation utilities below (``interpolate_tessellation``, ``get_tessellation_statistics``,
``compute_cell_extents``) are batched onto the CUDA rasterisation kernels."""
from __future__ import annotations

from numbers import Number
from typing import List, Union

import numpy as np

from .exceptions import DeviceUnsupported
from .parameterization import ParameterSpace
from .perturbations import ParamPerturbation, ParamSpacePerturbation
from .prior import Prior

__all__ = ["Discretization", "Voronoi", "Voronoi1D", "Voronoi2D"]


class Discretization(ParameterSpace):
    """Base class of discretisations (reference discretization/_discretization.py:15-510)."""

    def __init__(self, name, spatial_dimensions, perturb_std=0.1, n_dimensions=None, n_dimensions_min=1,
                 n_dimensions_max=10, n_dimensions_init_range=0.3, parameters=None, birth_from="prior", **kwargs):
        if birth_from not in ("prior", "neighbour"):
            raise ValueError("`birth_from` should be either 'prior' or 'neighbour'")
        self.spatial_dimensions = spatial_dimensions
        self.perturb_std = perturb_std
        self.birth_from = birth_from
        self.position = None
        super().__init__(name=name, n_dimensions=n_dimensions, n_dimensions_min=n_dimensions_min,
                         n_dimensions_max=n_dimensions_max, n_dimensions_init_range=n_dimensions_init_range,
                         parameters=parameters)
        self._repr_args.update(spatial_dimensions=spatial_dimensions, perturb_std=perturb_std, **kwargs)
        if self.trans_d:
            self._repr_args["birth_from"] = birth_from

    def get_perturb_std(self, *args) -> Number:
        return self.perturb_std


class Voronoi(Discretization):
    """Voronoi tessellation (reference discretization/_voronoi.py:39-468)."""

    def __init__(self, name, spatial_dimensions, vmin, vmax, perturb_std, n_dimensions=None, n_dimensions_min=2,
                 n_dimensions_max=10, n_dimensions_init_range=0.3, parameters=None, birth_from="neighbour"):
        if type(self) is Voronoi:
            if spatial_dimensions <= 2:
                sub = {1: "Voronoi1D", 2: "Voronoi2D"}[spatial_dimensions]
                raise ValueError(f"Use {sub} for {spatial_dimensions}D tessellations")
            raise DeviceUnsupported("only 1-D and 2-D Voronoi tessellations are supported by the device engine")
        self.vmin = vmin
        self.vmax = vmax
        if n_dimensions is not None:
            assert isinstance(n_dimensions, int) and n_dimensions > 0, "`n_dimensions` should be a positive integer"
        super().__init__(name=name, spatial_dimensions=spatial_dimensions, perturb_std=perturb_std,
                         n_dimensions=n_dimensions, n_dimensions_min=n_dimensions_min,
                         n_dimensions_max=n_dimensions_max, n_dimensions_init_range=n_dimensions_init_range,
                         parameters=parameters, birth_from=birth_from, vmin=vmin, vmax=vmax)

    def _init_perturbation_funcs(self):
        # birth 1, death 1, values 3, site move 1 (reference _voronoi.py:402-430)
        funcs, weights = self._inner_perturbations()
        funcs.append(ParamPerturbation(self.name, [self]))
        weights.append(1)
        self._perturbation_funcs = [ParamSpacePerturbation(self.name, funcs, weights)]
        self._perturbation_weights = [sum(weights)]

    def log_prior(self, *args):
        raise NotImplementedError  # as in the reference (_voronoi.py:387-400)


class Voronoi1D(Voronoi):
    """``Voronoi1D(name, vmin, vmax, perturb_std, n_dimensions=None, n_dimensions_min=1,
    n_dimensions_max=10, n_dimensions_init_range=0.3, parameters=None, birth_from='neighbour')``
    (reference _voronoi.py:471-1280).  Sites are kept sorted."""

    def __init__(self, name, vmin, vmax, perturb_std, n_dimensions=None, n_dimensions_min=1, n_dimensions_max=10,
                 n_dimensions_init_range=0.3, parameters: List[Prior] = None, birth_from="neighbour"):
        super().__init__(name=name, spatial_dimensions=1, vmin=vmin, vmax=vmax, perturb_std=perturb_std,
                         n_dimensions=n_dimensions, n_dimensions_min=n_dimensions_min,
                         n_dimensions_max=n_dimensions_max, n_dimensions_init_range=n_dimensions_init_range,
                         parameters=parameters, birth_from=birth_from)

    @staticmethod
    def compute_cell_extents(voronoi_sites: np.ndarray, lb=0, ub=None, fill_value=0):
        """Cell extents from midpoints between consecutive sites (reference _voronoi.py:630-670 ->
        _utils_1d.pyx:30-52).  Unbounded cells get ``fill_value``."""
        s = np.asarray(voronoi_sites, dtype=np.float64)
        mids = (s[:-1] + s[1:]) / 2.0
        lo = np.concatenate(([np.nan if lb is None else float(lb)], mids))
        hi = np.concatenate((mids, [np.nan if ub is None else float(ub)]))
        ext = hi - lo
        if lb is None:
            ext[0] = fill_value
        if ub is None:
            ext[-1] = fill_value
        return ext

    @staticmethod
    def compute_interface_positions(voronoi_cells: np.ndarray, input_type="nuclei", lb_tessellation=None):
        if input_type == "nuclei":
            return (voronoi_cells[:-1] + voronoi_cells[1:]) / 2
        if input_type == "extents":
            assert lb_tessellation is not None, "`lb_tessellation` should not be None when `input_type` is'extents'"
            return np.cumsum(voronoi_cells)[:-1] + lb_tessellation
        raise ValueError("`input_type` should either be 'nuclei' or 'extents'")

    @staticmethod
    def interpolate_tessellation(voronoi_cells, param_values, interp_positions, input_type="nuclei"):
        """Reference _voronoi.py:705-732.  ``'nuclei'``: nearest-nuclei interpolation on the device
        (``interpolate_nearest_1d``, _utils_1d.pyx:171-192).  ``'extents'``: the layer whose cumulative extent interval
        ``[sum(extents[:i]), sum(extents[:i+1]))`` holds the position (``interpolate_depth_profile``,
        _utils_1d.pyx:119-140) -- a results-side utility on one host array, evaluated on the host with the reference's
        sweep semantics (the layer pointer never moves back; positions past the last interface, and everything after a
        position below zero, fall in the last layer)."""
        if input_type == "extents":
            extents = np.asarray(voronoi_cells, dtype=np.float64)
            values = np.asarray(param_values, dtype=np.float64)
            x0 = np.atleast_1d(np.asarray(interp_positions, dtype=np.float64))
            upper = np.cumsum(extents)          # running sums, accumulated in the reference's order
            lower = np.concatenate(([0.0], upper[:-1]))
            out = np.empty(x0.shape, dtype=np.float64)
            layer, last = 0, extents.size - 1
            for i, x in enumerate(x0):
                while not (lower[layer] <= x < upper[layer]) and layer < last:
                    layer += 1
                out[i] = values[layer]
            return out
        if input_type != "nuclei":
            raise ValueError("`input_type` should either be 'nuclei' or 'extents'")
        from ._batch import device_interpolate_1d

        return device_interpolate_1d([np.asarray(voronoi_cells)], [np.asarray(param_values)], interp_positions)[0]


class Voronoi2D(Voronoi):
    """``Voronoi2D(name, vmin=None, vmax=None, polygon=None, perturb_std=1, n_dimensions=None,
    n_dimensions_min=2, n_dimensions_max=100, n_dimensions_init_range=0.3, parameters=None,
    birth_from='neighbour', compute_kdtree=False)`` (reference _voronoi.py:1283-1630).

    ``compute_kdtree`` is accepted and ignored: the device keeps the nearest-site index of every
    grid point resident instead of a KD-tree."""

    def __init__(self, name, vmin=None, vmax=None, polygon=None, perturb_std: Union[Number, np.ndarray] = 1,
                 n_dimensions=None, n_dimensions_min=2, n_dimensions_max=100, n_dimensions_init_range=0.3,
                 parameters: List[Prior] = None, birth_from="neighbour", compute_kdtree=False):
        if polygon is not None:
            raise DeviceUnsupported("polygon domains (shapely) are not supported by the device engine")
        assert vmin is not None and vmax is not None, (
            "Either `vmin`/`vmax` or `polygon` must not be None to properly define the discretization domain.")
        self.polygon = None
        self.compute_kdtree = compute_kdtree
        super().__init__(name=name, spatial_dimensions=2, vmin=vmin, vmax=vmax, perturb_std=perturb_std,
                         n_dimensions=n_dimensions, n_dimensions_min=n_dimensions_min,
                         n_dimensions_max=n_dimensions_max, n_dimensions_init_range=n_dimensions_init_range,
                         parameters=parameters, birth_from=birth_from)

    @staticmethod
    def interpolate_tessellation(voronoi_sites, param_values, interp_positions):
        """Nearest-neighbour interpolation (reference _voronoi.py:1429-1454), CUDA rasteriser."""
        return Voronoi2D._interpolate_tessellations([voronoi_sites], [param_values], interp_positions)[0]

    @staticmethod
    def _interpolate_tessellations(samples_voronoi_sites, samples_param_values, interp_positions):
        from ._batch import device_interpolate_2d

        return device_interpolate_2d(samples_voronoi_sites, samples_param_values, interp_positions)

    @staticmethod
    def get_tessellation_statistics(samples_voronoi_cells, samples_param_values, interp_positions, percentiles=(10, 90)):
        """Mean / median / std / percentiles of an ensemble of tessellations (reference
        _voronoi.py:1465-1500); all samples are rasterised in one batched kernel launch."""
        import torch

        from ._batch import device_interpolate_2d

        fields = device_interpolate_2d(samples_voronoi_cells, samples_param_values, interp_positions, as_tensor=True)
        q = torch.tensor([p / 100.0 for p in percentiles], dtype=torch.float64, device=fields.device)
        return {
            "mean": fields.mean(dim=0).cpu().numpy(),
            "median": torch.quantile(fields, 0.5, dim=0).cpu().numpy(),
            "std": fields.std(dim=0, unbiased=False).cpu().numpy(),
            "percentiles": torch.quantile(fields, q, dim=0).cpu().numpy(),
        }
