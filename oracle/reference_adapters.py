"""Turn a device forward-operator descriptor (``bayesbay_b200.forward``) into the equivalent plain
``state -> d_pred`` callable, so the SAME problem can be run inside the unmodified reference.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): used to generate golden vectors and to validate; the
product never calls these."""
from __future__ import annotations

import numpy as np

from . import kernels as K


def as_reference_forward(op):
    kind = getattr(op, "kind", None)
    if kind == "ray2d":
        import scipy.sparse
        import scipy.spatial

        jac = scipy.sparse.csr_matrix((op.data, op.indices, op.indptr), shape=(op.indptr.size - 1, op.grid_points.shape[0]))

        def forward(state):  # docs/source/tutorials/31_sw_tomography.ipynb cell 12
            ps = state[op.param_space_name]
            if ps.saved_in_cache("kdtree"):
                tree = ps.load_from_cache("kdtree")
            else:
                tree = scipy.spatial.KDTree(ps["discretization"])
            field = ps[op.param_name][tree.query(op.grid_points)[1]]
            return jac @ (1 / field if op.transform == "reciprocal" else field)

        return forward
    if kind == "lookup1d":
        def forward(state):  # Voronoi1D.interpolate_tessellation(..., 'nuclei'), discretization/_voronoi.py:705-732
            ps = state[op.param_space_name]
            return K.lookup1d_forward(op.x, ps["discretization"], ps[op.param_name])

        return forward
    if kind == "linear":
        def forward(state):  # docs/source/tutorials/02_hierarchical_polyfit.ipynb cell 12
            ps = state[op.param_space_name]
            return op.G @ np.concatenate([ps[n] for n in op.param_names])

        return forward
    if kind == "polynomial":
        def forward(state):  # docs/source/tutorials/03_transd_polyfit.ipynb cell 5
            m = state[op.param_space_name][op.param_name]
            return np.vander(op.x, len(m), True) @ m

        return forward
    raise TypeError(f"not a device forward operator: {op!r}")
