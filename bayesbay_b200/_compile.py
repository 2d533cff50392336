"""Compile API objects into the flat problem description ("spec dict") the device engine runs.

Duck-typed on attribute names the reference classes share with this package (``_n_dimensions_min``,
``parameters``, ``birth_from``, ``vmin`` ...), so the same function also accepts a
``Parameterization`` / ``LogLikelihood`` built with the unmodified reference -- that is what lets a
batched sampler plug into the reference's own ``BayesianInversion.run(sampler=...)`` seam
(INTEGRATION.md)."""
from __future__ import annotations

from numbers import Number

import numpy as np

from .exceptions import DeviceUnsupported

MAX_SPACES, MAX_PARAMS, MAX_TARGETS = 4, 8, 4


def _cls(obj):
    return type(obj).__name__


def _scalar(x, what):
    if x is None or not np.isscalar(x) or not isinstance(x, Number):
        raise DeviceUnsupported(f"{what} must be a scalar (position-dependent hyper-parameters are not on the device yet)")
    return float(x)


def compile_prior(p) -> dict:
    """Prior -> dict(name, kind, a, b, perturb_std, perturb_std_birth[, table]).  Scalar hyper-parameters, or --
    with a 1-D ``position`` -- a table of per-position hyper-parameters the device interpolates linearly."""
    name = _cls(p)
    if name in ("UniformPrior", "UniformParameter"):
        kind, a_name, b_name = "uniform", "vmin", "vmax"
    elif name in ("GaussianPrior", "GaussianParameter"):
        kind, a_name, b_name = "gaussian", "mean", "std"
    elif name == "LaplacePrior":
        kind, a_name, b_name = "laplace", "mean", "scale"
    else:
        raise DeviceUnsupported(
            f"prior {p.name!r} of type {name}: only Uniform/Gaussian/LaplacePrior run on the device "
            "(user callables cannot run in a CUDA kernel and there is no CPU fallback)")

    def raw(attr):
        # the reference stores position-dependent hyper-parameters as interpolating partials; the raw arrays are in
        # its _repr_args (prior/_prior.py:35-40)
        v = getattr(p, attr, None)
        if callable(v) or v is None:
            v = getattr(p, "_repr_args", {}).get(attr)
        return v

    std = raw("perturb_std")
    std_birth = raw("perturb_std_birth")
    std_birth = std if std_birth is None else std_birth
    a, b = raw(a_name), raw(b_name)
    position = getattr(p, "position", None)
    if position is None:
        return dict(name=p.name, kind=kind, a=_scalar(a, a_name), b=_scalar(b, b_name),
                    perturb_std=_scalar(std, "perturb_std"), perturb_std_birth=_scalar(std_birth, "perturb_std_birth"))
    pos = np.asarray(position, dtype=np.float64)
    if pos.ndim != 1:
        raise DeviceUnsupported(f"prior {p.name!r}: only 1-D `position` arrays are supported on the device")
    col = lambda v: np.broadcast_to(np.asarray(v, dtype=np.float64), pos.shape).copy()  # noqa: E731
    return dict(name=p.name, kind=kind, a=None, b=None, perturb_std=None, perturb_std_birth=None,
                table=dict(position=pos.copy(), a=col(a), b=col(b), perturb_std=col(std), perturb_std_birth=col(std_birth)))


def compile_space(ps) -> dict:
    name = _cls(ps)
    params = list(ps.parameters.values())
    for p in params:
        if hasattr(p, "parameters"):
            raise DeviceUnsupported(f"parameter space {ps.name!r}: nested parameter spaces are not supported on the device")
    if len(params) > MAX_PARAMS:
        raise DeviceUnsupported(f"parameter space {ps.name!r}: more than {MAX_PARAMS} parameters")
    trans_d = bool(ps.trans_d)
    kmin, kmax = int(ps._n_dimensions_min), int(ps._n_dimensions_max)
    init_range = ps._n_dimensions_init_range
    kinit_max = int((kmax - kmin) * init_range + kmin) if trans_d else kmin
    kinit_max = min(max(kinit_max, kmin), kmax)
    spec = dict(name=ps.name, trans_d=trans_d, kmin=kmin, kmax=kmax, kinit_max=kinit_max,
                params=[compile_prior(p) for p in params])
    if name != "Voronoi1D" and any(q.get("table") is not None for q in spec["params"]):
        raise DeviceUnsupported(f"parameter space {ps.name!r}: position-dependent priors need a Voronoi1D space")
    if name in ("Voronoi1D", "Voronoi2D"):
        ndim = 1 if name == "Voronoi1D" else 2
        if getattr(ps, "polygon", None) is not None:
            raise DeviceUnsupported("polygon domains are not supported on the device")
        vmin = np.atleast_1d(np.asarray(ps.vmin, dtype=float))
        vmax = np.atleast_1d(np.asarray(ps.vmax, dtype=float))
        std = np.atleast_1d(np.asarray(ps.perturb_std, dtype=float))
        if std.size == 1:
            std = np.repeat(std, ndim)
        if vmin.size != ndim or vmax.size != ndim or std.size != ndim:
            raise ValueError(f"{ps.name}: vmin/vmax/perturb_std should have {ndim} entries")
        if ps.birth_from not in ("prior", "neighbour"):
            raise ValueError("birth_from should be 'prior' or 'neighbour'")
        spec.update(kind="voronoi1d" if ndim == 1 else "voronoi2d", ndim=ndim, vmin=[float(v) for v in vmin],
                    vmax=[float(v) for v in vmax], site_std=[float(v) for v in std], birth_from=ps.birth_from,
                    weights=dict(birth=1 if trans_d else 0, death=1 if trans_d else 0, values=3 if params else 0, sites=1))
    elif name == "ParameterSpace":
        spec.update(kind="plain", ndim=0, vmin=[], vmax=[], site_std=[], birth_from="prior",
                    weights=dict(birth=1 if trans_d else 0, death=1 if trans_d else 0, values=3 if params else 0, sites=0))
        if not trans_d and not params:
            raise DeviceUnsupported(f"parameter space {ps.name!r} has nothing to perturb")
    else:
        raise DeviceUnsupported(f"parameter space type {name} is not supported on the device")
    return spec


def compile_target(tgt, fwd, space_index: dict, spaces: list) -> dict:
    dobs = np.ascontiguousarray(tgt.dobs, dtype=np.float64).ravel()
    cov = getattr(tgt, "covariance_mat_inv", None)
    spec = dict(name=tgt.name, dobs=dobs, hierarchical=cov is None,
                std_min=float(tgt.std_min), std_max=float(tgt.std_max), std_perturb_std=float(tgt.std_perturb_std),
                cov_inv=None, correlated=bool(getattr(tgt, "noise_is_correlated", False)) and cov is None,
                corr_min=float(tgt.correlation_min), corr_max=float(tgt.correlation_max),
                corr_perturb_std=float(tgt.correlation_perturb_std))
    if cov is not None:
        if np.isscalar(cov):
            spec["cov_inv"] = float(cov)
        else:
            cov = np.asarray(cov, dtype=np.float64)
            if not ((cov.ndim == 1 and cov.size == dobs.size) or (cov.ndim == 2 and cov.shape == (dobs.size, dobs.size))):
                raise ValueError(f"target {tgt.name!r}: covariance_mat_inv must be a scalar, a vector [{dobs.size}] or a "
                                 f"matrix [{dobs.size}][{dobs.size}], got shape {cov.shape}")
            spec["cov_inv"] = cov
    kind = getattr(fwd, "kind", None)
    if kind not in ("ray2d", "lookup1d", "linear", "polynomial"):
        raise TypeError(
            f"forward function of target {tgt.name!r} is {fwd!r}: the device engine only runs the forward operators "
            "of bayesbay_b200.forward (RayTomographyForward, NearestLookup1D, LinearForward, PolynomialForward). Arbitrary Python "
            "callables cannot run in a CUDA kernel and there is no CPU fallback.")
    if fwd.param_space_name not in space_index:
        raise ValueError(f"forward of target {tgt.name!r} reads unknown parameter space {fwd.param_space_name!r}")
    s = space_index[fwd.param_space_name]
    sp = spaces[s]
    pnames = [p["name"] for p in sp["params"]]
    if fwd.n_data != dobs.size:
        raise ValueError(f"forward of target {tgt.name!r} predicts {fwd.n_data} data but dobs has {dobs.size}")
    if kind == "ray2d":
        if sp["kind"] != "voronoi2d":
            raise ValueError("RayTomographyForward needs a Voronoi2D parameter space")
        spec["forward"] = dict(kind="ray2d", space=s, param=pnames.index(fwd.param_name), transform=fwd.transform,
                               indptr=fwd.indptr, indices=fwd.indices, data=fwd.data, points=fwd.grid_points)
        if getattr(fwd, "save_field_as", None):
            spec["forward"].update(save_field_as=fwd.save_field_as, store_field_samples=fwd.store_field_samples)
    elif kind == "lookup1d":
        if sp["kind"] != "voronoi1d":
            raise ValueError("NearestLookup1D needs a Voronoi1D parameter space")
        spec["forward"] = dict(kind="lookup1d", space=s, param=pnames.index(fwd.param_name), x=fwd.x)
    elif kind == "polynomial":
        if sp["kind"] != "plain":
            raise ValueError("PolynomialForward needs a plain ParameterSpace")
        spec["forward"] = dict(kind="polynomial", space=s, param=pnames.index(fwd.param_name), x=fwd.x)
    else:
        if sp["trans_d"]:
            raise ValueError("LinearForward needs a fixed-dimension parameter space")
        params = [pnames.index(n) for n in fwd.param_names]
        if fwd.G.shape[1] != len(params) * sp["kmax"]:
            raise ValueError(f"LinearForward: G has {fwd.G.shape[1]} columns, expected {len(params) * sp['kmax']}")
        spec["forward"] = dict(kind="linear", space=s, params=params, G=fwd.G)
    return spec


def compile_problem(parameterization, log_likelihood=None) -> dict:
    """``Parameterization`` (+ ``LogLikelihood``) -> spec dict.  Raises ``TypeError`` /
    ``DeviceUnsupported`` for anything the device cannot run (there is no CPU fallback)."""
    spaces = [compile_space(ps) for ps in parameterization.parameter_spaces.values()]
    if len(spaces) > MAX_SPACES:
        raise DeviceUnsupported(f"more than {MAX_SPACES} parameter spaces")
    index = {sp["name"]: i for i, sp in enumerate(spaces)}
    targets = []
    if log_likelihood is not None:
        if getattr(log_likelihood, "log_like_func", None) or getattr(log_likelihood, "log_like_ratio_func", None):
            raise TypeError("user log-likelihood callables cannot run on the device; use targets + forward operators")
        tg = list(log_likelihood.targets)
        fw = list(log_likelihood.fwd_functions)
        if len(tg) != len(fw):
            raise ValueError("the number of targets and forward functions must be equal")
        if len(tg) > MAX_TARGETS:
            raise DeviceUnsupported(f"more than {MAX_TARGETS} targets")
        for t, f in zip(tg, fw):
            f = getattr(f, "func", f)  # the reference wraps callables in _FunctionWrapper(func, args, kwargs)
            targets.append(compile_target(t, f, index, spaces))
    return dict(spaces=spaces, targets=targets)


def perturbation_names(spec: dict):
    """Statistics keys in device move-id order: ``[(outer_name, inner_name | None)] * (4*n_spaces+1)``
    (None where the move does not exist), following the ``__name__`` rules of perturbations/*.py."""
    out = []
    for sp in spec["spaces"]:
        outer = f"ParamSpacePerturbation(param_space_name={sp['name']})"
        pn = [f"{sp['name']}.{p['name']}" for p in sp["params"]]
        names = {
            "birth": f"BirthPerturbation({sp['name']})",
            "death": f"DeathPerturbation({sp['name']})",
            "values": f"ParamPerturbation({str(pn) if len(pn) != 1 else pn[0]})",
            "sites": f"ParamPerturbation({sp['name']}.discretization)",
        }
        for kind in ("birth", "death", "values", "sites"):
            if sp["weights"][kind] <= 0:
                out.append(None)
            elif sp.get("flat"):  # the inner perturbations sit in the chain's own list (set_perturbation_funcs)
                out.append((names[kind], None))
            else:
                out.append((outer, names[kind]))
    hier = [t["name"] for t in spec["targets"] if t["hierarchical"]]
    out.append((f"NoisePerturbation({str(hier) if len(hier) > 1 else hier[0]})", None)
               if hier and spec.get("w_noise") != 0 else None)
    return out
