"""Parity of the stand-alone CUDA kernels (through the C ABI) against the golden fixtures recorded
from the reference and against the oracle on seeded random states.

Bar: bit-exact for indices; 1e-12 relative for fp64 results (summation order differs)."""
import ctypes as C
import math
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import kernels as OK  # noqa: E402
from tests import problems  # noqa: E402

RTOL = 1e-12


@pytest.fixture(scope="module")
def lib():
    from bayesbay_b200 import _lib

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return _lib.load()


@pytest.fixture(scope="module")
def kg(golden_dir):
    return np.load(os.path.join(golden_dir, "kernels.npz"))


_KEEP = []  # device inputs must outlive the asynchronous launches that read them


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a, dtype=dtype)).cuda()
    _KEEP.append(t)
    if len(_KEEP) > 256:
        torch.cuda.synchronize()
        del _KEEP[:128]
    return t


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def chk(rc):
    from bayesbay_b200 import _lib

    _lib.check(rc)


def pack_sites_2d(sites_list, K):
    Cn = len(sites_list)
    s = np.zeros((2, K, Cn))
    k = np.zeros(Cn, np.int32)
    for c, a in enumerate(sites_list):
        k[c] = a.shape[0]
        s[:, : k[c], c] = a.T
    return s, k


def raster2d(lib, sites_list, K, pts):
    s, k = pack_sites_2d(sites_list, K)
    Cn, G = len(sites_list), pts.shape[0]
    out = torch.zeros((G, Cn), dtype=torch.uint8 if K <= 256 else torch.int16, device="cuda")
    chk(lib.bbb_raster2d(ptr(dev(s)), ptr(dev(k)), K, Cn, ptr(dev(pts[:, 0])), ptr(dev(pts[:, 1])), G, ptr(out), None))
    torch.cuda.synchronize()
    return out.cpu().numpy().astype(np.int64) & (0xFF if K <= 256 else 0xFFFF)


def test_raster2d_golden_kdtree(lib, kg):
    """F1: bit-exact against scipy.spatial.KDTree.query as called by Voronoi2D.interpolate_tessellation."""
    pts = kg["r2_points"]
    sites_list = [kg[f"r2_sites_{i}"] for i in range(len(kg["r2_ks"]))]
    for K in (200, 256, 300):
        idx = raster2d(lib, sites_list, K, pts)
        for i in range(len(sites_list)):
            assert np.array_equal(idx[:, i], kg[f"r2_idx_{i}"]), (K, i)


def test_raster2d_random_ragged(lib):
    rng = np.random.default_rng(5)
    pts = rng.uniform(-1, 1, (777, 2))  # arbitrary (non-grid) query points
    ks = rng.integers(1, 64, 70)
    sites_list = [rng.uniform(-1, 1, (k, 2)) for k in ks]
    idx = raster2d(lib, sites_list, 64, pts)
    for c, s in enumerate(sites_list):
        assert np.array_equal(idx[:, c], OK.nearest_site_2d(s, pts)), c
    import scipy.spatial

    for c in (0, 13, 69):
        assert np.array_equal(idx[:, c], scipy.spatial.KDTree(sites_list[c]).query(pts)[1])


def test_raster2d_ties_lowest_index(lib):
    # two coincident sites and a point equidistant from two sites: lowest index wins (numpy/KDTree rule)
    sites = np.array([[0.5, 0.0], [-0.5, 0.0], [0.5, 0.0]])
    pts = np.array([[0.0, 0.3], [0.5, 0.1], [-0.4, 0.0]])
    idx = raster2d(lib, [sites], 8, pts)
    assert idx[:, 0].tolist() == [0, 0, 1]
    assert np.array_equal(idx[:, 0], OK.nearest_site_2d(sites, pts))


def test_raster1d_golden(lib, kg):
    """F1/F4: interpolate_nearest_1d / nearest-nuclei lookup, incl. exact ties and out-of-range points."""
    x = kg["r1_x"]
    n = len(kg["r1_ks"])
    K = 40
    sites = np.zeros((K, n + 1))
    vals = np.zeros((K, n + 1))
    k = np.zeros(n + 1, np.int32)
    for i in range(n):
        k[i] = kg["r1_ks"][i]
        sites[: k[i], i] = kg[f"r1_sites_{i}"]
        vals[: k[i], i] = kg[f"r1_vals_{i}"]
    k[n] = 4
    sites[:4, n] = kg["r1_tie_sites"]
    vals[:4, n] = np.arange(4.0)
    for xs, order in ((x, "asc"), (x[::-1].copy(), "desc")):
        R = xs.size
        idx = torch.zeros((R, n + 1), dtype=torch.int32, device="cuda")
        field = torch.zeros((R, n + 1), dtype=torch.float64, device="cuda")
        chk(lib.bbb_raster1d(ptr(dev(sites)), ptr(dev(vals)), ptr(dev(k)), K, n + 1, ptr(dev(xs)), R, ptr(idx), ptr(field), None))
        f = field.cpu().numpy()
        if order == "desc":
            f = f[::-1]
        for i in range(n):
            assert np.array_equal(f[:, i], kg[f"r1_field_{i}"]), (order, i)
    xt = kg["r1_tie_x"]
    idx = torch.zeros((xt.size, n + 1), dtype=torch.int32, device="cuda")
    field = torch.zeros((xt.size, n + 1), dtype=torch.float64, device="cuda")
    chk(lib.bbb_raster1d(ptr(dev(sites)), ptr(dev(vals)), ptr(dev(k)), K, n + 1, ptr(dev(xt)), xt.size, ptr(idx), ptr(field), None))
    assert np.array_equal(field.cpu().numpy()[:, n], kg["r1_tie_field"])
    assert np.array_equal(idx.cpu().numpy()[:, n], OK.nearest_site_1d(kg["r1_tie_sites"], xt))


def test_ray_forward_vs_oracle_and_golden(lib, kg):
    """F2: jacobian @ (1 / vel[nearest]) and the fused sum of squared residuals."""
    ing = problems.tomo_small()
    pts, indptr, indices, data = ing["points"], ing["indptr"], ing["indices"], ing["data"]
    G, R = pts.shape[0], indptr.size - 1
    rng = np.random.default_rng(3)
    sites_list = [kg["fw_sites"]] + [rng.uniform(-1, 1, (k, 2)) for k in rng.integers(1, 20, 40)]
    vals_list = [kg["fw_vel"]] + [rng.uniform(2, 4, s.shape[0]) for s in sites_list[1:]]
    K, Cn = 20, len(sites_list)
    s, k = pack_sites_2d(sites_list, K)
    v = np.zeros((K, Cn))
    for c, a in enumerate(vals_list):
        v[: a.size, c] = a
    idx = torch.zeros((G, Cn), dtype=torch.uint8, device="cuda")
    chk(lib.bbb_raster2d(ptr(dev(s)), ptr(dev(k)), K, Cn, ptr(dev(pts[:, 0])), ptr(dev(pts[:, 1])), G, ptr(idx), None))
    dpred = torch.zeros((R, Cn), dtype=torch.float64, device="cuda")
    phi = torch.zeros((Cn,), dtype=torch.float64, device="cuda")
    tiles = int(lib.bbb_ray_tiles(R))
    scratch = torch.zeros((tiles, Cn), dtype=torch.float64, device="cuda")
    w = rng.uniform(0.5, 2, R)
    fwd = dict(points=pts, indptr=indptr, indices=indices, data=data, transform="reciprocal")
    for wt in (None, w):
        chk(lib.bbb_ray_forward(ptr(idx), ptr(dev(v)), ptr(dev(k)), K, Cn, G, R, ptr(dev(indptr, np.int32)),
                                ptr(dev(indices, np.int32)), ptr(dev(data)), 1, ptr(dev(ing["d_obs"])),
                                None if wt is None else ptr(dev(wt)), ptr(dpred), ptr(phi), ptr(scratch), None))
        d = dpred.cpu().numpy()
        np.testing.assert_allclose(d[:, 0], kg["fw_dpred"], rtol=RTOL)
        for c in range(Cn):
            ref, ridx = OK.tomography_forward(fwd, sites_list[c], vals_list[c])
            np.testing.assert_allclose(d[:, c], ref, rtol=RTOL)
            res = ref - ing["d_obs"]
            np.testing.assert_allclose(phi[c].item(), np.sum((1.0 if wt is None else wt) * res * res), rtol=1e-12)
    # identity transform
    chk(lib.bbb_ray_forward(ptr(idx), ptr(dev(v)), ptr(dev(k)), K, Cn, G, R, ptr(dev(indptr, np.int32)),
                            ptr(dev(indices, np.int32)), ptr(dev(data)), 0, ptr(dev(ing["d_obs"])), None, ptr(dpred),
                            None, None, None))
    fwd["transform"] = "identity"
    np.testing.assert_allclose(dpred.cpu().numpy()[:, 3], OK.tomography_forward(fwd, sites_list[3], vals_list[3])[0], rtol=RTOL)


def test_dense_forward_loglike(lib, kg):
    """F3 + L1/L2 against LogLikelihood._get_misfit_and_det of the reference."""
    from bayesbay_b200 import synthetic

    ing = synthetic.polyfit()
    rng = np.random.default_rng(1)
    Cn = 100
    m = rng.normal(0, 5, (4, Cn))
    out = torch.zeros((15, Cn), dtype=torch.float64, device="cuda")
    chk(lib.bbb_dense_forward(ptr(dev(ing["G"])), ptr(dev(m)), 15, 4, Cn, ptr(out), None))
    np.testing.assert_allclose(out.cpu().numpy(), ing["G"] @ m, rtol=RTOL, atol=1e-12)
    dobs = kg["ll_dobs"]
    dp = np.stack((kg["ll_dp_old"], kg["ll_dp_new"]), axis=1)
    sg = np.array([0.21, 0.19])
    mis = torch.zeros(2, dtype=torch.float64, device="cuda")
    ld = torch.zeros(2, dtype=torch.float64, device="cuda")
    chk(lib.bbb_loglike(ptr(dev(dp)), ptr(dev(dobs)), ptr(dev(sg)), dobs.size, 2, ptr(mis), ptr(ld), None))
    mis, ld = mis.cpu().numpy(), ld.cpu().numpy()
    np.testing.assert_allclose([mis[0], ld[0]], kg["ll_old"], rtol=RTOL)
    np.testing.assert_allclose([mis[1], ld[1]], kg["ll_new"], rtol=RTOL)
    ratio = (ld[0] - ld[1] + mis[0] - mis[1]) / 2
    np.testing.assert_allclose(ratio, float(kg["ll_ratio_T1"]), rtol=1e-11)


def test_logprior(lib, kg):
    """L3: ParameterSpace.log_prior, incl. the exact known answer of tests/20_test_param_space_log_prior.py."""
    from bayesbay_b200 import _lib as L

    sp = L.SpaceSpec()
    sp.n_params, sp.kmax = 2, 1
    sp.prior_kind[0], sp.prior_a[0], sp.prior_b[0] = L.PRIOR_UNIFORM, -1.0, 1.0
    sp.prior_kind[1], sp.prior_a[1], sp.prior_b[1] = L.PRIOR_GAUSSIAN, 0.0, 1.0
    vals = np.zeros((2, 1, 3))
    vals[0, 0, 2] = 1.5  # out of the uniform support -> -inf
    out = torch.zeros(3, dtype=torch.float64, device="cuda")
    chk(lib.bbb_logprior(C.byref(sp), ptr(dev(vals)), ptr(dev(np.ones(3, np.int32))), 3, ptr(out), None))
    o = out.cpu().numpy()
    expected = math.log(1 / 2) + math.log(1 / math.sqrt(2 * math.pi))
    np.testing.assert_allclose(o[0], expected, rtol=1e-15)
    np.testing.assert_allclose(o[0], float(kg["lp_test20"]), rtol=1e-15)
    assert o[2] == -np.inf
    sp2 = L.SpaceSpec()
    sp2.n_params, sp2.kmax = 3, 7
    for j, (kind, a, b) in enumerate(((L.PRIOR_UNIFORM, 2.0, 4.0), (L.PRIOR_GAUSSIAN, -3.0, 2.0), (L.PRIOR_LAPLACE, 0.5, 2.0))):
        sp2.prior_kind[j], sp2.prior_a[j], sp2.prior_b[j] = kind, a, b
    v = kg["lp_vals"][:, :, None].repeat(2, axis=2)
    out = torch.zeros(2, dtype=torch.float64, device="cuda")
    chk(lib.bbb_logprior(C.byref(sp2), ptr(dev(v)), ptr(dev(np.array([7, 3], np.int32))), 2, ptr(out), None))
    np.testing.assert_allclose(out[0].item(), float(kg["lp_joint"]), rtol=1e-13)
    np.testing.assert_allclose(out[1].item(), kg["lp_each"][:, :3].sum(), rtol=1e-13)


def test_error_reporting(lib):
    from bayesbay_b200 import _lib as L

    rc = lib.bbb_raster2d(None, None, 8, 1, None, None, 4, None, None)
    assert rc == -1 and b"NULL" in lib.bbb_last_error()
    with pytest.raises(ValueError):
        L.check(rc)
