"""Golden vectors of the full-matrix inverse covariance misfit, recorded from the UNMODIFIED reference
(``LogLikelihood._get_misfit_and_det`` with ``Target(covariance_mat_inv=<2-D array>)``,
src/bayesbay/likelihood/_log_likelihood.py:308-332, likelihood/_target.py:136-152).

    python tests/golden/make_cov_matrix.py      # needs baseline/_ref (``__graft_entry__.build()``)

Writes tests/golden/cov_matrix.npz: per case the data, the predicted data, the matrix and the reference's misfit,
log-determinant and log-likelihood ratio between two predicted-data vectors."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import reference_runner as RR  # noqa: E402


def main():
    bb = RR.import_reference()
    rng = np.random.default_rng(20261020)
    out = {}
    for case, n in enumerate((1, 2, 7, 25, 120)):
        a = rng.normal(size=(n, n))
        cov_inv = a @ a.T / n + np.diag(rng.uniform(0.5, 2.0, n))  # symmetric positive definite, dense
        dobs = rng.normal(size=n) * 3.0
        dpred = [dobs + rng.normal(size=n), dobs + rng.normal(size=n)]
        target = bb.likelihood.Target("t", dobs, covariance_mat_inv=cov_inv)
        holder = {}
        ll = bb.likelihood.LogLikelihood(targets=target, fwd_functions=lambda state: holder["d"])
        vals = []
        for d in dpred:
            holder["d"] = d
            state = bb.State({"dummy": bb.ParameterSpaceState(1, {"x": np.zeros(1)})})
            vals.append(ll._get_misfit_and_det(state))
        s_old = bb.State({"dummy": bb.ParameterSpaceState(1, {"x": np.zeros(1)})})
        s_new = bb.State({"dummy": bb.ParameterSpaceState(1, {"x": np.zeros(1)})})
        s_old.save_to_cache("t.dpred", dpred[0])
        s_new.save_to_cache("t.dpred", dpred[1])
        ratio = ll.log_likelihood_ratio(s_old, s_new, 1.7)
        out[f"c{case}_cov_inv"], out[f"c{case}_dobs"] = cov_inv, dobs
        out[f"c{case}_dpred"] = np.stack(dpred)
        out[f"c{case}_misfit"] = np.array([float(v[0]) for v in vals])
        out[f"c{case}_logdet"] = np.array([float(v[1]) for v in vals])
        out[f"c{case}_ratio_T1p7"] = np.array(float(ratio))
    out["n_cases"] = np.array(5)
    np.savez(os.path.join(HERE, "cov_matrix.npz"), **out)
    print("wrote cov_matrix.npz")


if __name__ == "__main__":
    main()
