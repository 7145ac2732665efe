"""Mints tests/golden/taf_bra.npz: outputs of the reference's OWN TAF class (metabo/policies/policies.py:310-468, imported
from /root/reference through oracle/ref_loader.py; its GPy models are the oracle's restated-GPy shim) in both modes on a
state of the golden Branin episode.  Test infrastructure; authoring container only:  python -m oracle.make_golden_taf

The source-task data follow policies/taf/generate_taf_data_branin.py (M = 50 tasks x N = 100 Sobol points, fixed GP
hyper-parameters); they are regenerated here with the reference's own bra_var and stored in the fixture, so the GPU test
also pins metabo_b200.policies.taf.generate_taf_data_branin against them."""
import os
import pickle as pkl
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402


def main():
    load_reference()
    import sobol_seq
    from metabo.environment.objectives import bra_var
    from metabo.environment.util import scale_from_unit_square_to_domain, create_uniform_grid
    from metabo.policies.policies import TAF
    M, N, D = 50, 100, 2
    dom = np.array([[-0.1, 0.1], [-0.1, 0.1], [0.9, 1.1]])
    grid = scale_from_unit_square_to_domain(X=sobol_seq.i4_sobol_generate(dim_num=3, n=M), domain=dom)
    X = sobol_seq.i4_sobol_generate(dim_num=D, n=N)
    data = {"D": D, "M": M, "X": M * [X], "Y": M * [None], "kernel_lengthscale": M * [[0.235, 0.578]],
            "kernel_variance": M * [2.0], "noise_variance": M * [8.9e-16], "use_prior_mean_function": M * [False]}
    for i, p in enumerate(grid):
        data["Y"][i] = bra_var(x=X, t=np.array([p[0], p[1]]), s=p[2])
    tmp = os.path.join(tempfile.mkdtemp(), "taf_branin_M_50_N_100.pkl")
    with open(tmp, "wb") as f:
        pkl.dump(data, f)
    ep = np.load(os.path.join(ROOT, "tests", "golden", "episode_bra.npz"))
    rec = {"source_Y": np.stack([y[:, 0] for y in data["Y"]]), "source_X": X}
    for t in (0, 2, 7):
        p = "t%02d_" % t
        Xt, Yt = ep[p + "X"], ep[p + "Y"].reshape(-1, 1)
        taf_me = TAF(datafile=tmp, mode="me")
        cand = create_uniform_grid(np.stack([np.zeros(D), np.ones(D)], axis=1), 17)
        cand = cand[0] if isinstance(cand, tuple) else cand
        state = np.zeros((cand.shape[0], 2 + D + 2), dtype=np.float32)
        if t == 0:
            mean, var = np.zeros(cand.shape[0]), 2.0 * np.ones(cand.shape[0])
            model_target, X_target, inc = None, None, float(ep["y_min"])
        else:
            model_target = taf_me.train_gp(X=Xt, Y=Yt, kernel_lengthscale=[0.235, 0.578], kernel_variance=2.0,
                                           noise_variance=8.9e-16, use_prior_mean_function=False)
            m, v = model_target.predict_noiseless(cand)
            mean, var, X_target, inc = m[:, 0], v[:, 0], Xt, float(Yt.max())
        state[:, 0], state[:, 1], state[:, 2:2 + D], state[:, 2 + D], state[:, 3 + D] = mean, np.sqrt(var), cand, inc, t
        rec[p + "state"] = state
        rec[p + "X_target"] = np.zeros((0, D)) if X_target is None else X_target
        rec[p + "Y_target"] = np.zeros((0,)) if X_target is None else Yt[:, 0]
        rec[p + "taf_me"] = taf_me.af(state, X_target, model_target)
        rec[p + "taf_ranking"] = TAF(datafile=tmp, mode="ranking", rho=1.0).af(state, X_target, model_target)
    out = os.path.join(ROOT, "tests", "golden", "taf_bra.npz")
    np.savez_compressed(out, **rec)
    print(out, "%.1f KB" % (os.path.getsize(out) / 1024), {k: v.shape for k, v in rec.items()})


if __name__ == "__main__":
    main()
