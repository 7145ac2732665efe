"""GPU parity: NeuralAF evaluation (features -> MLP -> logits) and the per-episode reductions,
through the C ABI, against the oracle / the vectors minted from the reference.
Reference: metabo_gym.py:624-677 (get_state), policies.py:104-161, mlp.py:52-61.
Tolerance (north_star): logits within 1e-3 relative (relative to max|logit| of the candidate set,
logits cross zero mid-episode); the CUDA-core fp32 mode is held to 1e-5."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import metabo_oracle as O
from oracle import sampling_oracle as S

pytestmark = pytest.mark.gpu

CASES = [("bra", "weights_bra_1635.npz"), ("hm3", "weights_hm3_1988.npz"), ("gps", "weights_gps_1161.npz")]
MODES = [("fp32", 0, 1e-5), ("tf32", 1, 1e-3), ("bf16x3", 2, 2e-5), ("fp16", 3, 1e-3)]


def feat_codes(features, D, mask):
    from metabo_b200 import _lib as L
    codes = []
    if "posterior_mean" in features: codes.append(L.FEAT_MEAN)
    if "posterior_std" in features: codes.append(L.FEAT_STD)
    if "x" in features: codes += [L.FEAT_X0 + d for d in range(D)]
    if "incumbent" in features: codes.append(L.FEAT_INCUMBENT)
    if "timestep_perc" in features: codes.append(L.FEAT_TPERC)
    if "timestep" in features: codes.append(L.FEAT_TIMESTEP)
    if "budget" in features: codes.append(L.FEAT_BUDGET)
    return [c for c, keep in zip(codes, mask) if keep]


def make_policy(pw, net="policy_net"):
    from metabo_b200 import ops
    h = ops.PolicyHandle()
    layers = pw.layers(net)
    h.set_weights([w.cuda() for w, _ in layers], [b.cuda() for _, b in layers], pw.act)
    return h


def tc_supported(pw):
    l = pw.layers("policy_net")
    return len(l) == 5 and pw.act == "relu" and all(w.shape[0] == l[0][0].shape[0] for w, _ in l[:-1])


@pytest.mark.parametrize("name,wfile", CASES)
@pytest.mark.parametrize("mode_name,mode,tol", MODES)
def test_logits_on_reference_states(golden_dir, name, wfile, mode_name, mode, tol):
    """teacher-forced: the reference's own (mean, std, x, t, T) -> logits."""
    from metabo_b200 import ops
    dev = torch.device("cuda:0")
    ep = np.load(os.path.join(golden_dir, "episode_%s.npz" % name))
    spec = json.loads(str(ep["__env_spec__"]))
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, wfile))
    if mode != 0 and not tc_supported(pw):
        pytest.skip("tensor-core path needs the 4x200 ReLU architecture")
    pol = make_policy(pw)
    D, T = spec["D"], spec["T"]
    codes = feat_codes(spec["features"], D, pw.mask)
    for t in ep["keep"]:
        p = "t%02d_" % int(t)
        state, xi = ep[p + "state"], ep[p + "xi_t"]
        cs = ops.CandidateSet(D, tail=torch.as_tensor(xi, dtype=torch.float64, device=dev).contiguous())
        mean = torch.as_tensor(state[None, :, 0].copy(), device=dev)
        std = torch.as_tensor(state[None, :, 1].copy(), device=dev)
        epi = torch.tensor([[float(ep[p + "incumbent"]), int(t) / T, float(int(t)), float(T)]],
                           dtype=torch.float32, device=dev)
        lg = ops.af_logits(pol, codes, cs, mean, std, epi, 1, mode=mode)
        torch.cuda.synchronize()
        pol.check()
        got, ref = lg[0].cpu().numpy(), ep[p + "logits"]
        scale = np.max(np.abs(ref))
        assert np.max(np.abs(got - ref)) <= tol * scale + 1e-6, (name, int(t), mode_name)
        # deterministic-eval action (north_star): identical to the reference's unless the reference's own top two
        # logits are within the mode's tolerance of each other -- asserted for EVERY arithmetic mode
        a_ref = int(ep[p + "action"])
        a, _ = ops.logits_argmax(lg)
        a = int(a[0])
        assert a == a_ref or abs(ref[a] - ref[a_ref]) <= 2 * tol * scale, (name, int(t), mode_name, a, a_ref)


def test_value_net_matches_reference(golden_dir):
    from metabo_b200 import ops
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, "weights_bra_1635.npz"))
    vh = make_policy(pw, "value_net")
    tT = torch.tensor([[0., 30.], [7., 30.], [29., 30.]], device="cuda:0")
    v = ops.value_net(vh, tT).cpu().numpy()
    states = np.zeros((3, 4, 6), dtype=np.float32)
    states[:, :, 4] = tT[:, 0].cpu().numpy()[:, None]
    states[:, :, 5] = 30.0
    _, vo = O.policy_forward(pw, states)
    np.testing.assert_allclose(v, vo.numpy(), rtol=1e-5, atol=1e-5)


def test_end_to_end_gp_features_logits_vs_oracle(golden_dir):
    """GP fit -> posterior (fp32 cast) -> features -> logits for B ragged episodes against the
    oracle's get_state + NeuralAF.forward on well-conditioned synthetic Hartmann-3 states."""
    from metabo_b200 import ops
    dev = torch.device("cuda:0")
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, "weights_hm3_1988.npz"))
    pol = make_policy(pw)
    rng = np.random.RandomState(11)
    D, T, B = 3, 30, 6
    feats = ["posterior_mean", "posterior_std", "timestep", "budget", "x"]
    var, noise, ls = 0.83, 1e-5, np.array([0.716, 0.298, 0.186])
    ts = [0, 1, 5, 12, 20, 30]
    grids = O.Grids(D, N_MS=2000, N_LS=2000, k=5)
    X = np.zeros((B, T, D)); Y = np.zeros((B, T))
    for b, t in enumerate(ts):
        if t:
            X[b, :t] = rng.rand(t, D)
            Y[b, :t] = O.hm3_var(X[b, :t], np.zeros(3), 1.0)[:, 0]
    tt = lambda a, dt=torch.float64: torch.as_tensor(a, dtype=dt, device=dev).contiguous()
    state, status = ops.gp_fit(tt(np.array(ts), torch.int32), tt(X), tt(Y), tt(np.tile(ls, (B, 1))),
                               tt(np.full(B, var)), tt(np.full(B, noise)))
    cs = ops.CandidateSet(D, tail=tt(grids.multistart_grid))
    mean, std = ops.gp_posterior(state, T, cs, B, out_dtype=torch.float32)
    epi = tt(np.array([[0.0, t / T, t, T] for t in ts]), torch.float32)
    codes = feat_codes(feats, D, pw.mask)
    lg = ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=0).cpu().numpy()
    for b, t in enumerate(ts):
        gp = O.EpisodeGP(D, "RBF", var, ls, noise)
        gp.set_data(X[b, :t], Y[b, :t])
        s = O.get_state(gp, grids.multistart_grid, feats, t, T)
        ref = O.policy_forward(pw, s[None])[0][0].numpy()
        assert np.max(np.abs(lg[b] - ref)) <= 1e-3 * np.max(np.abs(ref)) + 1e-5, (b, t)


def test_reductions_vs_numpy_and_torch():
    from metabo_b200 import ops
    rng = np.random.RandomState(2)
    B, M = 9, 4999
    lg = (rng.randn(B, M) * 3).astype(np.float32)
    lg[0, 100] = lg[0, 4000] = lg[0].max() + 1.0        # exact tie -> first index
    lg[1, :] = -7.5                                     # all equal
    lgt = torch.as_tensor(lg, device="cuda:0")
    a, v = ops.logits_argmax(lgt)
    assert np.array_equal(a.cpu().numpy(), np.argmax(lg, axis=1))
    assert np.array_equal(v.cpu().numpy(), lg.max(axis=1))
    idx, val = ops.logits_topk(lgt, 5)
    ref_idx = np.argsort(-lg, axis=1, kind="stable")[:, :5]
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    acts = torch.as_tensor(rng.randint(0, M, size=B).astype(np.int32), device="cuda:0")
    lp, en = ops.logits_logp_entropy(lgt, acts)
    lpo, eno = O.categorical_logp_entropy(torch.as_tensor(lg), acts.cpu().numpy())
    np.testing.assert_allclose(lp.cpu().numpy(), lpo.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(en.cpu().numpy(), eno.numpy(), rtol=1e-5, atol=1e-5)
    # large-magnitude logits as the shipped Branin net produces (|logit| ~ 1500)
    big = (1349.0 + rng.rand(2, 966)).astype(np.float32)
    lp2, en2 = ops.logits_logp_entropy(torch.as_tensor(big, device="cuda:0"),
                                       torch.zeros(2, dtype=torch.int32, device="cuda:0"))
    # torch's own fp32 Categorical quantises (logit - lse) to the fp32 ulp of 1349 (1.2e-4): compare with the
    # fp64 evaluation tightly and with the fp32 one at its own noise level
    lpo2, eno2 = O.categorical_logp_entropy(torch.as_tensor(big, dtype=torch.float64), np.zeros(2, dtype=np.int64))
    np.testing.assert_allclose(lp2.cpu().numpy(), lpo2.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(en2.cpu().numpy(), eno2.numpy(), rtol=1e-5, atol=1e-5)
    lpo3, eno3 = O.categorical_logp_entropy(torch.as_tensor(big), np.zeros(2, dtype=np.int64))
    np.testing.assert_allclose(en2.cpu().numpy(), eno3.numpy(), rtol=0, atol=1e-3)


def test_sampler_matches_oracle_noise_and_distribution():
    from metabo_b200 import ops
    rng = np.random.RandomState(5)
    B, M = 64, 50
    lg = (rng.randn(B, M) * 2).astype(np.float32)
    lgt = torch.as_tensor(lg, device="cuda:0")
    a = ops.logits_sample(lgt, seed=1234, step=7).cpu().numpy()
    for b in range(B):
        noisy = lg[b] + S.gumbel_noise(1234, b, 7, M)
        assert noisy[a[b]] >= noisy.max() - 1e-4            # same noise, device logf within an ulp
    a2 = ops.logits_sample(lgt, seed=1234, step=7).cpu().numpy()
    assert np.array_equal(a, a2)                             # counter-based: reproducible
    # chi-square of the empirical distribution of one row over 20000 steps (rows = steps trick)
    n = 20000
    row = np.tile(lg[0, :8][None], (n, 1)).astype(np.float32)
    acts = ops.logits_sample(torch.as_tensor(row, device="cuda:0"), seed=99, step=0).cpu().numpy()
    p = np.exp(row[0] - row[0].max()); p /= p.sum()
    cnt = np.bincount(acts, minlength=8)
    chi2 = np.sum((cnt - n * p) ** 2 / (n * p))
    assert chi2 < 30.0                                       # 7 dof, p ~ 1e-4


@pytest.mark.parametrize("mode_name,mode,tol", [("tf32", 1, 1e-3), ("bf16x3", 2, 2e-5), ("fp16", 3, 1e-3)])
def test_fused_gp_mlp_kernel_matches_two_kernel_path(golden_dir, mode_name, mode, tol):
    """mbo_af_eval_fused (fp64 GP warps inside the tensor-core kernel) == mbo_gp_posterior (fp64) +
    mbo_af_logits on the same inputs: identical mean/std (bitwise, same arithmetic) and logits within
    the mode's tolerance of the fp32 CUDA-core evaluation; ragged n incl. empty GPs; local-grid and
    head+tail candidate descriptors."""
    from metabo_b200 import ops
    dev = torch.device("cuda:0")
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, "weights_hm3_1988.npz"))
    pol = make_policy(pw)
    rng = np.random.RandomState(21)
    D, T, B = 3, 30, 7
    ns = [0, 1, 7, 8, 19, 30, 30]
    X = np.zeros((B, T, D)); Y = np.zeros((B, T))
    for b, n in enumerate(ns):
        if n:
            X[b, :n] = rng.rand(n, D)
            Y[b, :n] = O.hm3_var(X[b, :n], np.zeros(3), 1.0)[:, 0]
    tt = lambda a, dt=torch.float64: torch.as_tensor(a, dtype=dt, device=dev).contiguous()
    state, status = ops.gp_fit(tt(np.array(ns), torch.int32), tt(X), tt(Y), tt(np.tile([0.716, 0.298, 0.186], (B, 1))),
                               tt(np.full(B, 0.83)), tt(np.full(B, 1.688e-11)))
    epi = tt(np.array([[0.0, n / T, n, T] for n in ns]), torch.float32)
    feats = ["posterior_mean", "posterior_std", "timestep", "budget", "x"]
    codes = feat_codes(feats, D, pw.mask)
    grids = O.Grids(D, N_MS=2000, N_LS=300, k=5)
    starts = np.stack([grids.multistart_grid[rng.choice(grids.N_MS, 5, replace=False)] for _ in range(B)])
    cubes = ops.cubes_around(tt(starts), grids.af_max_search_diam)
    sets = [ops.CandidateSet(D, tail=tt(grids.multistart_grid)),
            ops.CandidateSet(D, tail=tt(grids.local_search_grid), cubes=cubes),
            ops.CandidateSet(D, head=tt(rng.rand(B, 5, D)), tail=tt(grids.multistart_grid))]
    for cs in sets:
        mean, std = ops.gp_posterior(state, T, cs, B, out_dtype=torch.float32)
        ref = ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=0)
        m2 = torch.empty_like(mean); s2 = torch.empty_like(std)
        got = ops.af_eval_fused(pol, codes, cs, state, T, T, epi, B, mode=mode, mean=m2, std=s2)
        torch.cuda.synchronize()
        pol.check()
        assert torch.equal(m2, mean) and torch.equal(s2, std)
        scale = float(ref.abs().max())
        assert float((got - ref).abs().max()) <= tol * scale + 1e-6, (mode_name, cs.M)
        got2 = ops.af_eval_fused(pol, codes, cs, state, T, T, epi, B, mode=mode)      # without the posterior copies
        torch.cuda.synchronize()
        assert torch.equal(got, got2)


def test_engine_round_reuse_of_multistart_results_is_bitwise(golden_dir):
    """AFEngine.reuse_multistart: phase 3 evaluates only the k new maxima and reuses the phase-1 results of the
    multistart grid -> logits, action and posterior identical to the full re-evaluation the reference does
    (metabo_gym.py:214 get_state(xi_t))."""
    from metabo_b200 import _lib as L
    from metabo_b200.engine import AFEngine
    from metabo_b200.sobol import i4_sobol_generate
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, "weights_bra_1635.npz"))
    rng = np.random.RandomState(3)
    B, D, T = 5, 2, 30
    ns = [0, 2, 9, 17, 30]
    X = np.zeros((B, T, D)); Y = np.zeros((B, T))
    for b, n in enumerate(ns):
        if n:
            X[b, :n] = rng.rand(n, D)
            Y[b, :n] = O.bra_var(X[b, :n], np.zeros(2), 1.0)[:, 0]
    outs = []
    for reuse in (False, True):
        eng = AFEngine(B, D, T, ["posterior_mean", "posterior_std", "timestep", "budget", "x"], N_MS=1000, N_LS=1000, k=5,
                       local_search_grid=i4_sobol_generate(D, 1000), mlp_mode=L.MLP_BF16X3)
        layers = pw.layers("policy_net")
        eng.set_policy([w for w, _ in layers], [b for _, b in layers], pw.act)
        eng.set_hyperparameters([0.235, 0.578], 2.0, 1e-6)
        eng.set_data(torch.as_tensor(X, device=eng.dev), torch.as_tensor(Y, device=eng.dev),
                     torch.as_tensor(np.array(ns, dtype=np.int32), device=eng.dev))
        eng.set_episode_scalars(np.zeros(B), np.array(ns, dtype=np.float32), np.full(B, T, dtype=np.float32))
        eng.reuse_multistart = reuse
        outs.append(eng.af_round(deterministic=True, want_posterior=True))
        torch.cuda.synchronize()
    a, b = outs
    for k in ("logits", "mean", "std", "action", "x_action", "af_maxima"):
        assert torch.equal(a[k], b[k]), k


def test_sampler_is_invariant_under_sharding():
    """Gumbel noise is keyed by the GLOBAL episode index: sampling a slice with episode_offset equals the
    corresponding rows of the full batch (what makes rollouts independent of the number of GPUs)."""
    from metabo_b200 import ops
    rng = np.random.RandomState(8)
    lg = torch.as_tensor((rng.randn(12, 300) * 2).astype(np.float32), device="cuda:0")
    full = ops.logits_sample(lg, seed=5, step=3).cpu().numpy()
    part = ops.logits_sample(lg[7:].contiguous(), seed=5, step=3, episode_offset=7).cpu().numpy()
    assert np.array_equal(full[7:], part)
    # the oracle restates the noise bit for bit up to libm-vs-CUDA logf (<= 2 ulp): the sampled action may only differ
    # where the top two perturbed logits are that close
    ref = S.sample(lg.cpu().numpy(), 5, 3)
    pert = lg.cpu().numpy() + np.stack([S.gumbel_noise(5, b, 3, lg.shape[1]) for b in range(lg.shape[0])])
    for b in range(lg.shape[0]):
        assert ref[b] == full[b] or abs(pert[b, ref[b]] - pert[b, full[b]]) <= 1e-5 * np.abs(pert[b]).max(), b
    assert (ref == full).mean() >= 0.9


@pytest.mark.parametrize("D", [4, 6])
def test_tensor_core_kernel_with_wide_feature_vectors(D):
    """F = D + 4 features (8 and 10 > 7) take the two-slice layer-0 path of the tcgen05 kernel."""
    from metabo_b200 import ops, _lib as L
    dev = torch.device("cuda:0")
    rng = np.random.RandomState(D)
    F = D + 4
    dims = [F, 200, 200, 200, 200, 1]
    W = [torch.as_tensor(rng.normal(0, 0.08, (dims[i + 1], dims[i])).astype(np.float32), device=dev) for i in range(5)]
    b = [torch.as_tensor(rng.normal(0, 0.05, dims[i + 1]).astype(np.float32), device=dev) for i in range(5)]
    pol = ops.PolicyHandle(); pol.set_weights(W, b, "relu")
    B, M = 3, 777
    cs = ops.CandidateSet(D, tail=torch.as_tensor(rng.rand(M, D), device=dev))
    mean = torch.as_tensor(rng.randn(B, M).astype(np.float32), device=dev)
    std = torch.as_tensor(rng.rand(B, M).astype(np.float32), device=dev)
    epi = torch.as_tensor(np.array([[0.1, 0.2, 6.0, 30.0]] * B, dtype=np.float32), device=dev)
    codes = [L.FEAT_MEAN, L.FEAT_STD] + [L.FEAT_X0 + d for d in range(D)] + [L.FEAT_TIMESTEP, L.FEAT_BUDGET]
    ref = ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=0)
    # torch fp64 ground truth of the same network
    x = cs.tail.float()[None].expand(B, M, D)
    feats = torch.cat([mean[..., None], std[..., None], x, epi[:, None, 2:3].expand(B, M, 1), epi[:, None, 3:4].expand(B, M, 1)], 2).double()
    h = feats
    for i in range(4):
        h = torch.relu(h @ W[i].double().T + b[i].double())
    truth = (h @ W[4].double().T + b[4].double())[..., 0]
    scale = float(truth.abs().max())
    assert float((ref.double() - truth).abs().max()) <= 1e-5 * scale
    # random N(0, 0.08) nets cancel heavily (|logit| ~ 0.3 from O(1) hidden activations): TF32's operand rounding is
    # then ~4e-3 of max|logit| (bf16x3: ~1e-4); the 1e-3 bar is stated (and tested above) on the shipped policies
    for mode, tol in ((1, 1e-2), (2, 3e-4), (3, 1e-2)):
        got = ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=mode)
        torch.cuda.synchronize(); pol.check()
        assert float((got.double() - truth).abs().max()) <= tol * scale, (D, mode)


def test_full_size_properties_hartmann3_batch(golden_dir):
    """BASELINE configs[1] at full size (B=1024 episodes, 10 000 local-grid candidates each, n=30 observations):
    size-independent properties instead of an oracle run -- (a) every episode of the batch is bitwise what it is
    when evaluated alone, (b) the arithmetic modes agree within their tolerances, (c) top-k is the sorted true top-k,
    (d) the posterior std is invariant to Y and the mean linear in Y."""
    from metabo_b200 import ops
    dev = torch.device("cuda:0")
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, "weights_hm3_1988.npz"))
    pol = make_policy(pw)
    B, D, T, n, k, NLS = 1024, 3, 30, 30, 5, 2000
    g = torch.Generator(device=dev); g.manual_seed(7)
    X = torch.rand((B, T, D), generator=g, dtype=torch.float64, device=dev)
    Y1 = torch.randn((B, T), generator=g, dtype=torch.float64, device=dev)
    Y2 = torch.randn((B, T), generator=g, dtype=torch.float64, device=dev)
    nn = torch.full((B,), n, dtype=torch.int32, device=dev)
    ls = torch.tensor([[0.716, 0.298, 0.186]] * B, dtype=torch.float64, device=dev)
    var = torch.full((B,), 0.83, dtype=torch.float64, device=dev)
    noise = torch.full((B,), 1e-6, dtype=torch.float64, device=dev)
    from metabo_b200.sobol import i4_sobol_generate
    table = torch.as_tensor(i4_sobol_generate(D, NLS), device=dev)
    starts = torch.rand((B, k, D), generator=g, dtype=torch.float64, device=dev)
    cubes = ops.cubes_around(starts, 2.0 / 12)
    cs = ops.CandidateSet(D, tail=table, cubes=cubes)
    assert cs.M == 10000
    epi = torch.zeros(B, 4, device=dev); epi[:, 1] = 0.5; epi[:, 2] = 15; epi[:, 3] = 30
    codes = feat_codes(["posterior_mean", "posterior_std", "timestep", "budget", "x"], D, pw.mask)
    st1, s1 = ops.gp_fit(nn, X, Y1, ls, var, noise)
    st2, _ = ops.gp_fit(nn, X, Y2, ls, var, noise)
    st3, _ = ops.gp_fit(nn, X, 2.0 * Y1 - 0.5 * Y2, ls, var, noise)
    assert int(s1.abs().max()) == 0
    m1, sd1 = ops.gp_posterior(st1, T, cs, B)
    m2, sd2 = ops.gp_posterior(st2, T, cs, B)
    m3, sd3 = ops.gp_posterior(st3, T, cs, B)
    assert torch.equal(sd1, sd2) and torch.equal(sd1, sd3)                                        # (d)
    assert float((m3 - (2.0 * m1 - 0.5 * m2)).abs().max()) < 1e-8 * max(1.0, float(m1.abs().max()))
    mean, std = m1.float(), sd1.float()
    lg = {m: ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=m) for m in (0, 1, 2, 3)}
    torch.cuda.synchronize(); pol.check()
    scale = float(lg[0].abs().max())
    assert float((lg[1] - lg[0]).abs().max()) <= 1e-3 * scale                                     # (b)
    assert float((lg[2] - lg[0]).abs().max()) <= 2e-5 * scale
    assert float((lg[3] - lg[0]).abs().max()) <= 1e-3 * scale          # fp16 = the mode bench.py times
    # deterministic argmax per episode: every mode picks the fp32 kernel's action or one within its tolerance of it
    a0, v0 = ops.logits_argmax(lg[0])
    for m, tol in ((1, 1e-3), (2, 2e-5), (3, 1e-3)):
        am, _ = ops.logits_argmax(lg[m])
        picked = lg[0].gather(1, am.long()[:, None])[:, 0]
        assert bool(((am == a0) | ((v0 - picked) <= 2 * tol * scale)).all()), m
    for b in (0, 517, 1023):                                                                      # (a)
        csb = ops.CandidateSet(D, tail=table, cubes=cubes[b:b + 1].contiguous())
        mb, sb = ops.gp_posterior(st1[b:b + 1].contiguous(), T, csb, 1)
        assert torch.equal(mb[0], m1[b]) and torch.equal(sb[0], sd1[b])
        for m in (1, 2, 3):
            one = ops.af_logits(pol, codes, csb, mean[b:b + 1].contiguous(), std[b:b + 1].contiguous(),
                                epi[b:b + 1].contiguous(), 1, mode=m)
            assert torch.equal(one[0], lg[m][b])
    idx, val = ops.logits_topk(lg[2], k)                                                          # (c)
    tv, ti = torch.topk(lg[2], k, dim=1)
    assert torch.equal(val, tv)
    assert torch.equal(torch.gather(lg[2], 1, idx.long()), val)
    assert bool((val[:, :-1] >= val[:, 1:]).all())
    a, v = ops.logits_argmax(lg[2])
    assert torch.equal(v, lg[2].max(dim=1).values) and torch.equal(a.long(), idx[:, 0].long())


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_fused_topk_epilogue_equals_logits_then_topk(golden_dir, mode):
    """mbo_af_topk (selection inside the tensor-core kernel's epilogue + merge kernel, logits optional) returns
    exactly what mbo_af_logits followed by mbo_logits_topk returns; ragged M, exact ties, k = 1 and 5."""
    from metabo_b200 import ops
    dev = torch.device("cuda:0")
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, "weights_hm3_1988.npz"))
    pol = make_policy(pw)
    rng = np.random.RandomState(31)
    codes = [0, 1, 2, 3, 4, 12, 13]
    for B, M in ((1, 5), (3, 127), (2, 1733), (16, 10000)):
        x = rng.rand(M, 3)
        if M > 300:
            x[200] = x[7]; x[M - 1] = x[7]                        # duplicated candidates -> exact logit ties
        cs = ops.CandidateSet(3, tail=torch.as_tensor(x, device=dev))
        mean = torch.as_tensor(np.tile(rng.randn(1, M), (B, 1)).astype(np.float32), device=dev)
        std = torch.as_tensor(np.tile(rng.rand(1, M), (B, 1)).astype(np.float32), device=dev)
        if M > 300:
            mean[:, 200] = mean[:, 7]; mean[:, M - 1] = mean[:, 7]; std[:, 200] = std[:, 7]; std[:, M - 1] = std[:, 7]
        epi = torch.as_tensor(np.array([[0.0, 0.3, 9.0, 30.0]] * B, dtype=np.float32), device=dev)
        lg = ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=mode)
        for k in (1, min(5, M)):
            ri, rv = ops.logits_topk(lg, k)
            lg2 = torch.empty_like(lg)
            i1, v1 = ops.af_topk(pol, codes, cs, mean, std, epi, B, k, mode=mode, logits=lg2)
            i2, v2 = ops.af_topk(pol, codes, cs, mean, std, epi, B, k, mode=mode)          # logits never written
            torch.cuda.synchronize(); pol.check()
            assert torch.equal(lg2, lg)
            assert torch.equal(i1, ri) and torch.equal(v1, rv), (B, M, k)
            assert torch.equal(i2, ri) and torch.equal(v2, rv), (B, M, k)


def _random_relu_net(rng, F, H, w_std, b_std, dev):
    dims = [F, H, H, H, H, 1]
    W = [torch.as_tensor(rng.normal(0, w_std, (dims[i + 1], dims[i])).astype(np.float32), device=dev) for i in range(5)]
    b = [torch.as_tensor(rng.normal(0, b_std, dims[i + 1]).astype(np.float32), device=dev) for i in range(5)]
    return W, b


def _truth_logits(W, b, feats):
    h = feats.double()
    for i in range(4):
        h = torch.relu(h @ W[i].double().T + b[i].double())
    return (h @ W[4].double().T + b[4].double())[..., 0]


@pytest.mark.parametrize("w_std,b_std", [(1e-4, 1e-5), (0.01, 0.0), (0.07, 0.05), (30.0, 10.0)])
def test_fp16_mode_renormalises_layers_by_powers_of_two(w_std, b_std):
    """MBO_MLP_FP16 folds a power-of-two scale into layers whose weights are far from O(1) (exact, ReLU is
    positively homogeneous): nets whose activations would underflow (w ~ 1e-4: a4 ~ 1e-12) or overflow
    (w ~ 30: a4 ~ 1e9) in plain fp16 agree with the fp64 truth as closely as the tf32 mode does."""
    from metabo_b200 import ops, _lib as L
    dev = torch.device("cuda:0")
    rng = np.random.RandomState(5)
    D, B, M = 3, 2, 1000
    W, b = _random_relu_net(rng, 7, 200, w_std, b_std, dev)
    b[4].zero_()                      # keep the (tiny or huge) hidden signal visible in the fp32 logit
    pol = ops.PolicyHandle(); pol.set_weights(W, b, "relu")
    cs = ops.CandidateSet(D, tail=torch.as_tensor(rng.rand(M, D), device=dev))
    mean = torch.as_tensor(rng.randn(B, M).astype(np.float32), device=dev)
    std = torch.as_tensor(rng.rand(B, M).astype(np.float32), device=dev)
    epi = torch.as_tensor(np.array([[0.1, 0.2, 6.0, 30.0]] * B, dtype=np.float32), device=dev)
    codes = [0, 1, 2, 3, 4, 12, 13]
    x = cs.tail.float()[None].expand(B, M, D)
    feats = torch.cat([mean[..., None], std[..., None], x, epi[:, None, 2:3].expand(B, M, 1), epi[:, None, 3:4].expand(B, M, 1)], 2)
    truth = _truth_logits(W, b, feats)
    scale = float(truth.abs().max())
    e16 = float((ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=L.MLP_FP16).double() - truth).abs().max())
    e32 = float((ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=L.MLP_TF32).double() - truth).abs().max())
    torch.cuda.synchronize(); pol.check()
    assert e16 <= 4.0 * e32 + 1e-6 * scale, (w_std, e16, e32, scale)
    assert e16 <= 2e-2 * scale                                    # random nets cancel heavily, see the wide-feature test


def test_fp16_mode_flags_activations_beyond_its_range():
    """A layer-0 row norm inside the unscaled band but inputs of 1e6 push activations past 65504: the kernel clamps
    (satfinite) and raises bit 1 of the workspace flag, which PolicyHandle.check() turns into an error."""
    from metabo_b200 import ops, _lib as L
    dev = torch.device("cuda:0")
    rng = np.random.RandomState(6)
    W, b = _random_relu_net(rng, 7, 200, 0.07, 0.0, dev)
    W[0] = W[0] * 5.0
    pol = ops.PolicyHandle(); pol.set_weights(W, b, "relu")
    B, M = 1, 300
    cs = ops.CandidateSet(3, tail=torch.as_tensor(rng.rand(M, 3), device=dev))
    mean = torch.full((B, M), 1e6, dtype=torch.float32, device=dev)
    std = torch.ones((B, M), dtype=torch.float32, device=dev)
    epi = torch.as_tensor(np.array([[0.1, 0.2, 6.0, 30.0]], dtype=np.float32), device=dev)
    lg = ops.af_logits(pol, [0, 1, 2, 3, 4, 12, 13], cs, mean, std, epi, B, mode=L.MLP_FP16)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(lg).all())
    with pytest.raises(L.MboError, match="fp16 range"):
        pol.check()
    ops.af_logits(pol, [0, 1, 2, 3, 4, 12, 13], cs, mean, std, epi, B, mode=L.MLP_TF32)   # tf32 has fp32's range
    torch.cuda.synchronize(); pol.check()


def test_fp16_mode_needs_two_spare_columns():
    from metabo_b200 import ops, _lib as L
    dev = torch.device("cuda:0")
    rng = np.random.RandomState(8)
    for H, ok in ((206, True), (208, False)):
        W, b = _random_relu_net(rng, 7, H, 0.07, 0.05, dev)
        pol = ops.PolicyHandle(); pol.set_weights(W, b, "relu")
        M = 130
        cs = ops.CandidateSet(3, tail=torch.as_tensor(rng.rand(M, 3), device=dev))
        mean = torch.as_tensor(rng.randn(1, M).astype(np.float32), device=dev)
        std = torch.as_tensor(rng.rand(1, M).astype(np.float32), device=dev)
        epi = torch.as_tensor(np.array([[0.1, 0.2, 6.0, 30.0]], dtype=np.float32), device=dev)
        codes = [0, 1, 2, 3, 4, 12, 13]
        ref = ops.af_logits(pol, codes, cs, mean, std, epi, 1, mode=L.MLP_FP32)
        if ok:
            got = ops.af_logits(pol, codes, cs, mean, std, epi, 1, mode=L.MLP_FP16)
            torch.cuda.synchronize(); pol.check()
            assert float((got - ref).abs().max()) <= 1e-2 * float(ref.abs().max())
        else:
            with pytest.raises(L.MboError, match="H <= 206"):
                ops.af_logits(pol, codes, cs, mean, std, epi, 1, mode=L.MLP_FP16)
