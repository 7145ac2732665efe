"""GPU parity of the path bench.py times -- `AFEngine.af_round` (GP refit -> multistart phase -> local Sobol phase ->
xi_t -> greedy action, metabo_gym.py:195-217, 615-622, 706-721) -- against the per-step traces of the REFERENCE'S
OWN deterministic episodes (tests/golden/episode_{bra,hm3,gps}.npz, minted by oracle/make_golden.py), for every
arithmetic mode of the NeuralAF kernel, and 100-episode regret statistics against tests/golden/regret_*.npz
(plot_results.py:61-63: median / 30th / 70th percentile).

Tolerances.  Logits: `tol` of the mode (north_star: 1e-3 relative to max|logit|; fp32 kernel 1e-5, bf16x3 2e-5) PLUS
the reference GP route's own error on that state: the deterministic policy re-samples points, cond(Ky) reaches 1e9 and
GPy's explicit-inverse route is then only ~5e-4 accurate in sigma (DESIGN.md "GP precision"), while the CUDA kernel
follows the long-double truth.  The slack is measured, per state, as the logit difference the ORACLE produces between
its GPy-route and its truth-route posterior; it is ~0 for the early, well-conditioned steps, where the bare mode
tolerance applies.  Selections (top-k starts, top-k maxima, action): identical to the reference's, except where the
reference's own scores of the two alternatives are within twice that bound ("near-tie")."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import gp_oracle as G
from oracle import metabo_oracle as O

pytestmark = pytest.mark.gpu

MODES = [("fp32", 0, 1e-5), ("tf32", 1, 1e-3), ("bf16x3", 2, 2e-5), ("fp16", 3, 1e-3)]
# Phases 1-2 (multistart / local grids, 1 728-10 000 candidates that only feed the top-k START selection): the 11-bit
# modes are held to 1.5e-3 there.  Measured maximum over the 11 golden states x 2 phases: 1.2e-3 (hm3 t=12, local grid,
# both fp16 and tf32), every other state < 1e-3; the states the policy ACTS on (xi_t, phase 3) are held to 1e-3 here and
# in test_af_gpu.py::test_logits_on_reference_states.  bf16x3 / fp32 modes: no such allowance.
SEARCH_PHASE_TOL = {"tf32": 1.5e-3, "fp16": 1.5e-3}
WEIGHTS = {"bra": "weights_bra_1635.npz", "hm3": "weights_hm3_1988.npz", "gps": "weights_gps_1161.npz"}


def _engine_for(ep, spec, pw, mode, B):
    from metabo_b200.engine import AFEngine
    from metabo_b200.sobol import i4_sobol_generate
    D, T = spec["D"], spec["T"]
    kw = dict(kernel=spec.get("kernel", "RBF"), use_prior_mean=bool(spec.get("use_prior_mean_function", False)),
              mlp_mode=mode, policy_mask=pw.mask)
    if spec["local_af_opt"]:
        eng = AFEngine(B, D, T, spec["features"], N_MS=spec["N_MS"], N_LS=spec["N_LS"], k=spec["k"],
                       local_search_grid=i4_sobol_generate(D, spec["N_LS"]), **kw)
    else:
        eng = AFEngine(B, D, T, spec["features"], discrete_domain=i4_sobol_generate(D, spec["cardinality_domain"]), **kw)
    layers = pw.layers("policy_net")
    eng.set_policy([w.cuda() for w, _ in layers], [b.cuda() for _, b in layers], pw.act)
    eng.set_hyperparameters(ep["kernel_lengthscale"], float(ep["kernel_variance"]), float(ep["noise_variance"]))
    return eng


def _oracle_logits(pw, spec, ep, p, t, Xstar, route):
    gp = O.EpisodeGP(spec["D"], spec.get("kernel", "RBF"), float(ep["kernel_variance"]), ep["kernel_lengthscale"],
                     float(ep["noise_variance"]), use_prior_mean=bool(spec.get("use_prior_mean_function", False)),
                     y_min=float(ep["y_min"]), route=route)
    gp.set_data(ep[p + "X"], ep[p + "Y"])
    s = O.get_state(gp, Xstar, spec["features"], t, spec["T"])
    return O.policy_forward(pw, s[None])[0][0].numpy()


def _gp_slack(pw, spec, ep, p, t, Xstar):
    """max logit difference the reference-route GP error causes on this state (0 for the empty GP)"""
    if ep[p + "X"].shape[0] == 0:
        return 0.0
    a = _oracle_logits(pw, spec, ep, p, t, Xstar, "gpy")
    b = _oracle_logits(pw, spec, ep, p, t, Xstar, "truth")
    return float(np.max(np.abs(a - b)))


def _same_or_near_tie(ref_scores, got_idx, ref_idx, bound):
    """got_idx == ref_idx, or every position where they differ is a near-tie in the reference's own scores"""
    got_idx, ref_idx = np.asarray(got_idx), np.asarray(ref_idx)
    if np.array_equal(got_idx, ref_idx):
        return True, True
    ok = all(abs(ref_scores[g] - ref_scores[r]) <= 2 * bound for g, r in zip(got_idx, ref_idx) if g != r)
    return ok, False


@pytest.mark.parametrize("name", ["bra", "hm3"])
@pytest.mark.parametrize("mode_name,mode,tol", MODES)
def test_af_round_on_golden_episode_states(golden_dir, name, mode_name, mode, tol):
    from metabo_b200 import ops
    ep = np.load(os.path.join(golden_dir, "episode_%s.npz" % name))
    spec = json.loads(str(ep["__env_spec__"]))
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, WEIGHTS[name]))
    keep = [int(t) for t in ep["keep"]]
    B, D, T, k = len(keep), spec["D"], spec["T"], spec["k"]
    eng = _engine_for(ep, spec, pw, mode, B)
    dev = eng.dev
    X = np.zeros((B, T, D)); Y = np.zeros((B, T)); inc = np.zeros(B)
    for b, t in enumerate(keep):
        p = "t%02d_" % t
        X[b, :t], Y[b, :t] = ep[p + "X"], ep[p + "Y"]
        inc[b] = float(ep[p + "incumbent"])
    eng.set_data(torch.as_tensor(X, device=dev), torch.as_tensor(Y, device=dev),
                 torch.as_tensor(np.array(keep, dtype=np.int32), device=dev))
    eng.set_episode_scalars(inc, np.array(keep, dtype=np.float32), np.full(B, T, dtype=np.float32))
    eng.keep_phase_logits = True
    out = eng.af_round(deterministic=True, want_posterior=True)
    torch.cuda.synchronize()
    eng.policy.check()
    assert int(eng.status.abs().max()) == 0            # no jitter retries on the reference's own states
    grids = O.Grids(D, N_MS=spec["N_MS"], N_LS=spec["N_LS"], k=k)
    full_chain = 0
    for b, t in enumerate(keep):
        p = "t%02d_" % t
        # ---- phase 1: AF on the multistart grid, top-k starts (metabo_gym.py:707-709)
        ref1, got1 = ep[p + "af_ms"], out["af_ms"][b].cpu().numpy()
        scale1 = np.max(np.abs(ref1))
        tol12 = SEARCH_PHASE_TOL.get(mode_name, tol)
        bound1 = tol12 * scale1 + 1.5 * _gp_slack(pw, spec, ep, p, t, grids.multistart_grid) + 1e-6
        assert np.max(np.abs(got1 - ref1)) <= bound1, (name, t, mode_name, "af_ms")
        ref_top1 = np.argsort(-ref1, kind="stable")[:k]
        ok, same = _same_or_near_tie(ref1, out["top_ms"][b].cpu().numpy(), ref_top1, bound1)
        assert ok, (name, t, mode_name, "top_ms", out["top_ms"][b].tolist(), ref_top1.tolist())
        if not same:
            continue
        # ---- phase 2: local Sobol grids in clipped cubes, bit for bit (metabo_gym.py:711-716, util.py:45-51)
        cs_ls = ops.CandidateSet(D, tail=eng.local_table, cubes=out["cubes"][b:b + 1].contiguous())
        idx = torch.arange(cs_ls.M, dtype=torch.int32, device=dev)[None].contiguous()
        local = ops.candidates_gather(cs_ls, idx, 1)[0].cpu().numpy()
        assert np.array_equal(local, ep[p + "local"]), (name, t, "local grids")
        ref2, got2 = ep[p + "af_ls"], out["af_ls"][b].cpu().numpy()
        scale2 = np.max(np.abs(ref2))
        bound2 = tol12 * scale2 + 1.5 * _gp_slack(pw, spec, ep, p, t, ep[p + "local"]) + 1e-6
        assert np.max(np.abs(got2 - ref2)) <= bound2, (name, t, mode_name, "af_ls")
        ref_top2 = np.argsort(-ref2, kind="stable")[:k]
        ok, same = _same_or_near_tie(ref2, out["top_ls"][b].cpu().numpy(), ref_top2, bound2)
        assert ok, (name, t, mode_name, "top_ls")
        if not same:
            continue
        # ---- phase 3: xi_t = [maxima; multistart] (:619), logits, greedy action (policies.py:139-141)
        xi = np.concatenate([out["af_maxima"][b].cpu().numpy(), grids.multistart_grid])
        xi_ref = ep[p + "xi_t"]
        assert np.array_equal(xi[k:], xi_ref[k:]), (name, t, "xi_t multistart part")
        # the k maxima: the same points; EXACTLY tied scores may come in another order (the reference sorts with numpy's
        # default unstable argsort, :719) -> align ours to the reference's order
        perm = []
        for j in range(k):
            hit = [i for i in range(k) if np.array_equal(xi[i], xi_ref[j]) and i not in perm]
            assert hit, (name, t, "xi_t maxima", xi[:k], xi_ref[:k])
            perm.append(hit[0])
        assert perm == list(range(k)) or len(set(np.sort(ref2)[::-1][:k].tolist())) < k, (name, t, "xi_t order", perm)
        order = np.array(perm + list(range(k, xi.shape[0])))
        xi = xi[order]
        ref3, got3 = ep[p + "logits"], out["logits"][b].cpu().numpy()[order]
        scale3 = np.max(np.abs(ref3))
        bound3 = tol * scale3 + 1.5 * _gp_slack(pw, spec, ep, p, t, xi) + 1e-6
        assert np.max(np.abs(got3 - ref3)) <= bound3, (name, t, mode_name, "logits")
        a, a_ref = int(np.where(order == int(out["action"][b]))[0][0]), int(ep[p + "action"])
        assert a == a_ref or abs(ref3[a] - ref3[a_ref]) <= 2 * bound3, (name, t, mode_name, "action", a, a_ref)
        # the fp32 posterior that enters the state (metabo_gym.py:628-640)
        m_ref, s_ref = ep[p + "mean_final"], ep[p + "std_final"]
        if t > 0:
            m_t, s_t = G.truth_gp(spec.get("kernel", "RBF"), float(ep["kernel_variance"]), ep["kernel_lengthscale"],
                                  float(ep["noise_variance"]), ep[p + "X"], ep[p + "Y"], xi)
            err_ref = max(np.max(np.abs(m_ref - m_t)), np.max(np.abs(s_ref - s_t)))
        else:
            err_ref = 0.0
        floor = 1e-2 * np.sqrt(float(ep["kernel_variance"]))
        rel = lambda a_, b_: np.max(np.abs(a_ - b_) / np.maximum(np.abs(b_), floor))
        assert rel(out["mean"][b].cpu().numpy()[order], m_ref) <= 1e-5 + 4 * err_ref / floor, (name, t, "mean")
        assert rel(out["std"][b].cpu().numpy()[order], s_ref) <= 1e-5 + 4 * err_ref / floor, (name, t, "std")
        full_chain += 1
    # The local Sobol grids are dense around a smooth maximum: their top scores lie within 1e-6..1e-4 (relative) of each
    # other, so the 11-bit modes (and occasionally bf16x3) legitimately pick other near-tied maxima and leave the
    # chain early; the fp32 kernel must carry at least two states through every comparison above.
    if mode_name == "fp32":
        assert full_chain >= 2, (name, mode_name, full_chain)


@pytest.mark.parametrize("name", ["bra", "hm3"])
@pytest.mark.parametrize("mode_name,mode,tol", MODES)
def test_round_without_logits_equals_round_with_logits(golden_dir, name, mode_name, mode, tol, monkeypatch):
    """`af_round(want_logits=False)` -- what bench.py times: selections of all three phases in the tensor-core kernel's
    epilogue, no logit written to HBM -- returns the top-k starts, maxima, action and x_action of the round that
    materialises every logit and selects with topk_kernel / argmax_kernel (MBO_FUSE_TOPK=0), bit for bit, on the
    reference's own episode states (first-index ties like torch.argmax; the fp32 CUDA-core mode has no selection
    epilogue and takes the same path both times)."""
    ep = np.load(os.path.join(golden_dir, "episode_%s.npz" % name))
    spec = json.loads(str(ep["__env_spec__"]))
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, WEIGHTS[name]))
    keep = [int(t) for t in ep["keep"]]
    B, D, T = len(keep), spec["D"], spec["T"]
    outs = []
    for fuse, want in (("0", True), ("1", False)):
        monkeypatch.setenv("MBO_FUSE_TOPK", fuse)
        eng = _engine_for(ep, spec, pw, mode, B)
        dev = eng.dev
        X = np.zeros((B, T, D)); Y = np.zeros((B, T)); inc = np.zeros(B)
        for b, t in enumerate(keep):
            p = "t%02d_" % t
            X[b, :t], Y[b, :t] = ep[p + "X"], ep[p + "Y"]
            inc[b] = float(ep[p + "incumbent"])
        eng.set_data(torch.as_tensor(X, device=dev), torch.as_tensor(Y, device=dev),
                     torch.as_tensor(np.array(keep, dtype=np.int32), device=dev))
        eng.set_episode_scalars(inc, np.array(keep, dtype=np.float32), np.full(B, T, dtype=np.float32))
        out = eng.af_round(deterministic=True, want_logits=want)
        torch.cuda.synchronize()
        eng.policy.check()
        outs.append(out)
    a, b = outs
    assert a["logits"] is not None
    if mode != 0:
        assert b["logits"] is None and b["af_ms"] is None and b["af_ls"] is None
    for key in ("top_ms", "top_ls", "af_maxima", "action", "x_action"):
        assert torch.equal(a[key], b[key]), (name, mode_name, key)


@pytest.mark.parametrize("mode_name", ["fp32"])
def test_gps_discrete_round_on_golden_states(golden_dir, mode_name):
    """C4 (train_metabo_gps.py: Matern-5/2, prior mean = mean(Y), discrete Sobol domain, 2x20 policy with t / T masked
    from the policy): `af_round` on the reference episode's states -> logits and greedy action."""
    ep = np.load(os.path.join(golden_dir, "episode_gps.npz"))
    spec = json.loads(str(ep["__env_spec__"]))
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, WEIGHTS["gps"]))
    keep = [int(t) for t in ep["keep"]]
    B, D, T = len(keep), spec["D"], spec["T"]
    eng = _engine_for(ep, spec, pw, 0, B)
    dev = eng.dev
    X = np.zeros((B, T, D)); Y = np.zeros((B, T)); inc = np.zeros(B)
    for b, t in enumerate(keep):
        p = "t%02d_" % t
        X[b, :t], Y[b, :t] = ep[p + "X"], ep[p + "Y"]
        inc[b] = float(ep[p + "incumbent"])
    eng.set_data(torch.as_tensor(X, device=dev), torch.as_tensor(Y, device=dev),
                 torch.as_tensor(np.array(keep, dtype=np.int32), device=dev))
    eng.set_episode_scalars(inc, np.array(keep, dtype=np.float32), np.full(B, T, dtype=np.float32))
    out = eng.af_round(deterministic=True, want_posterior=True)
    torch.cuda.synchronize()
    for b, t in enumerate(keep):
        p = "t%02d_" % t
        xi = ep[p + "xi_t"]
        ref, got = ep[p + "logits"], out["logits"][b].cpu().numpy()
        scale = np.max(np.abs(ref))
        bound = 1e-5 * scale + 1.5 * _gp_slack(pw, spec, ep, p, t, xi) + 1e-6
        assert np.max(np.abs(got - ref)) <= bound, (t, np.max(np.abs(got - ref)), bound)
        a, a_ref = int(out["action"][b]), int(ep[p + "action"])
        assert a == a_ref or abs(ref[a] - ref[a_ref]) <= 2 * bound, (t, a, a_ref)
        st = eng.state_tensor(out["cand"], out["mean"], out["std"])[b].cpu().numpy()
        ref_state = ep[p + "state"]
        assert np.array_equal(st[:, 2:], ref_state[:, 2:])               # incumbent, t/T, t, T columns: exact
        assert np.max(np.abs(st[:, :2] - ref_state[:, :2])) <= 1e-5 * max(1.0, np.abs(ref_state[:, :2]).max()) + bound


def test_gps_episode_through_vec_env(golden_dir):
    """C4 end to end through VecMetaBO (lane seed 100 = the reference episode's seed): same GP-sample objective
    (y_max on the discrete domain), per-episode length-scale hyper-parameter, and -- free running, deterministic --
    the reference's action sequence up to the first near-tie, with a final reward at least as good as a near-tie
    explains."""
    from metabo_b200.environment.vec_env import VecMetaBO
    from metabo_b200 import gym_compat as gym
    from metabo_b200.policies.policies import NeuralAF
    ep = np.load(os.path.join(golden_dir, "episode_gps.npz"))
    spec = json.loads(str(ep["__env_spec__"]))
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, WEIGHTS["gps"]))
    M, F = spec["cardinality_domain"], int(pw.n_features)
    obs = gym.spaces.Box(low=-1e5, high=1e5, shape=(M, F), dtype=np.float32)
    pi = NeuralAF(obs, gym.spaces.Discrete(M), deterministic=True, options=dict(pw.options, mlp_mode="fp32"))
    pi.load_state_dict({k_: v.clone() for k_, v in pw.sd.items()})
    env = VecMetaBO(1, [100], mlp_mode="fp32", **{k_: v for k_, v in spec.items()})
    env.set_policy(pi)
    env.reset(deterministic=True)
    assert abs(float(env.fb.y_max[0]) - float(ep["y_max"])) <= 1e-9 * max(1.0, abs(float(ep["y_max"])))
    assert np.allclose(env.engine.lengthscale[0].cpu().numpy(), ep["kernel_lengthscale"], rtol=1e-12)
    T = spec["T"]
    rewards, xs = [], []
    for t in range(T):
        r, done = env.step(deterministic=True)
        rewards.append(float(r[0]))
    torch.cuda.synchronize()
    Xg, Xr = env.engine.X[0].cpu().numpy(), ep["actions_all_X"]
    agree = 0
    while agree < T and np.array_equal(Xg[agree], Xr[agree]):
        agree += 1
    assert agree >= 3, agree                   # t = 0..2: empty / tiny GPs, no near-ties in the reference trace
    ref_rewards = ep["rewards"]
    assert np.allclose(rewards[:agree], ref_rewards[:agree], rtol=1e-9, atol=1e-9)
    assert rewards[-1] >= ref_rewards[-1] - 1.0            # neg_log10 regret: within a decade of the reference's final


def _regret_curves(name, golden_dir, mode):
    """evaluate.py:146-152 through the drop-in BatchRecorder: 10 workers (seeds 100..109) x 10 episodes, deterministic"""
    from metabo_b200 import gym_compat as gym
    from metabo_b200.policies.policies import NeuralAF
    from metabo_b200.ppo.batchrecorder import BatchRecorder
    ep = np.load(os.path.join(golden_dir, "episode_%s.npz" % name))
    spec = json.loads(str(ep["__env_spec__"]))
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, WEIGHTS[name]))
    env_id = "Regret-%s-%s-v0" % (name, mode)
    gym.register(id=env_id, entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=spec["T"],
                 kwargs=dict(spec, env_id=env_id))
    opts = dict(pw.options, mlp_mode=mode)
    br = BatchRecorder(size=spec["T"] * 100, env_id=env_id, env_seeds=list(range(100, 110)),
                       policy_fn=lambda o, a, d: NeuralAF(o, a, d, options=opts), n_workers=10, deterministic=True,
                       store_states=False, mlp_mode=mode)
    pi = NeuralAF(br.observation_space, br.action_space, True, options=opts)
    pi.load_state_dict({k_: v.clone() for k_, v in pw.sd.items()})
    br.set_worker_weights(pi)
    br.record_batch(gamma=1.0, lam=1.0)
    r = br.buf["rewards"].cpu().numpy()               # [100, T], worker-major like the reference's memory (:272)
    return r, spec, float(ep["y_max"])


@pytest.mark.parametrize("name,modes", [("bra", ["bf16x3", "fp16", "tf32"]), ("hm3", ["bf16x3", "fp16"]), ("gps", ["fp32"])])
def test_regret_statistics_100_episodes(golden_dir, name, modes):
    """SURVEY 7 hard-part 5: free-running episodes are graded on regret STATISTICS.  Same seeds => same objectives as the
    reference run; median / p30 / p70 of log10(simple regret) per step (plot_results.py:61-63) must agree with the
    reference's within max(0.35 decades, 3 bootstrap standard errors of the reference's own statistic) for the fp32-class
    modes (fp32, bf16x3).  The 11-bit modes (fp16, tf32) get 0.6 decades and no first-step check: on Branin the AF of
    the empty GP is flat to 5e-5 relative (|logit| = 1349, top multistart scores 0.07 apart) -- far inside their 1e-3
    logit tolerance -- so their first action is a different near-tied maximum in 91 % of the episodes and the regret
    curves only agree statistically (measured r2: Branin fp16 p30 at t=10 0.53 decades above the reference, every
    other statistic within 0.35; Hartmann-3 fp16 within 0.35 everywhere)."""
    ref = np.load(os.path.join(golden_dir, "regret_%s.npz" % name))
    R_ref = ref["regret"]
    if name == "gps":
        R_ref = 10.0 ** (-R_ref)                      # that config logs neg_log10 rewards
    rng = np.random.RandomState(0)
    boots = rng.randint(0, 100, size=(400, 100))
    for mode in modes:
        R, spec, _ = _regret_curves(name, golden_dir, mode)
        if name == "gps":
            R = 10.0 ** (-R)
        assert R.shape == R_ref.shape
        lg, lg_ref = np.log10(np.maximum(R, 1e-12)), np.log10(np.maximum(R_ref, 1e-12))
        bad = []
        # same objectives as the reference run (same lane seeds) and a first action that is a function of the weights
        # only (empty GP): the regret after step 1 agrees up to the choice among near-tied local maxima of the AF
        first_same = (np.abs(lg[:, 0] - lg_ref[:, 0]) <= 0.02).mean()
        coarse = mode in ("fp16", "tf32")
        if first_same < 0.9 and not coarse:
            bad.append(("first step", first_same))
        for t in (4, 9, 19, spec["T"] - 1):
            for q in (30, 50, 70):
                s, s_ref = np.percentile(lg[:, t], q), np.percentile(lg_ref[:, t], q)
                se = np.std([np.percentile(lg_ref[bi, t], q) for bi in boots])
                if abs(s - s_ref) > max(0.6 if coarse else 0.35, 3 * se):
                    bad.append((t, q, round(float(s), 3), round(float(s_ref), 3), round(float(se), 3)))
        assert not bad, (name, mode, bad)
