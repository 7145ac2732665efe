"""ORACLE (test infrastructure, never shipped/imported by the product path).

CPU restatement (numpy fp64 + torch-CPU fp32) of the reference hot path, one BO episode at a
time, exactly as the reference executes it:

  get_state            metabo/environment/metabo_gym.py:624-677   (feature packing, fp32 cast)
  get_incumbent        metabo/environment/metabo_gym.py:723-730
  get_af_maxima        metabo/environment/metabo_gym.py:706-721   (multistart -> local grids -> top-k)
  optimize_AF          metabo/environment/metabo_gym.py:615-622
  create_uniform_grid  metabo/environment/util.py:24-32
  get_cube_around      metabo/environment/util.py:45-51
  NeuralAF.forward     metabo/policies/policies.py:104-125  +  MLP.forward  metabo/policies/mlp.py:52-61
  NeuralAF.act         metabo/policies/policies.py:135-148
  predict_vals_logps_ents  metabo/policies/policies.py:150-161  (torch Categorical semantics)
  GAE                  metabo/ppo/batchrecorder.py:117-137
  PPO loss             metabo/ppo/ppo.py:141-168

The GP numerics come from oracle/gp_oracle.py.  tests/test_oracle_vs_reference.py checks this
file against the reference's own classes (imported through oracle/ref_loader.py) whenever
/root/reference is mounted, and tests/golden/*.npz hold vectors minted from the reference
(oracle/make_golden.py).
"""
import json
import numpy as np
import torch

from . import gp_oracle
from .sobol import i4_sobol_generate

FEATURE_ORDER = ["posterior_mean", "posterior_std", "x", "incumbent", "timestep_perc", "timestep", "budget"]


# ----------------------------------------------------------------------------------------------
# grids (util.py:24-51, metabo_gym.py:74-91)
# ----------------------------------------------------------------------------------------------
def create_uniform_grid(domain, N_samples_dim):
    D = domain.shape[0]
    x_grid = [np.linspace(domain[i, 0], domain[i, 1], int(N_samples_dim)) for i in range(D)]
    X_mesh = np.meshgrid(*x_grid)
    return np.vstack(X_mesh).reshape((D, -1)).T


def get_cube_around(x, diam, domain):
    cube = np.zeros(domain.shape)
    cube[:, 0] = np.max((x - 0.5 * diam, domain[:, 0]), axis=0)
    cube[:, 1] = np.min((x + 0.5 * diam, domain[:, 1]), axis=0)
    return cube


def scale_from_unit_square_to_domain(X, domain):
    return X * np.ptp(domain, axis=1) + domain[:, 0]


class Grids:
    """Candidate tables of one env config (metabo_gym.py:66-91)."""

    def __init__(self, D, N_MS=None, N_LS=None, k=None, cardinality_domain=None):
        self.D = D
        self.domain = np.stack([np.zeros(D), np.ones(D)], axis=1)
        if cardinality_domain is None:
            per_dim = int(np.floor(N_MS ** (1 / D)))
            self.N_MS_per_dim = per_dim
            self.multistart_grid = create_uniform_grid(self.domain, per_dim)
            self.N_MS = self.multistart_grid.shape[0]
            self.k = k
            self.N_LS = N_LS
            self.local_search_grid = i4_sobol_generate(D, N_LS)
            self.af_max_search_diam = 2 * 1 / per_dim
            self.discrete = False
        else:
            self.xi_t = i4_sobol_generate(D, cardinality_domain)
            self.discrete = True

    def local_grids(self, startpoints):
        return np.concatenate([scale_from_unit_square_to_domain(
            self.local_search_grid, get_cube_around(x, self.af_max_search_diam, self.domain))
            for x in startpoints], axis=0)


# ----------------------------------------------------------------------------------------------
# policy (policies.py / mlp.py)
# ----------------------------------------------------------------------------------------------
class PolicyWeights:
    """NeuralAF parameters as plain tensors; state_dict keys as shipped
    (policy_net.fc.{i}.{weight,bias}, value_net.fc.{i}.{weight,bias})."""

    def __init__(self, state_dict, options, n_features):
        self.sd = {k: torch.as_tensor(np.asarray(v), dtype=torch.float32) for k, v in state_dict.items()}
        self.options = dict(options)
        self.n_features = n_features
        self.act = options["activations"]
        self.mask = [True] * n_features
        if options.get("exclude_t_from_policy", False):
            self.mask[options["t_idx"]] = False
        if options.get("exclude_T_from_policy", False):
            self.mask[options["T_idx"]] = False
        self.use_value_network = bool(options.get("use_value_network", False))

    @staticmethod
    def from_npz(path):
        z = np.load(path, allow_pickle=False)
        opts = json.loads(str(z["__policy_options__"]))
        nf = int(z["__n_features__"])
        sd = {k: z[k] for k in z.files if not k.startswith("__")}
        return PolicyWeights(sd, opts, nf)

    def layers(self, net):
        i, out = 0, []
        while "%s.fc.%d.weight" % (net, i) in self.sd:
            out.append((self.sd["%s.fc.%d.weight" % (net, i)], self.sd["%s.fc.%d.bias" % (net, i)]))
            i += 1
        return out


def _mlp(layers, x, act):
    f = torch.relu if act == "relu" else torch.tanh
    for W, b in layers[:-1]:
        x = f(torch.nn.functional.linear(x, W.to(x.dtype), b.to(x.dtype)))
    W, b = layers[-1]
    return torch.nn.functional.linear(x, W.to(x.dtype), b.to(x.dtype))


def policy_forward(pw, states, dtype=torch.float32):
    """NeuralAF.forward: states [B,M,F] -> (logits [B,M], values [B])."""
    states = torch.as_tensor(np.asarray(states), dtype=torch.float32).to(dtype)
    assert states.dim() == 3 and states.shape[-1] == pw.n_features
    with torch.no_grad():
        logits = _mlp(pw.layers("policy_net"), states[:, :, pw.mask], pw.act).squeeze(2)
        if pw.use_value_network:
            t_idx, T_idx = pw.options["t_idx"], pw.options["T_idx"]
            tT = states[:, [0], [t_idx, T_idx]]
            values = _mlp(pw.layers("value_net"), tT, pw.act).squeeze(1)
        else:
            values = torch.zeros(states.shape[0], dtype=dtype)
    return logits, values


def categorical_logp_entropy(logits, actions):
    """torch.distributions.Categorical(logits=...).log_prob / .entropy (policies.py:157-159)."""
    logits = torch.as_tensor(logits)
    norm = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
    actions = torch.as_tensor(np.asarray(actions)).long()
    logp = norm.gather(-1, actions[:, None]).squeeze(-1)
    min_real = torch.finfo(norm.dtype).min
    nl = torch.clamp(norm, min=min_real)
    ent = -(nl * torch.exp(nl)).sum(-1)
    return logp, ent


# ----------------------------------------------------------------------------------------------
# environment-side state assembly
# ----------------------------------------------------------------------------------------------
class EpisodeGP:
    """The GP surrogate + bookkeeping of one BO episode (metabo_gym.py:141-153, 549-613)."""

    def __init__(self, D, kernel, variance, lengthscale, noise_variance, use_prior_mean=False, y_min=0.0,
                 route="gpy"):
        self.D, self.kernel, self.variance = D, kernel, float(variance)
        self.lengthscale = np.broadcast_to(np.asarray(lengthscale, dtype=np.float64), (D,)).copy()
        self.noise_variance = float(noise_variance)
        self.use_prior_mean = use_prior_mean
        self.y_min = y_min
        self.X = np.zeros((0, D))
        self.Y = np.zeros((0,))
        self.route = route

    def set_data(self, X, Y):
        self.X = np.asarray(X, dtype=np.float64).reshape(-1, self.D)
        self.Y = np.asarray(Y, dtype=np.float64).reshape(-1)

    def prior_mean(self):
        return float(np.mean(self.Y)) if (self.use_prior_mean and self.Y.size) else 0.0

    def eval_gp(self, Xstar):
        if self.Y.size == 0:
            return gp_oracle.empty_gp_posterior(self.variance, Xstar.shape[0])
        fn = gp_oracle.gpy_route_posterior if self.route == "gpy" else gp_oracle.truth_gp
        return fn(self.kernel, self.variance, self.lengthscale, self.noise_variance,
                  self.X, self.Y, Xstar, self.prior_mean())

    def incumbent(self):
        return float(np.max(self.Y)) if self.Y.size else float(self.y_min)


def get_state(gp, Xstar, features, t, T, T_training=None):
    """metabo_gym.py:624-677: fixed column order, float32."""
    M, D = Xstar.shape
    cols = []
    mean, std = gp.eval_gp(Xstar)
    if "posterior_mean" in features:
        cols.append(mean.reshape(M, 1))
    if "posterior_std" in features:
        cols.append(std.reshape(M, 1))
    if "x" in features:
        cols.append(Xstar)
    if "incumbent" in features:
        cols.append(np.full((M, 1), gp.incumbent()))
    if "timestep_perc" in features:
        cols.append(np.full((M, 1), t / T))
    if "timestep" in features:
        cols.append(np.full((M, 1), min(t, T_training) if T_training is not None else t))
    if "budget" in features:
        cols.append(np.full((M, 1), T_training if T_training is not None else T))
    return np.concatenate(cols, axis=1).astype(np.float32)


def n_features(features, D):
    return sum(D if f == "x" else 1 for f in FEATURE_ORDER if f in features)


def af_round(gp, grids, pw, features, t, T, T_training=None, dtype=torch.float32):
    """One AF-optimisation round of an episode (optimize_AF + get_state(xi_t) + greedy act).
    Returns a dict with every intermediate the parity tests compare."""
    out = {}
    if grids.discrete:
        xi_t = grids.xi_t
    else:
        s_ms = get_state(gp, grids.multistart_grid, features, t, T, T_training)
        af_ms = policy_forward(pw, s_ms[None], dtype)[0][0].numpy()
        top_ms = np.argsort(-af_ms, kind="stable")[:grids.k]
        starts = grids.multistart_grid[top_ms]
        local = grids.local_grids(starts)
        s_ls = get_state(gp, local, features, t, T, T_training)
        af_ls = policy_forward(pw, s_ls[None], dtype)[0][0].numpy()
        top_ls = np.argsort(-af_ls, kind="stable")[:grids.k]
        af_maxima = local[top_ls]
        xi_t = np.concatenate((af_maxima, grids.multistart_grid), axis=0)
        out.update(state_ms=s_ms, af_ms=af_ms, top_ms=top_ms, local=local, state_ls=s_ls, af_ls=af_ls,
                   top_ls=top_ls)
    state = get_state(gp, xi_t, features, t, T, T_training)
    logits, values = policy_forward(pw, state[None], dtype)
    out.update(xi_t=xi_t, state=state, logits=logits[0].numpy(), value=float(values[0]),
               action=int(torch.argmax(logits[0])))
    return out


# ----------------------------------------------------------------------------------------------
# rollout bookkeeping: GAE (batchrecorder.py:117-137) and the PPO loss (ppo.py:141-168)
# ----------------------------------------------------------------------------------------------
def gae(rewards, values, news, next_new, next_value, gamma, lam):
    n = len(rewards)
    adv = np.zeros(n)
    tdlamret = np.zeros(n)
    next_adv = 0.0
    for i in reversed(range(n)):
        nonterminal = 1 - next_new
        delta = -values[i] + rewards[i] + gamma * nonterminal * next_value
        adv[i] = next_adv = delta + lam * gamma * nonterminal * next_adv
        tdlamret[i] = adv[i] + values[i]
        next_new = news[i]
        next_value = values[i]
    return adv, tdlamret


def ppo_loss(pw, pw_old, states, actions, advs, tdlamrets, epsilon, value_coeff, ent_coeff,
             normalize_advs=True):
    """Forward value of the PPO objective of one minibatch (no autograd; ppo.py:141-168)."""
    advs = torch.as_tensor(np.asarray(advs), dtype=torch.float32)
    tdl = torch.as_tensor(np.asarray(tdlamrets), dtype=torch.float32)
    if normalize_advs:
        s = torch.std(advs, unbiased=False)
        if not s == 0 and not torch.isnan(s):
            advs = (advs - torch.mean(advs)) / s
    lo, _ = policy_forward(pw_old, states)
    ln, vn = policy_forward(pw, states)
    lp_old, _ = categorical_logp_entropy(lo, actions)
    lp, ent = categorical_logp_entropy(ln, actions)
    ratios = torch.exp(lp - lp_old)
    clipped = ratios.clamp(1 - epsilon, 1 + epsilon)
    loss_ppo = -torch.mean(torch.min(ratios * advs, clipped * advs))
    loss_value = torch.mean((vn - tdl) ** 2)
    loss_ent = -torch.mean(ent)
    return dict(loss=float(loss_ppo + value_coeff * loss_value + ent_coeff * loss_ent),
                loss_ppo=float(loss_ppo), loss_value=float(loss_value), loss_ent=float(loss_ent))


def ppo_policy_grads(layers, states, actions, advs, logp_old, epsilon, ent_coeff, n_global, masks=None,
                     dtype=torch.float64):
    """Gradient of loss_ppo + ent_coeff * loss_ent (ppo.py:148-168, both divided by the minibatch size n_global)
    w.r.t. the policy net's (W, b) by torch-CPU autograd, with Categorical(logits) restated as in
    policies.py:150-161.  states [E, M, F] (policy columns only), layers = [(W, b)] * 5 in torch layout.
    `masks` (optional, one bool [E*M, H] array per hidden layer) replaces the ReLU decisions: the backward of a
    reduced-precision forward is only comparable element by element under that forward's own decisions (a
    pre-activation within rounding error of 0 flips its ReLU and changes dZ by its full value).
    Returns dict(dW, db, logits, logp, entropy, ratio, loss_ppo, loss_ent, dlogits, hs)."""
    Ws = [torch.as_tensor(np.asarray(w), dtype=dtype).clone().requires_grad_() for w, _ in layers]
    bs = [torch.as_tensor(np.asarray(b), dtype=dtype).clone().requires_grad_() for _, b in layers]
    x = torch.as_tensor(np.asarray(states), dtype=dtype)
    E, M, F = x.shape
    h = x.reshape(E * M, F)
    hs = []
    for l in range(len(Ws) - 1):
        z = h @ Ws[l].t() + bs[l]
        h = torch.relu(z) if masks is None else z * torch.as_tensor(np.asarray(masks[l]), dtype=dtype)
        hs.append(h)
    logits = (h @ Ws[-1].t() + bs[-1]).reshape(E, M)
    logits.retain_grad()
    a = torch.as_tensor(np.asarray(actions), dtype=torch.int64)
    norm = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
    logp = norm.gather(-1, a[:, None]).squeeze(-1)
    ent = -(norm * torch.exp(norm)).sum(-1)
    adv = torch.as_tensor(np.asarray(advs), dtype=dtype)
    ratio = torch.exp(logp - torch.as_tensor(np.asarray(logp_old), dtype=dtype))
    clipped = ratio.clamp(1 - epsilon, 1 + epsilon)
    loss_ppo = -torch.sum(torch.min(ratio * adv, clipped * adv)) / n_global
    loss_ent = -torch.sum(ent) / n_global
    (loss_ppo + ent_coeff * loss_ent).backward()
    return dict(dW=[w.grad for w in Ws], db=[b.grad for b in bs], logits=logits.detach(), logp=logp.detach(),
                entropy=ent.detach(), ratio=ratio.detach(), loss_ppo=loss_ppo.item(), loss_ent=loss_ent.item(),
                dlogits=logits.grad, hs=[t.detach() for t in hs])


# ----------------------------------------------------------------------------------------------
# objective functions needed for full-episode checks (objectives.py:26-93, 165-235)
# ----------------------------------------------------------------------------------------------
def bra(x):
    a, b, c, r, s, t = 1, 5.1 / (4 * np.pi ** 2), 5 / np.pi, 6, 10, 1 / (8 * np.pi)
    x1 = x[:, 0] * 15. - 5.
    x2 = x[:, 1] * 15.
    f = a * (x2 - b * x1 ** 2 + c * x1 - r) ** 2 + s * (1 - t) * np.cos(x1) + s
    f = 1 / 51.44 * (f - 54.44)      # objectives.py:46-49 normalisation
    return (-f).reshape(x.shape[0], 1)


def bra_var(x, t, s):
    t_range = np.array([[-0.12, 0.87], [-0.81, 0.18]])
    t = np.clip(t, t_range[:, 0], t_range[:, 1])
    return s * bra(x - t)


def bra_max(t, s):
    t_range = np.array([[-0.12, 0.87], [-0.81, 0.18]])
    pos = np.array([[(-np.pi + 5.) / 15., 12.275 / 15.]])
    return float(s * bra(pos)[0, 0]), pos + np.clip(t, t_range[:, 0], t_range[:, 1])


def hm3(x):
    alpha = np.array([1.0, 1.2, 3.0, 3.2])
    A = np.array([[3.0, 10, 30], [0.1, 10, 35], [3.0, 10, 30], [0.1, 10, 35]])
    P = 1e-4 * np.array([[3689, 1170, 2673], [4699, 4387, 7470], [1091, 8732, 5547], [381, 5743, 8828]])
    B = (x.reshape(x.shape[0], 1, -1) - P) ** 2
    C = np.exp(-np.einsum("ijk->ij", A * B))
    f = -np.einsum("i, ki", alpha, C)
    f = 1 / 0.95 * (f - (-0.93))
    return (-f).reshape(-1, 1)


def hm3_var(x, t, s):
    t_range = np.array([[-0.11, 0.88], [-0.55, 0.44], [-0.85, 0.14]])
    t = np.clip(t, t_range[:, 0], t_range[:, 1])
    return s * hm3(x - t)


def hm3_max(t, s):
    t_range = np.array([[-0.11, 0.88], [-0.55, 0.44], [-0.85, 0.14]])
    pos = np.array([[0.114614, 0.555649, 0.852547]])
    return float(s * hm3(pos)[0, 0]), pos + np.clip(t, t_range[:, 0], t_range[:, 1])


# ---------------------------------------------------------------------------------------------
# Closed-form baseline acquisition functions (metabo/policies/policies.py:181-307), restated on the float32 state
# with the reference's dtype flow: differences and z in float32, pdf / cdf in float64 (scipy.stats.norm).
# ---------------------------------------------------------------------------------------------
def ei_af(mean, std, incumbent):
    """EI.af, policies.py:249-261.  mean, std, incumbent float32 [M] -> float64 [M] (0 where std == 0)."""
    from scipy.stats import norm
    mean, std, incumbent = (np.asarray(a, dtype=np.float32) for a in (mean, std, incumbent))
    mask = std != 0.0
    out = np.zeros(mean.shape, dtype=np.float64)
    d = (mean - incumbent)[mask]
    z = (d / std[mask]).astype(np.float64)
    out[mask] = d * norm.cdf(z) + std[mask] * norm.pdf(z)
    return out


def pi_af(mean, std, incumbent, xi):
    """PI.af, policies.py:289-301."""
    from scipy.stats import norm
    mean, std, incumbent = (np.asarray(a, dtype=np.float32) for a in (mean, std, incumbent))
    mask = std != 0.0
    out = np.zeros(mean.shape, dtype=np.float64)
    z = ((mean[mask] - (incumbent[mask] + xi).astype(np.float32)) / std[mask]).astype(np.float64)
    out[mask] = norm.cdf(z)
    return out


def ucb_af(mean, std, kappa, timestep=None, D=None, delta=None):
    """UCB.af / compute_kappa, policies.py:201-229 (kappa == "gp_ucb": the GP-UCB schedule on timestep + 1)."""
    mean, std = np.asarray(mean, dtype=np.float32), np.asarray(std, dtype=np.float32)
    if kappa == "gp_ucb":
        t = np.asarray(timestep, dtype=np.float32) + 1
        tau = 2 * np.log(t.astype(np.float64) ** (D / 2 + 2) * np.pi ** 2 / (3 * delta))
        k = np.sqrt(tau)
    else:
        k = float(kappa)
    return mean.astype(np.float64) + k * std.astype(np.float64)
