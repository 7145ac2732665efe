"""CPU tests of the host-side drop-in layer: C-ABI surface, gym shim, minibatch partition, GAE,
PPO loss algebra, function-parameter draws, and the 2-rank (gloo) gradient all-reduce equivalence."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import metabo_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    import ctypes
    from metabo_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "metabo_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mbo_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "metabo_b200", "csrc"), "-j8"])
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    bound = {s[0] for s in _lib.SYMBOLS}
    assert declared == bound, declared ^ bound
    assert _lib.lib().mbo_version() >= 100          # loads; no compute call without a GPU


def test_missing_library_fails_loudly(monkeypatch):
    from metabo_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmetabo_b200.so")
    with pytest.raises(_lib.MboError):
        _lib.lib()


def test_engine_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from metabo_b200.engine import AFEngine
    from metabo_b200 import _lib
    with pytest.raises(_lib.MboError):
        AFEngine(1, 2, 30, ["posterior_mean", "posterior_std", "x"], discrete_domain=np.zeros((4, 2)))


def test_gym_shim_registers_and_routes_reference_entry_point():
    from metabo_b200 import gym_compat as gym
    gym.register(id="X-v0", entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=3, kwargs={"a": 1})
    spec = gym.registry.env_specs["X-v0"]
    assert spec.id == "X-v0" and spec._kwargs == {"a": 1} and spec.max_episode_steps == 3
    box = gym.spaces.Box(low=-1e5, high=1e5, shape=(4, 3), dtype=np.float32)
    assert box.contains(np.zeros((4, 3), dtype=np.float32)) and not box.contains(np.zeros((3, 3)))
    assert gym.spaces.Discrete(5).contains(4) and not gym.spaces.Discrete(5).contains(5)


@pytest.mark.parametrize("n,mb", [(1200, 60), (100, 30), (59, 60), (121, 60), (3000, 128)])
def test_minibatch_partition_matches_reference_rule(n, mb):
    """batchrecorder.py:318-333: the last minibatch absorbs the remainder (size in [mb, 2mb))."""
    from metabo_b200.ppo.batchrecorder import BatchRecorder
    br = BatchRecorder.__new__(BatchRecorder)
    br.buf, br.B, br.T = {}, n, 1
    bounds = br._minibatch_bounds(mb)
    ref, pos = [], 0
    while pos < n:                              # the reference loop, restated
        cur = n - pos if pos + 2 * mb > n else mb
        ref.append((pos, pos + cur))
        pos += cur
    assert bounds == ref
    assert bounds[-1][1] == n and all(b - a >= min(mb, n) for a, b in bounds)


def test_gae_device_matches_oracle():
    from metabo_b200.ppo.batchrecorder import gae_device
    rng = np.random.RandomState(0)
    B, T = 5, 30
    r, v = rng.randn(B, T), rng.randn(B, T)
    adv, ret = gae_device(torch.as_tensor(r), torch.as_tensor(v), 0.98, 0.97)
    for b in range(B):
        news = [1] + [0] * (T - 1)
        a_o, r_o = O.gae(r[b], v[b], news, next_new=1, next_value=v[b, -1], gamma=0.98, lam=0.97)
        np.testing.assert_allclose(adv[b].numpy(), a_o, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(ret[b].numpy(), r_o, rtol=1e-12, atol=1e-12)


def _tiny_policy(seed=0):
    from metabo_b200 import gym_compat as gym
    from metabo_b200.policies.policies import NeuralAF
    torch.manual_seed(seed)
    opts = {"activations": "relu", "arch_spec": [16, 16], "use_value_network": True, "arch_spec_value": [8],
            "t_idx": -2, "T_idx": -1}
    obs = gym.spaces.Box(low=-1e5, high=1e5, shape=(11, 6), dtype=np.float32)
    pi = NeuralAF(obs, gym.spaces.Discrete(11), deterministic=False, options=opts)
    for p in pi.parameters():
        p.data.normal_(0, 0.3)
    return pi, opts


def _batch(seed, n):
    rng = np.random.RandomState(seed)
    states = rng.randn(n, 11, 6).astype(np.float32)
    states[:, :, 4] = rng.randint(0, 30, size=(n, 1))
    states[:, :, 5] = 30.0
    return (states, rng.randint(0, 11, size=n), rng.randn(n).astype(np.float32), rng.randn(n).astype(np.float32),
            (rng.randn(n) * 0.1 - 2.0).astype(np.float32))


def test_ppo_loss_matches_oracle_on_cpu():
    """NeuralAF autograd path + ppo_minibatch_loss + normalize_advantages == ppo.py:141-168."""
    from metabo_b200.ppo.ppo import normalize_advantages, ppo_minibatch_loss
    pi, opts = _tiny_policy()
    states, actions, advs, tdl, _ = _batch(1, 24)
    params = {"epsilon": 0.15, "value_coeff": 1.0, "ent_coeff": 0.01}
    sd = {k: v.detach().numpy() for k, v in pi.state_dict().items()}
    pw = O.PolicyWeights(sd, opts, 6)
    pw_old = O.PolicyWeights({k: v * 0.9 for k, v in sd.items()}, opts, 6)
    lo, _ = O.policy_forward(pw_old, states)
    lp_old, _ = O.categorical_logp_entropy(lo, actions)
    ref = O.ppo_loss(pw, pw_old, states, actions, advs, tdl, 0.15, 1.0, 0.01, normalize_advs=True)
    a = normalize_advantages(torch.as_tensor(advs))
    loss, lp, lv, le = ppo_minibatch_loss(pi, torch.as_tensor(states), torch.as_tensor(actions), a, torch.as_tensor(tdl),
                                          lp_old, params, n_global=24)
    assert abs(float(loss) - ref["loss"]) < 1e-5 * max(1.0, abs(ref["loss"]))
    assert abs(float(lp) - ref["loss_ppo"]) < 1e-5 and abs(float(lv) - ref["loss_value"]) < 1e-4
    assert abs(float(le) - ref["loss_ent"]) < 1e-5


_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[3]); sys.path.insert(0, os.path.join(sys.argv[3], "tests"))
from test_host_logic import _tiny_policy, _batch
from metabo_b200.ppo.ppo import normalize_advantages, ppo_minibatch_loss, allreduce_gradients
rank, world = int(sys.argv[1]), 2
os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = sys.argv[2]
dist.init_process_group("gloo", rank=rank, world_size=world)
pi, _ = _tiny_policy()
states, actions, advs, tdl, lp_old = _batch(1, 24)
sl = slice(rank * 12, (rank + 1) * 12)
params = {"epsilon": 0.15, "value_coeff": 1.0, "ent_coeff": 0.01}
a = normalize_advantages(torch.as_tensor(advs[sl]), world)
loss, *_ = ppo_minibatch_loss(pi, torch.as_tensor(states[sl]), torch.as_tensor(actions[sl]), a, torch.as_tensor(tdl[sl]),
                              torch.as_tensor(lp_old[sl]), params, n_global=24)
loss.backward()
allreduce_gradients(list(pi.parameters()), world)
g = torch.cat([p.grad.reshape(-1) for p in pi.parameters()])
if rank == 0:
    np.save(sys.argv[4], g.numpy())
dist.destroy_process_group()
'''


def test_two_rank_gradient_allreduce_equals_single_process(tmp_path):
    """world_size-2 gloo run: sharded minibatch + (advantage-stat, gradient) all-reduce gives the
    single-process gradient on the concatenated minibatch."""
    from metabo_b200.ppo.ppo import normalize_advantages, ppo_minibatch_loss
    pi, _ = _tiny_policy()
    states, actions, advs, tdl, lp_old = _batch(1, 24)
    params = {"epsilon": 0.15, "value_coeff": 1.0, "ent_coeff": 0.01}
    a = normalize_advantages(torch.as_tensor(advs))
    loss, *_ = ppo_minibatch_loss(pi, torch.as_tensor(states), torch.as_tensor(actions), a, torch.as_tensor(tdl),
                                  torch.as_tensor(lp_old), params, n_global=24)
    loss.backward()
    g_ref = torch.cat([p.grad.reshape(-1) for p in pi.parameters()]).numpy()
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    out = str(tmp_path / "g.npy")
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), port, ROOT, out]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=240) == 0
    g = np.load(out)
    np.testing.assert_allclose(g, g_ref, rtol=2e-5, atol=1e-7)


def test_function_draws_follow_reference_rng_order():
    """The host-side parameter draws must consume RandomState like metabo_gym.py:231-304,360-406."""
    from oracle.ref_loader import reference_available, load_reference
    if not reference_available():
        pytest.skip("/root/reference not mounted")
    load_reference()
    from metabo.environment.metabo_gym import MetaBO as RefEnv
    from metabo_b200.environment.functions import FunctionSampler
    base = dict(D=3, features=["posterior_mean", "posterior_std", "x"], T=5, n_init_samples=0, pass_X_to_pi=False,
                kernel_lengthscale=[0.7, 0.3, 0.2], kernel_variance=0.8, noise_variance=1e-9, local_af_opt=True,
                N_MS=100, N_LS=50, k=2, reward_transformation="none", use_prior_mean_function=False)
    for f_type, f_opts in (("HM3-var", {"bound_translation": 0.1, "bound_scaling": 0.1}),
                           ("HM3-var", {"bound_translation": 0.1, "bound_scaling": 0.1, "M": 50})):
        ref = RefEnv(f_type=f_type, f_opts=f_opts, **base)
        ref.seed(7)
        rng = np.random.RandomState(); rng.seed(7)
        smp = FunctionSampler(f_type, f_opts, 3, "cpu")
        for _ in range(3):                         # three consecutive episodes of the same lane
            ref.reset_step_counters(); ref.draw_new_function()
            fb, _ = smp.draw([rng])
            x = np.random.RandomState(1).rand(6, 3)
            y_ref = ref.f(x)[:, 0]
            y = torch.stack([fb.eval(torch.as_tensor(x[i:i + 1])) for i in range(6)])[:, 0].numpy()
            np.testing.assert_allclose(y, y_ref, rtol=1e-12, atol=1e-12)
            assert abs(float(fb.y_max[0]) - float(np.asarray(ref.y_max).reshape(-1)[0])) < 1e-12


def test_ppo_update_abi_argument_checks_and_workspace_layout():
    """mbo_ppo_* entry points without a GPU: size queries, the debug layout and the argument checks that
    fire before any CUDA call (status code + mbo_last_error, no exception across the boundary)."""
    import ctypes as C
    from metabo_b200 import _lib as L
    lib = L.lib()
    n60 = lib.mbo_ppo_workspace_bytes(60, 966)
    assert lib.mbo_ppo_workspace_bytes(0, 966) == 0 and lib.mbo_ppo_workspace_bytes(61, 966) > n60 > 60 * 966 * 208 * 4 * 8
    off = (C.c_size_t * 12)()
    assert lib.mbo_ppo_debug_offsets(60, 966, off) == 0
    rows = ((60 * 966 + 127) // 128) * 128
    o = list(off)
    assert all(o[i + 1] - o[i] == rows * 208 * 4 for i in range(3))            # h1..h4 are consecutive [R_pad, 208] fp32 arrays
    assert all(o[i + 1] - o[i] >= rows * 208 * 4 for i in range(4, 7)) and max(o) < n60
    assert lib.mbo_ppo_debug_offsets(0, 966, off) == L.lib().mbo_ppo_debug_offsets(60, 966, None) < 0
    null = C.c_void_p(0)
    cols = (C.c_int32 * 6)(*range(6))
    arr = (C.c_void_p * 5)()
    rc = lib.mbo_ppo_policy_grad(0, 966, 6, 6, cols, 200, null, arr, arr, null, null, null, 0.15, 0.01, 1.0, arr, arr, null, null, 0, null)
    assert rc == -1 and b"E, M must be > 0" in lib.mbo_last_error()
    rc = lib.mbo_ppo_policy_grad(4, 10, 6, 6, cols, 300, null, arr, arr, null, null, null, 0.15, 0.01, 1.0, arr, arr, null, null, 0, null)
    assert rc == -1 and b"hidden width" in lib.mbo_last_error()
    rc = lib.mbo_ppo_policy_grad(4, 10, 6, 9, cols, 200, null, arr, arr, null, null, null, 0.15, 0.01, 1.0, arr, arr, null, null, 0, null)
    assert rc == -1 and b"policy features" in lib.mbo_last_error()
    assert lib.mbo_ppo_value_workspace_bytes(60, 200) >= 6 * 60 * 200 * 4 and lib.mbo_ppo_value_workspace_bytes(0, 200) == 0
    rc = lib.mbo_ppo_value_grad(4, 60, 4, 5, 300, null, arr, arr, null, 0.1, arr, arr, null, null, null, 0, null)
    assert rc == -1 and b"hidden width" in lib.mbo_last_error()
    sizes = (C.c_int64 * 1)(4)
    rc = lib.mbo_adam_step(0, arr, arr, arr, arr, sizes, arr, 1e-3, 0.9, 0.999, 1e-8, null)
    assert rc == -1 and b"tensors per call" in lib.mbo_last_error()
    assert lib.mbo_ppo_loss_sums(0, null, null, 1.0, 1.0, 0.01, null, null) == -1
    assert lib.mbo_ppo_prepare_minibatch(0, 10, null, null, null, null, null, null, 1, null, null, null, null, null, null, null) == -1
    assert lib.mbo_ppo_adv_stats(0, null, null, null, null, null) == -1


def test_comm_abi_argument_checks():
    """mbo_comm_* (peer-memory gradient exchange): argument checks that fire before any CUDA call."""
    import ctypes as C
    from metabo_b200 import _lib as L
    lib = L.lib()
    h = C.c_void_p()
    assert lib.mbo_comm_create(2, 2, 1024, C.byref(h)) == -1 and b"rank 2 / world 2" in lib.mbo_last_error()
    assert lib.mbo_comm_create(0, 17, 1024, C.byref(h)) == -1
    assert lib.mbo_comm_create(0, 1, 0, C.byref(h)) == -1 and b"slot_bytes" in lib.mbo_last_error()
    assert lib.mbo_comm_export(None, None) == -1 and lib.mbo_comm_connect(None, None) == -1
    assert lib.mbo_comm_allreduce(None, 0, None, None, 0, None) == -1
    arr = (C.c_void_p * 1)()
    sizes = (C.c_int64 * 1)(4)
    assert lib.mbo_comm_allreduce_adam(None, None, None, 1, arr, arr, arr, sizes, arr, 1e-3, 0.9, 0.999, 1e-8, None) == -1
    assert lib.mbo_comm_status(None, None, None, None) == -1
    lib.mbo_comm_destroy(None)                      # no-op


def test_shard_bounds_partition_every_minibatch_exactly():
    """Data-parallel minibatches (ppo.py:127-178 sharded over ranks): for every world size the local shares of a step
    add up to the reference's global minibatch (batchrecorder.py:326-330 incl. the merged remainder) and every rank
    consumes exactly its n_local transitions per epoch."""
    from metabo_b200.ppo.ppo import minibatch_bounds, shard_bounds
    assert minibatch_bounds(1200, 60) == [(60 * i, 60 * (i + 1)) for i in range(20)]
    assert minibatch_bounds(130, 60) == [(0, 60), (60, 130)]                 # remainder joins the last minibatch
    for world in (1, 2, 3, 4, 8):
        for n_local, mb in ((150, 60), (1200 // world if 1200 % world == 0 else 120, 60), (30720, 60), (130, 64), (96, 8)):
            per_rank = [shard_bounds(n_local, mb, r, world) for r in range(world)]
            gb = minibatch_bounds(n_local * world, mb)
            for r, (loc, ng) in enumerate(per_rank):
                assert ng == [b - a for a, b in gb]
                assert loc[0][0] == 0 and loc[-1][1] == n_local
                assert all(loc[i][1] == loc[i + 1][0] for i in range(len(loc) - 1))
                assert all(b - a >= mb // world for a, b in loc) and all(b - a <= -(-max(ng) // world) for a, b in loc)
            for j, (a, b) in enumerate(gb):
                assert sum(per_rank[r][0][j][1] - per_rank[r][0][j][0] for r in range(world)) == b - a


def test_env_step_abi_argument_checks():
    """mbo_env_step / mbo_state_pack reject bad arguments before any CUDA call (status code + mbo_last_error)."""
    import ctypes as C
    from metabo_b200 import _lib
    lib = _lib.lib()
    p = C.c_void_p(256)                      # never dereferenced: the checks fire first
    args = lambda family, B, T, D, t_idx, reward: (family, B, T, D, t_idx, p, p, p, p, p, p, p, p, p, reward, p, None,
                                                   C.c_float(1.0), C.c_float(30.0), C.c_float(-1.0), None)
    assert lib.mbo_env_step(*args(7, 4, 30, 2, 0, 2)) == -1 and b"unknown objective family" in lib.mbo_last_error()
    assert lib.mbo_env_step(*args(_lib.ENV_BRANIN, 4, 30, 3, 0, 2)) == -1 and b"dimensional domain" in lib.mbo_last_error()
    assert lib.mbo_env_step(*args(_lib.ENV_HARTMANN3, 4, 30, 3, 30, 2)) == -1 and b"t_idx" in lib.mbo_last_error()
    assert lib.mbo_env_step(*args(_lib.ENV_HARTMANN3, 4, 30, 3, 0, 5)) == -1 and b"reward transformation" in lib.mbo_last_error()
    assert lib.mbo_state_pack(4, None, 6, None, p, p, p, p, None) == -1 and b"mbo_state_pack" in lib.mbo_last_error()
    assert set(_lib.REWARDS) == {"none", "neg_linear", "neg_log10"}          # metabo_gym.py:679-704
