"""GPU tests of the peer-memory exchange (csrc/comm.cu, `mbo_comm_*`) and of the data-parallel PPO update built on it.
The driver's GPU tier has ONE device, so the ranks are two PROCESSES on cuda:0 (CUDA IPC works between processes on the
same device; the two contexts time-slice, the kernels' peer waits are bounded); `tools/ppo_multi_gpu.py` is the same
check under torchrun on 2-8 real GPUs (profiles/r02_*).  Rendezvous: gloo on 127.0.0.1."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
rank, world, port, root, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
sys.path.insert(0, root)
import numpy as np, torch, torch.distributed as dist
from metabo_b200 import ops
torch.cuda.set_device(0)
dev = torch.device("cuda:0")
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + port, rank=rank, world_size=world)
n = 243602
comm = ops.Comm.from_process_group(4 * n, dev)
res = {}
# (1) plain all-reduce, fp32 and fp64, several consecutive calls (slot alternation), odd sizes
g = torch.Generator(device="cpu"); g.manual_seed(100 + rank)
for it, cnt in enumerate([7, 1000, n, 12345, n]):
    x = torch.randn(cnt, generator=g).to(dev)
    y = comm.allreduce(x.clone())
    res["f32_%d" % it] = y.cpu().numpy()
    res["f32_in_%d" % it] = x.cpu().numpy()
xd = torch.randn(60, generator=g, dtype=torch.float64).to(dev)
res["f64"] = comm.allreduce(xd.clone()).cpu().numpy(); res["f64_in"] = xd.cpu().numpy()
# (2) fused all-reduce + Adam == mbo_adam_step on the exchanged gradient, inside a replayed CUDA graph
shapes = [(200, 6), (200,), (200, 200), (200,), (1, 200), (1,)]
gw = torch.Generator(device="cpu"); gw.manual_seed(7)             # identical parameters on both ranks
P = [torch.randn(s, generator=gw).to(dev).requires_grad_() for s in shapes]
Q = [p.detach().clone().requires_grad_() for p in P]
tot = sum(p.numel() for p in P)
flat = torch.zeros(tot, device=dev); flat_sum = torch.zeros(tot, device=dev)
off = 0
for p, q in zip(P, Q):
    p.grad = flat[off:off + p.numel()].view_as(p); q.grad = flat_sum[off:off + p.numel()].view_as(q); off += p.numel()
opt = ops.FusedAdam(P, lr=1e-2); opt.set_comm(comm, flat)
ref = ops.FusedAdam(Q, lr=1e-2)
def draw(step):
    gg = torch.Generator(device="cpu"); gg.manual_seed(1000 * step + rank)
    return torch.randn(tot, generator=gg).to(dev)
flat.copy_(draw(0)); opt.step()                                    # eager first step allocates the optimiser state
flat_sum.copy_(comm.allreduce(draw(0))); ref.step()
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
static = torch.zeros(tot, device=dev)
with torch.cuda.graph(gr):
    flat.copy_(static); opt.step()
for step in range(1, 6):
    static.copy_(draw(step)); gr.replay()
    flat_sum.copy_(comm.allreduce(draw(step))); ref.step()
torch.cuda.synchronize()
res["adam_fused"] = torch.cat([p.detach().reshape(-1) for p in P]).cpu().numpy()
res["adam_ref"] = torch.cat([q.detach().reshape(-1) for q in Q]).cpu().numpy()
res["epoch"] = np.array(comm.check())
np.savez(out % rank, **res)
dist.barrier()
dist.destroy_process_group()
'''


def _spawn(script_text, tmp_path, world, extra=()):
    script = tmp_path / "w.py"
    script.write_text(script_text)
    port = str(29500 + os.getpid() % 2000)
    out = str(tmp_path / "r%d.npz")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), port, ROOT, out, *extra]) for r in range(world)]
    try:
        for p in procs:
            assert p.wait(timeout=600) == 0
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    return [np.load(out % r) for r in range(world)]


def test_peer_memory_allreduce_and_fused_adam_two_ranks(tmp_path):
    r0, r1 = _spawn(_WORKER, tmp_path, 2)
    for it in range(5):
        want = r0["f32_in_%d" % it] + r1["f32_in_%d" % it]            # rank order 0 + 1, one fp32 addition: exact
        assert np.array_equal(r0["f32_%d" % it], want) and np.array_equal(r1["f32_%d" % it], want), it
    want = r0["f64_in"] + r1["f64_in"]
    assert np.array_equal(r0["f64"], want) and np.array_equal(r1["f64"], want)
    # fused exchange + Adam: bitwise the unfused sequence (all-reduce, then mbo_adam_step), identical on both ranks
    assert np.array_equal(r0["adam_fused"], r0["adam_ref"])
    assert np.array_equal(r0["adam_fused"], r1["adam_fused"])
    assert int(r0["epoch"]) == int(r1["epoch"]) == 5 + 1 + 1 + 6 + 5


_PPO_WORKER = r'''
import os, sys
rank, world, port, root, out, logdir = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5], sys.argv[6]
sys.path.insert(0, root)
import numpy as np, torch, torch.distributed as dist
torch.cuda.set_device(0)
if world > 1:
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + port, rank=rank, world_size=world)
from metabo_b200 import gym_compat as gym
from metabo_b200.policies.policies import NeuralAF
from metabo_b200.ppo.ppo import PPO
spec = {"env_id": "MetaBO-DP-v0", "D": 2, "f_type": "BRA-var", "f_opts": {"bound_translation": 0.1, "bound_scaling": 0.1, "M": 50},
        "features": ["posterior_mean", "posterior_std", "timestep", "budget", "x"], "T": 5, "n_init_samples": 0,
        "pass_X_to_pi": False, "kernel_lengthscale": [0.235, 0.578], "kernel_variance": 2.0, "noise_variance": 8.9e-16,
        "use_prior_mean_function": False, "local_af_opt": True, "N_MS": 100, "N_LS": 50, "k": 3, "reward_transformation": "neg_log10"}
gym.register(id=spec["env_id"], entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=5, kwargs=spec)
params = {"batch_size": 120, "max_steps": 360, "minibatch_size": 15, "n_epochs": 2, "lr": 1e-3, "epsilon": 0.15,
          "value_coeff": 1.0, "ent_coeff": 0.01, "gamma": 0.98, "lambda": 0.98, "loss_type": "GAElam", "normalize_advs": True,
          "n_workers": 4, "env_id": spec["env_id"], "seed": 0, "env_seeds": [0, 1, 2, 3],
          "policy_options": {"activations": "relu", "arch_spec": [200] * 4, "use_value_network": True, "t_idx": -2, "T_idx": -1,
                             "arch_spec_value": [200] * 4}}
ppo = PPO(policy_fn=lambda o, a, d: NeuralAF(o, a, d, options=params["policy_options"]), params=params,
          logpath=os.path.join(logdir, "log"), save_interval=100, verbose=True)
ppo.train()
w = torch.cat([p.detach().reshape(-1) for p in ppo.pi.parameters()]).cpu().numpy()
losses = [float(x.split("=")[1]) for line in ppo.iter_log_str.splitlines() if "loss_ppo" in line for x in line.split(",")[1:]]
np.savez(out % rank, w=w, losses=np.array(losses), n_optsteps=np.array(ppo.stats["n_optsteps"]),
         graphed=np.array(int(ppo._ep is not None and ppo._ep.get("graph") is not None)),
         rewards=ppo.batch_recorder.buf["rewards"].cpu().numpy())
if world > 1:
    dist.barrier(); dist.destroy_process_group()
'''


def test_data_parallel_ppo_two_ranks_identical_weights(tmp_path):
    """PPO.train under world = 2 (minibatch 15 = 8 + 7 per step: uneven shares): 3 iterations x 2 epochs x 8 steps through
    the captured data-parallel epoch graph; weights bit-identical on both ranks, and the same rollouts as the
    single-process run of the same seeds in iteration 0 (lanes keep their seeds under sharding)."""
    r0, r1 = _spawn(_PPO_WORKER, tmp_path, 2, extra=(str(tmp_path),))
    assert int(r0["n_optsteps"]) == int(r1["n_optsteps"]) == 3 * 2 * 8
    assert int(r0["graphed"]) == 1 and int(r1["graphed"]) == 1
    assert np.array_equal(r0["w"], r1["w"])
    assert np.array_equal(r0["losses"], r1["losses"]) and np.all(np.isfinite(r0["losses"]))
    single = tmp_path / "single"
    single.mkdir()
    (s0,) = _spawn(_PPO_WORKER, single, 1, extra=(str(single),))
    assert int(s0["n_optsteps"]) == 3 * 2 * 8
    # same number of optimiser steps on the same amount of data: the weights moved by a comparable distance
    d_dp, d_s = np.abs(r0["w"]).mean(), np.abs(s0["w"]).mean()
    assert 0.5 < d_dp / d_s < 2.0
