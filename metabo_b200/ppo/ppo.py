"""PPO drop-in (metabo/ppo/ppo.py:43-264): same constructor, params keys, loss, update schedule,
log / checkpoint files; rollouts come from the GPU-resident BatchRecorder and minibatches stay on
the device.  Under torch.distributed (one process per GPU, NCCL over NVLink) the rollout lanes and
every minibatch are sharded across ranks and the policy update all-reduces

  * the flat fp32 gradient (sum; the loss of each rank is already divided by the GLOBAL minibatch
    size, so the reduced gradient equals the single-process gradient on the concatenated minibatch),
  * (sum, sum of squares, count) of the advantages, so the per-minibatch advantage normalisation of
    ppo.py:141-144 is the global one,

and every rank applies the identical Adam step (weights stay replicated, no broadcast).

The optimiser step runs entirely on the library's kernels, for every network NeuralAF can build: the shipped
F -> H -> H -> H -> H -> 1 ReLU policies on the tcgen05 kernels behind `mbo_ppo_policy_grad` (csrc/ppo_update.cu), every
other shape / tanh (e.g. train_metabo_gps.py's 4 -> 20 -> 20 -> 1) on the fp32 CUDA-core kernels of
`mbo_ppo_policy_grad_simt` (csrc/ppo_update_simt.cu), value nets likewise (`mbo_ppo_value_grad[_simt]`), Adam with
`mbo_adam_step` / `mbo_comm_allreduce_adam`.  There is no torch-autograd update path in the product (tests keep one as
the comparison implementation, tests/autograd_ppo.py).
"""
import collections
import copy
import json
import os
import pickle as pkl
import random
import time
from datetime import datetime

import numpy as np
import torch
import torch.distributed as dist
import torch.optim

from .. import gym_compat as gym
from .. import ops
from .batchrecorder import BatchRecorder


def dist_ctx():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def normalize_advantages(advs, world=1):
    """(advs - mean) / std with the biased std over the GLOBAL minibatch (ppo.py:141-144)."""
    cnt = torch.full((), float(advs.numel()), device=advs.device, dtype=advs.dtype)    # fill kernel, no H2D copy
    s = torch.stack([advs.sum(), (advs * advs).sum(), cnt]).double()
    if world > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    n = s[2]
    mean = s[0] / n
    var = torch.clamp(s[1] / n - mean * mean, min=0.0)
    std = torch.sqrt(var)
    ok = (std > 0) & ~torch.isnan(std)            # ppo.py:143: leave the advantages alone if std is 0 / NaN
    safe = torch.where(ok, std, torch.ones_like(std))
    normed = ((advs.double() - mean) / safe).to(advs.dtype)
    return torch.where(ok, normed, advs)          # branch-free: capturable in a CUDA graph


def ppo_minibatch_loss(pi, states, actions, advs, tdlamrets, logprobs_old, params, n_global):
    """Loss terms of one (local shard of a) minibatch, each already divided by the global minibatch
    size n_global so that summing gradients over ranks gives the reference's mean-reduced loss
    (ppo.py:148-168)."""
    vpreds, logprobs, entropies = pi.predict_vals_logps_ents(states=states, actions=actions)
    assert logprobs_old.dim() == vpreds.dim() == logprobs.dim() == entropies.dim() == 1
    ratios = torch.exp(logprobs - logprobs_old)
    clipped = ratios.clamp(1 - params["epsilon"], 1 + params["epsilon"])
    loss_ppo = -torch.sum(torch.min(ratios * advs, clipped * advs)) / n_global
    loss_value = torch.sum((vpreds - tdlamrets) ** 2) / n_global
    loss_ent = -torch.sum(entropies) / n_global
    loss = loss_ppo + params["value_coeff"] * loss_value + params["ent_coeff"] * loss_ent
    return loss, loss_ppo, loss_value, loss_ent


def minibatch_bounds(n, minibatch_size):
    """batchrecorder.py:326-330: consecutive minibatches, the remainder joins the last one"""
    pos, out = 0, []
    while pos < n:
        cur = n - pos if pos + 2 * minibatch_size > n else minibatch_size
        out.append((pos, pos + cur))
        pos += cur
    return out


def shard_bounds(n_local, minibatch_size, rank, world):
    """Data-parallel minibatches: every optimiser step works on a GLOBAL minibatch of the reference's size, of which
    rank r holds the share [c_r(a), c_r(b)) of its own (locally shuffled) n_local transitions, c_r(x) = (x + r) // world.
    By Hermite's identity the shares of all ranks add up to b - a for every step, and to n_local per rank over the
    epoch, for ANY world size (60 = 8+7+8+7+... on 8 ranks).  Returns (local bounds, global minibatch sizes)."""
    gb = minibatch_bounds(n_local * world, minibatch_size)
    c = lambda x: (x + rank) // world
    return [(c(a), c(b)) for a, b in gb], [b - a for a, b in gb]


def allreduce_gradients(parameters, world):
    """one flat fp32 all-reduce (sum) per optimiser step (~0.97 MB for the 4x200 nets)"""
    if world == 1:
        return
    grads = [p.grad for p in parameters if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


class PPO:
    def __init__(self, policy_fn, params, logpath, save_interval, verbose=False):
        self.params = params
        self.rank, self.world = dist_ctx()
        self.env_spec = gym.registry.env_specs[self.params["env_id"]]
        self.set_all_seeds()

        self.logpath = logpath
        if self.rank == 0:
            try:
                os.makedirs(logpath, exist_ok=False)
            except FileExistsError:
                raise ValueError("Logpath is not empty!")
        self.save_interval = save_interval
        self.verbose = verbose

        if not torch.cuda.is_available():
            raise RuntimeError("metabo_b200.ppo needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device())

        # rollout lanes: workers are sharded contiguously over ranks, seeds keep their meaning
        nw, bs = self.params["n_workers"], self.params["batch_size"]
        if nw % self.world or bs % self.world:
            raise ValueError("n_workers and batch_size must be divisible by the number of ranks")
        if self.params["minibatch_size"] < self.world:
            raise ValueError("minibatch_size must be at least the number of ranks")
        lw = nw // self.world
        seeds = list(self.params["env_seeds"])[self.rank * lw:(self.rank + 1) * lw]
        self.batch_recorder = BatchRecorder(size=bs // self.world, env_id=self.params["env_id"], env_seeds=seeds,
                                            policy_fn=policy_fn, n_workers=lw, device=self.device,
                                            mlp_mode=self.params.get("mlp_mode"))
        osp, asp = self.batch_recorder.observation_space, self.batch_recorder.action_space
        self.pi = policy_fn(osp, asp, False).to(self.device)
        self.old_pi = policy_fn(osp, asp, False).to(self.device)
        if self.world > 1:   # identical initial weights on every rank
            for p in self.pi.parameters():
                dist.broadcast(p.data, src=0)
        self.use_update_graph = os.environ.get("MBO_NO_GRAPHS") is None and self.params.get("update_graph", True)
        self.comm = None        # ops.Comm (peer-memory gradient exchange), created with the gradient buffers when world > 1
        self._upd = None        # (static inputs, CUDAGraph, static outputs) of one captured optimiser step
        self._ep = None         # native backend: one captured EPOCH (all minibatch steps) over static copies of the batch
        self._ep_stale = True
        backend = self.params.get("update_backend", "native")
        if backend != "native":
            raise ValueError("update_backend must be 'native': the optimiser step runs on the library's kernels only")
        self.update_backend = backend
        if self.world > 1 and not self.dp_native() and self.params["minibatch_size"] % self.world:
            raise ValueError("minibatch_size must be divisible by the number of ranks (NCCL exchange path)")
        self.optimizer = self._make_optimizer()
        self._pg = None         # ops.PPOPolicyGrad for the current minibatch shape
        self._pg_grads = None   # persistent gradient buffers of the policy net (written by the kernels)
        self._vg = None         # ops.PPOValueGrad
        self._vg_grads = None
        self._flat_grads = None
        self._loss4 = None
        self._side = None       # second stream: the value-net branch of an optimiser step
        self._mbufs = {}        # minibatch buffers of mbo_ppo_prepare_minibatch, by minibatch size

        self.stats = dict()
        self.stats["n_timesteps"] = 0
        self.stats["n_optsteps"] = 0
        self.stats["n_iters"] = 0
        self.stats["t_train"] = 0
        self.stats["avg_step_rews"] = np.array([])
        self.stats["avg_init_rews"] = np.array([])
        self.stats["avg_term_rews"] = np.array([])
        self.stats["avg_ep_rews"] = np.array([])
        self.t_batch = None
        self.rew_buffer = collections.deque(maxlen=50)
        if self.rank == 0:
            self.write_overview_logfile()

    def set_all_seeds(self):
        np.random.seed(self.params["seed"])
        random.seed(self.params["seed"])
        torch.manual_seed(self.params["seed"])
        torch.cuda.manual_seed_all(self.params["seed"])

    def write_overview_logfile(self):
        s = ""
        s += "********* OVERVIEW OF RUN *********\n"
        s += "Logpath        : {}\n".format(self.logpath)
        s += "Logfile created: {}\n".format(datetime.strftime(datetime.now(), "%Y-%m-%d-%H-%M-%S"))
        s += "Environment-ID:  {}\n".format(self.params["env_id"])
        s += "Environment-kwargs:\n"
        s += json.dumps(self.env_spec._kwargs, indent=2)
        s += "\n"
        s += "PPO-parameters:\n"
        s += json.dumps(self.params, indent=2)
        s += "\n"
        s += "Batchrecorder:\n"
        s += json.dumps(self.batch_recorder.overview_dict(), indent=2)
        with open(os.path.join(self.logpath, "000_overview.txt"), "w") as f:
            print(s, file=f)
        if not self.verbose:
            print(s)

    def _make_optimizer(self):
        """torch.optim.Adam(lr) semantics (ppo.py:93), one launch per step (mbo_adam_step)"""
        return ops.FusedAdam(self.pi.parameters(), lr=self.params["lr"])

    @staticmethod
    def policy_widths(pi):
        return [pi.N_features_policy] + list(pi.policy_net.arch_spec) + [1]

    @staticmethod
    def value_widths(pi):
        return [2] + list(pi.value_net.arch_spec) + [1]

    def _alloc_native_grads(self):
        """Persistent gradient buffers of the native backend, all views of ONE flat fp32 tensor: the kernels write into
        them, and under torch.distributed the gradient all-reduce is a single call on the flat buffer."""
        pi = self.pi
        layers = list(pi.policy_net.fc) + (list(pi.value_net.fc) if pi.use_value_network else [])
        n = sum(l.weight.numel() + l.bias.numel() for l in layers)
        self._flat_grads = torch.zeros(n, dtype=torch.float32, device=self.device)
        off, views = 0, []
        for l in layers:
            w = self._flat_grads[off:off + l.weight.numel()].view_as(l.weight); off += l.weight.numel()
            b = self._flat_grads[off:off + l.bias.numel()].view_as(l.bias); off += l.bias.numel()
            views.append((w, b))
        k = len(pi.policy_net.fc)
        self._pg_grads = ([w for w, _ in views[:k]], [b for _, b in views[:k]])
        self._vg_grads = ([w for w, _ in views[k:]], [b for _, b in views[k:]])
        if self.world > 1 and self.dp_native():
            self._ensure_comm(4 * n)

    def _ensure_comm(self, nbytes):
        """The ONE collective of the path: the flat gradient is pulled from every rank's slot over NVLink inside the Adam
        kernel (mbo_comm_allreduce_adam).  Creating the exchange object is itself collective (every rank gets here at the
        same point of its first optimiser step / epoch); the slot also carries the per-epoch advantage moments."""
        need = max(int(nbytes), 4096)
        if self.comm is None or self.comm.slot_bytes < need:
            self.comm = ops.Comm.from_process_group(need, self.device)
            self._ep = None                    # captured graphs hold the old regions' pointers
        if self._flat_grads is not None:
            self.optimizer.set_comm(self.comm, self._flat_grads)

    def dp_native(self):
        """gradient exchange inside the Adam kernel over NVLink peer memory (default) instead of an eager NCCL call"""
        return self.params.get("dp_exchange", "peer") == "peer"

    def _value_grad_op(self, E):
        if self._vg is None or self._vg.E < E:
            assert not torch.cuda.is_current_stream_capturing(), "value-net workspace must be sized before a capture"
            pi = self.pi
            self._vg = ops.make_value_grad(E, self.value_widths(pi), pi.activation, self.device)
            self._upd = self._ep = None      # captured graphs hold the old workspace's pointers
        return self._vg

    def _policy_grad_op(self, E, M, Fs):
        pg = self._pg
        if pg is None or pg.E < E or (pg.M, pg.F_state) != (M, Fs):
            pi = self.pi
            self._pg = ops.make_policy_grad(E, M, Fs, pi.policy_columns(), self.policy_widths(pi), pi.activation, self.device)
            self._upd = self._ep = None      # captured graphs hold the old workspace's pointers
        return self._pg

    def _side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def _old_logprobs(self, states, actions):
        """ppo.py:146-147 `logprobs_old` of the frozen old policy, by the update kernels' own forward (same
        arithmetic as the new policy's log-probs: the ratio is exactly 1 while theta == theta_old)."""
        fcs = list(self.old_pi.policy_net.fc)
        pg = self._policy_grad_op(*states.shape)
        lp, _ = pg.logp_entropy(states.contiguous(), [l.weight.data for l in fcs], [l.bias.data for l in fcs],
                                actions.to(torch.int32).contiguous())
        return lp

    def _update_step_native(self, mb, prepared=False, n_global=None):
        """ppo.py:129-178 on the hand-written kernels: policy net (mbo_ppo_policy_grad), value net (mbo_ppo_value_grad),
        Adam (mbo_adam_step).  prepared=True: `mb` comes from mbo_ppo_prepare_minibatch (gathered, advantages already
        normalised, int32 actions) and the logged loss sums are one more kernel -- no torch op is left in the step."""
        states, actions, tdlamrets = mb["states"], mb["actions"], mb["tdlamrets"]
        advs = mb["advs"] if (prepared or self.params["loss_type"] == "GAElam") else mb["tdlamrets"]
        E, M, Fs = states.shape
        n_global = E * self.world if n_global is None else n_global
        if self.params["normalize_advs"] and not prepared:
            advs = normalize_advantages(advs, self.world)
        pi = self.pi
        fcs = list(pi.policy_net.fc)
        self._policy_grad_op(E, M, Fs)
        logprobs_old = mb.get("logprobs_old")
        if logprobs_old is None:
            logprobs_old = self._old_logprobs(states, actions)
        if self._pg_grads is None:
            self._alloc_native_grads()
        dW, db = self._pg_grads
        self.optimizer.zero_grad(set_to_none=True)
        # The value net's ~15 small kernels (60 rows) depend only on the minibatch: they run on a second stream beside
        # the policy-net kernels and join before Adam (inside a captured graph this is a fork / join of two branches).
        side = None
        if pi.use_value_network:
            vfc = list(pi.value_net.fc)
            self._value_grad_op(E)
            vW, vb = self._vg_grads
            cur = torch.cuda.current_stream()
            side = self._side_stream() if self.params.get("value_branch", True) else None
            if side is not None:
                side.wait_stream(cur)
            with torch.cuda.stream(side if side is not None else cur):
                _, vterms = self._vg(states.contiguous(), pi.t_idx, pi.T_idx, [l.weight.data for l in vfc],
                                     [l.bias.data for l in vfc], tdlamrets.to(torch.float32).contiguous(),
                                     self.params["value_coeff"] / n_global, vW, vb)
            for l, gw, gb in zip(vfc, vW, vb):
                l.weight.grad, l.bias.grad = gw, gb
        else:
            vterms = tdlamrets.to(torch.float32) ** 2          # the reference's value head is then the constant 0
        terms = self._pg(states.contiguous(), [l.weight.data for l in fcs], [l.bias.data for l in fcs],
                         actions.to(torch.int32).contiguous(), advs.to(torch.float32).contiguous(),
                         logprobs_old.to(torch.float32).contiguous(), self.params["epsilon"], self.params["ent_coeff"],
                         n_global, dW, db)
        for l, gw, gb in zip(fcs, dW, db):
            l.weight.grad, l.bias.grad = gw, gb
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        if prepared and (self.world == 1 or self.comm is not None):
            self.optimizer.step()              # world > 1: gradient exchange + Adam in one kernel (FusedAdam.set_comm)
            if self._loss4 is None:
                self._loss4 = torch.zeros(4, device=self.device)
            return ops.ppo_loss_sums(terms, vterms.contiguous(), n_global, self.params["value_coeff"], self.params["ent_coeff"],
                                     self._loss4)
        l_val = vterms.sum() / n_global
        if self.comm is not None:
            pass                                   # the exchange happens inside optimizer.step()
        elif self.world > 1:
            dist.all_reduce(self._flat_grads, op=dist.ReduceOp.SUM)      # every gradient is a view of this one buffer
        self.optimizer.step()
        l_ppo = terms[:, 2].sum()
        l_ent = -terms[:, 1].sum() / n_global
        loss = l_ppo + self.params["value_coeff"] * l_val + self.params["ent_coeff"] * l_ent
        return torch.stack([l_ppo, l_val, l_ent, loss]).detach()

    def _update_step(self, mb):
        """one optimiser step on the (local shard of a) minibatch `mb` (dict of device tensors), ppo.py:129-178"""
        return self._update_step_native(mb)

    def _graphed_update(self, mb):
        """One optimiser step (~35 small kernels) captured once in a CUDA graph (static input buffers) and replayed per
        minibatch (`epoch_graph=False`; the default captures a whole epoch, `_graphed_epoch`)."""
        n = mb["states"].shape[0]
        if self._upd is None or self._upd[0]["states"].shape != mb["states"].shape:
            static = {k: torch.empty_like(v) for k, v in mb.items() if v is not None}
            for k in static:
                static[k].copy_(mb[k])
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            snap = [p.detach().clone() for p in self.pi.parameters()]
            opt_snap = [{k: v.clone() for k, v in st.items() if torch.is_tensor(v)} for st in self.optimizer.state.values()]
            with torch.cuda.stream(side):
                for _ in range(3):                                   # warm-up on a side stream (torch graph recipe)
                    self._update_step(static)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._update_step(static)
            # the warm-up steps were real optimiser steps on this minibatch: undo them so that the numbers of
            # optimiser steps and the parameter trajectory are exactly those of the eager schedule
            with torch.no_grad():
                for p, q in zip(self.pi.parameters(), snap):
                    p.copy_(q)
                for st, sn in zip(self.optimizer.state.values(), opt_snap):
                    for k, v in sn.items():
                        st[k].copy_(v)
            self._upd = (static, g, out, self.stats["n_optsteps"])
        static, g, out, _ = self._upd
        for k in static:
            static[k].copy_(mb[k])
        g.replay()
        return out.clone()

    _MB_KEYS = ("states", "actions", "tdlamrets", "advs", "logprobs_old")

    def _minibatch_buffers(self, E, states):
        if E not in self._mbufs:
            f = lambda dt: torch.zeros(E, dtype=dt, device=self.device)
            self._mbufs[E] = dict(states=torch.zeros((E,) + tuple(states.shape[1:]), dtype=torch.float32, device=self.device),
                                  actions=f(torch.int32), advs=f(torch.float32), tdlamrets=f(torch.float32),
                                  logprobs_old=f(torch.float32))
        return self._mbufs[E]

    def _graphed_epoch(self, order, local_mb):
        """All optimiser steps of one epoch (ppo.py:127-178) as ONE captured CUDA graph (native backend): the batch lives
        in static device buffers, the shuffled order in a static index tensor, every minibatch is gathered by a
        captured index_select at its fixed offset; a replay reads the new order / batch.  Removes ~0.15 ms of host
        launch overhead per optimiser step (80 graph launches per iteration -> 4)."""
        br = self.batch_recorder
        buf = br.buf
        n = br.B * br.T
        flat = {k: buf[k].reshape((n,) + tuple(buf[k].shape[2:])) for k in self._MB_KEYS}
        sig = tuple((k, tuple(v.shape), v.dtype) for k, v in flat.items())
        if self._ep is None or self._ep["sig"] != sig:
            static = {k: v.clone() for k, v in flat.items()}
            perm = torch.zeros(n, dtype=torch.int64, device=self.device)
            bounds = br._minibatch_bounds(local_mb)
            E_max = max(b - a for a, b in bounds)
            self._policy_grad_op(E_max, *flat["states"].shape[1:])          # size the workspaces before the capture
            if self.pi.use_value_network:
                self._value_grad_op(E_max)
            src = dict(static)
            if self.params["loss_type"] != "GAElam":
                src["advs"] = static["tdlamrets"]
            fused = (self.world == 1 and all(v.is_contiguous() for v in static.values()) and static["actions"].dtype == torch.int32
                     and self.params.get("fused_minibatch", True))
            for a, b in bounds:
                self._minibatch_buffers(b - a, static["states"])
            if self._loss4 is None:
                self._loss4 = torch.zeros(4, device=self.device)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                sums = torch.zeros(4, device=self.device)
                for a, b in bounds:
                    sel = perm[a:b]
                    if fused:       # gather + advantage normalisation in one kernel, loss sums in another
                        mbuf = ops.ppo_prepare_minibatch(sel, src, self.params["normalize_advs"],
                                                         self._minibatch_buffers(b - a, static["states"]))
                        sums = sums + self._update_step_native(mbuf, prepared=True)
                    else:
                        sums = sums + self._update_step({k: v.index_select(0, sel) for k, v in static.items()})
            self._ep = dict(sig=sig, static=static, perm=perm, graph=g, sums=sums, bounds=bounds)
            self._ep_stale = True
        ep = self._ep
        if self._ep_stale:                                  # first epoch of this optimize_on_batch: refresh the static copies
            for k, v in flat.items():
                ep["static"][k].copy_(v)
            self._ep_stale = False
        ep["perm"].copy_(torch.as_tensor(order, dtype=torch.int64), non_blocking=False)
        ep["graph"].replay()
        return ep["sums"].clone(), len(ep["bounds"])

    def _dp_epoch(self, order):
        """One epoch of optimiser steps under data parallelism (world > 1, native kernels): every rank shuffles ITS
        transitions (`order`), step j works on the global minibatch whose local share is perm[a_j:b_j] (shard_bounds),
        the advantage moments of all steps are exchanged once per epoch, each step is
            mbo_ppo_prepare_minibatch -> mbo_ppo_policy_grad -> mbo_ppo_value_grad -> mbo_comm_allreduce_adam -> mbo_ppo_loss_sums
        and the logged loss sums are exchanged at the end -- kernels only, no NCCL call, so the whole epoch is ONE
        captured CUDA graph per rank (eager on the first call, which allocates the optimiser state).  Every rank
        enqueues the same sequence of exchange calls; the peer waits inside the kernels are bounded."""
        br = self.batch_recorder
        buf = br.buf
        n = br.B * br.T
        flat = {k: buf[k].reshape((n,) + tuple(buf[k].shape[2:])) for k in self._MB_KEYS}
        sig = tuple((k, tuple(v.shape), v.dtype) for k, v in flat.items())
        if self._ep is None or self._ep["sig"] != sig:
            static = {k: v.clone().contiguous() for k, v in flat.items()}
            assert static["actions"].dtype == torch.int32
            perm = torch.zeros(n, dtype=torch.int64, device=self.device)
            bounds, n_globals = shard_bounds(n, self.params["minibatch_size"], self.rank, self.world)
            E_max = max(b - a for a, b in bounds)
            self._policy_grad_op(E_max, *flat["states"].shape[1:])          # size the workspaces before any capture
            if self.pi.use_value_network:
                self._value_grad_op(E_max)
            if self._pg_grads is None:
                self._alloc_native_grads()                                   # collective: creates the comm
            self._ensure_comm(max(4 * self._flat_grads.numel(), 24 * len(bounds)))    # the slot also carries [n_steps, 3] fp64 moments
            for a, b in bounds:
                self._minibatch_buffers(b - a, static["states"])
            if self._loss4 is None:
                self._loss4 = torch.zeros(4, device=self.device)
            src = dict(static)
            if self.params["loss_type"] != "GAElam":
                src["advs"] = static["tdlamrets"]
            bounds_dev = torch.tensor([bounds[0][0]] + [b for _, b in bounds], dtype=torch.int32, device=self.device)
            stats = torch.zeros((len(bounds), 3), dtype=torch.float64, device=self.device)
            self._ep = dict(sig=sig, static=static, perm=perm, graph=None, sums=None, bounds=bounds, n_globals=n_globals,
                            src=src, bounds_dev=bounds_dev, stats=stats, runs=0)
            self._ep_stale = True
        ep = self._ep
        if self._ep_stale:
            for k, v in flat.items():
                ep["static"][k].copy_(v)
            self._ep_stale = False
        ep["perm"].copy_(torch.as_tensor(order, dtype=torch.int64), non_blocking=False)
        norm = bool(self.params["normalize_advs"])

        def body():
            sums = torch.zeros(4, device=self.device)
            if norm:        # global per-minibatch advantage moments (ppo.py:141-144), one exchange per epoch
                ops.ppo_adv_stats(ep["bounds_dev"], ep["perm"], ep["src"]["advs"], ep["stats"])
                self.comm.allreduce(ep["stats"])
            for j, (a, b) in enumerate(ep["bounds"]):
                mbuf = ops.ppo_prepare_minibatch(ep["perm"][a:b], ep["src"], norm, self._minibatch_buffers(b - a, ep["static"]["states"]),
                                                 adv_stats=ep["stats"][j] if norm else None)
                sums = sums + self._update_step_native(mbuf, prepared=True, n_global=ep["n_globals"][j])
            self.comm.allreduce(sums)
            return sums

        use_graph = self.use_update_graph and self.params.get("epoch_graph", True)
        if not use_graph or ep["runs"] == 0:
            sums = body()
        else:
            if ep["graph"] is None:
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    ep["sums"] = body()
                ep["graph"] = g
            ep["graph"].replay()
            sums = ep["sums"]
        ep["runs"] += 1
        return sums.clone(), len(ep["bounds"])

    def _batch_old_logprobs(self, buf, chunk):
        """theta_old is frozen for the whole optimize_on_batch call: its log-probs (recomputed per minibatch in
        ppo.py:146-147) are computed once per transition here and travel with the minibatches"""
        B, T = buf["actions"].shape[:2]
        st = buf["states"].reshape((B * T,) + tuple(buf["states"].shape[2:]))
        ac = buf["actions"].reshape(B * T)
        with torch.no_grad():
            lp = [self._old_logprobs(st[i:i + chunk], ac[i:i + chunk]) for i in range(0, B * T, chunk)]
        buf["logprobs_old"] = torch.cat(lp).reshape(B, T)

    def optimize_on_batch(self):
        now = time.time()
        self.iter_log_str += " OPTIMIZING...\n"
        self.old_pi.load_state_dict(self.pi.state_dict())
        self.old_pi.set_requires_grad(False)
        self.old_pi.kernel_handles()      # re-pack the frozen policy's kernel weights NOW: a replayed graph cannot do it
        local_mb = -(-self.params["minibatch_size"] // self.world)       # largest local share of a minibatch
        dp = self.world > 1 and self.dp_native()
        buf = self.batch_recorder.buf
        buf.pop("logprobs_old", None)
        self._ep_stale = True             # the epoch graph reads static copies of the batch: refresh them once per call
        self._batch_old_logprobs(buf, local_mb)
        for ep in range(self.params["n_epochs"]):
            sums = torch.zeros(4, device=self.device)
            n_minibatches = 0
            if dp:
                sums, n_minibatches = self._dp_epoch(self.batch_recorder._order(True))       # sums are already global
                self.stats["n_optsteps"] += n_minibatches
            elif (self.use_update_graph and self.world == 1 and self.stats["n_optsteps"] > 0
                    and self.params.get("epoch_graph", True)):
                sums, n_minibatches = self._graphed_epoch(self.batch_recorder._order(True), local_mb)
                self.stats["n_optsteps"] += n_minibatches
            for mb in (() if n_minibatches else self.batch_recorder.iterate_device(local_mb, shuffle=True)):
                mb = {k: v for k, v in mb.items() if k in ("states", "actions", "tdlamrets", "advs", "logprobs_old")}
                graphable = (self.use_update_graph and self.world == 1 and mb["states"].shape[0] == local_mb
                             and self.stats["n_optsteps"] > 0)
                sums += self._graphed_update(mb) if graphable else self._update_step(mb)
                self.stats["n_optsteps"] += 1
                n_minibatches += 1
            if self.world > 1 and not dp:
                dist.all_reduce(sums, op=dist.ReduceOp.SUM)
            avg = (sums / n_minibatches).tolist()
            self.iter_log_str += "   loss_ppo = {: .4g}, loss_value = {: .4g}, loss_ent = {: .4g}, loss = {: .4g}\n".format(*avg)
        torch.cuda.synchronize(self.device)
        if self._pg is not None:
            self._pg.check()              # sticky error word of the update kernels (bounded pipeline waits)
        if self.comm is not None:
            self.comm.check()             # a peer wait of the gradient exchange timed out
        t_optim = time.time() - now
        self.iter_log_str += "  Took {:.2f}s".format(t_optim)
        return t_optim

    def _global_batch_stats(self, bs):
        """sum the per-rank batch statistics (each rank recorded batch_size / world steps)"""
        if self.world == 1:
            return bs
        br = self.batch_recorder
        v = torch.tensor([br.reward_sum, float(len(br)), float(br.n_new), float(np.sum(br.initial_rewards)),
                          float(np.sum(br.terminal_rewards))], dtype=torch.float64, device=self.device)
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
        rs, n, nn, si, st = v.tolist()
        bs = dict(bs)
        bs.update(size=int(n), avg_step_reward=rs / n, avg_initial_reward=si / nn, avg_terminal_reward=st / nn,
                  avg_ep_reward=rs / nn, avg_ep_len=n / nn, n_new=int(nn))
        return bs

    def train(self):
        while self.stats["n_timesteps"] < self.params["max_steps"]:
            self.iter_log_str = ""
            t_weights = self.batch_recorder.set_worker_weights(copy.deepcopy(self.pi))
            self.stats["t_train"] += t_weights
            t_batch = self.batch_recorder.record_batch(gamma=self.params["gamma"], lam=self.params["lambda"])
            self.stats["t_train"] += t_batch
            batch_stats = self._global_batch_stats(self.batch_recorder.get_batch_stats())
            self.stats["n_timesteps"] += batch_stats["size"]
            self.stats["batch_stats"] = batch_stats
            self.stats["avg_step_rews"] = np.append(self.stats["avg_step_rews"], batch_stats["avg_step_reward"])
            self.stats["avg_init_rews"] = np.append(self.stats["avg_init_rews"], batch_stats["avg_initial_reward"])
            self.stats["avg_term_rews"] = np.append(self.stats["avg_term_rews"], batch_stats["avg_terminal_reward"])
            self.stats["avg_ep_rews"] = np.append(self.stats["avg_ep_rews"], batch_stats["avg_ep_reward"])
            self.stats["perc"] = 100 * self.stats["n_timesteps"] / self.params["max_steps"]
            batch_stats["t_batch"] = t_batch
            batch_stats["sps"] = batch_stats["size"] / batch_stats["t_batch"]
            self.add_iter_log_str(batch_stats=batch_stats)
            if self.stats["n_iters"] % self.save_interval == 0 or self.stats["n_iters"] == 0 or \
                    self.stats["n_timesteps"] >= self.params["max_steps"]:
                self.store_weights()
            t_optim = self.optimize_on_batch()
            self.stats["t_train"] += t_optim
            self.stats["t_optim"] = t_optim
            self.store_log()
            self.stats["n_iters"] += 1
        self.batch_recorder.cleanup()

    def store_log(self):
        if self.rank != 0:
            return
        with open(os.path.join(self.logpath, "log"), "a") as f:
            print(self.iter_log_str, file=f)
        if not self.verbose:
            print(self.iter_log_str)
        with open(os.path.join(self.logpath, "stats_" + str(self.stats["n_iters"])), "wb") as f:
            pkl.dump(self.stats, f)
        with open(os.path.join(self.logpath, "params_" + str(self.stats["n_iters"])), "wb") as f:
            pkl.dump(self.params, f)

    def store_weights(self):
        if self.rank != 0:
            return
        with open(os.path.join(self.logpath, "weights_" + str(self.stats["n_iters"])), "wb") as f:
            torch.save({k: v.cpu() for k, v in self.pi.state_dict().items()}, f)

    def add_iter_log_str(self, batch_stats):
        s = "\n******************** ITERATION {:2d} ********************\n".format(self.stats["n_iters"])
        s += " RUN STATISTICS (BEFORE OPTIMIZATION):\n"
        s += "   environment          = {}\n".format(self.params["env_id"])
        s += "   n_timesteps  = {:d} ({:.2f}%)\n".format(self.stats["n_timesteps"], self.stats["perc"])
        s += "   n_optsteps   = {:d}\n".format(self.stats["n_optsteps"])
        s += "   t_total      = {:.2f}s\n".format(self.stats["t_train"])
        s += " BATCH STATISTICS (BEFORE OPTIMIZATION):\n"
        s += "   n_workers    = {:d}\n".format(self.params["n_workers"])
        s += "   worker_seeds = {}\n".format(self.params["env_seeds"])
        s += "   size         = {:d}\n".format(batch_stats["size"])
        s += "    per_worker  = {:}\n".format(batch_stats["worker_sizes"])
        s += "   avg_step_rew = {:.4g}\n".format(batch_stats["avg_step_reward"])
        s += "   avg_init_rew = {:.4g}\n".format(batch_stats["avg_initial_reward"])
        s += "   avg_term_rew = {:.4g}\n".format(batch_stats["avg_terminal_reward"])
        s += "   avg_ep_rew   = {:.4g}\n".format(batch_stats["avg_ep_reward"])
        s += "   n_new        = {:d}\n".format(batch_stats["n_new"])
        s += "    per_worker  = {:}\n".format(batch_stats["worker_n_news"])
        s += "   avg_ep_len   = {:.2f}\n".format(batch_stats["avg_ep_len"])
        s += "   t_batch      = {:.2f}s ({:.0f}sps)\n".format(batch_stats["t_batch"], batch_stats["sps"])
        self.iter_log_str += s
