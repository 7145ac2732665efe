"""Test infrastructure: the PPO optimiser step of ppo.py:129-178 with torch autograd + torch.optim.Adam (all-torch fp32),
as a subclass of the drop-in PPO.  The product has no autograd update path; this class is the comparison implementation
the GPU tests run the native kernels against (same rollouts, same minibatch order)."""
import torch

from metabo_b200.ppo.ppo import PPO, allreduce_gradients, normalize_advantages, ppo_minibatch_loss


class AutogradPPO(PPO):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.use_update_graph = False
        torch.set_float32_matmul_precision("highest")

    def _make_optimizer(self):
        return torch.optim.Adam(self.pi.parameters(), lr=self.params["lr"])

    def _batch_old_logprobs(self, buf, chunk):
        pass                              # recomputed per minibatch below, as the reference does (ppo.py:146-147)

    def _update_step(self, mb):
        states, actions, tdlamrets = mb["states"], mb["actions"], mb["tdlamrets"]
        advs = mb["advs"] if self.params["loss_type"] == "GAElam" else mb["tdlamrets"]
        n_global = states.shape[0] * self.world
        if self.params["normalize_advs"]:
            advs = normalize_advantages(advs, self.world)
        with torch.no_grad():
            _, logprobs_old, _ = self.old_pi.predict_vals_logps_ents(states=states, actions=actions)
        self.pi.set_requires_grad(True)
        loss, l_ppo, l_val, l_ent = ppo_minibatch_loss(self.pi, states, actions, advs, tdlamrets, logprobs_old,
                                                       self.params, n_global)
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        allreduce_gradients(list(self.pi.parameters()), self.world)
        self.optimizer.step()
        return torch.stack([l_ppo, l_val, l_ent, loss]).detach()
