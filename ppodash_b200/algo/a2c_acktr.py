"""A2C_ACKTR: drop-in for PKG/algo/a2c_acktr.py on the kernels of the PPO path.

Same constructor and ``update(rollouts) -> (value_loss, action_loss, dist_entropy)``.  One update = one forward / backward over the
whole rollout (T*N rows, time-major; recurrent policies unroll from ``recurrent_hidden_states[0]``), the A2C loss
(``ppd_a2c_loss_fwd_bwd``), global-norm clip and one RMSprop step (``ppd_clip_rmsprop_step``).  The reference's ``update`` still calls
``evaluate_actions`` with the 4-argument signature of the upstream repository (a2c_acktr.py:38-43) although this fork's Policy takes
the vector observations as a fifth argument; here they are passed, which is the only way the call can run.  ``acktr=True`` (K-FAC)
is out of scope and raises.
"""
import torch

from .. import _lib
from .. import dist as ppd_dist


class FusedClipRMSprop(torch.optim.Optimizer):
    """torch.optim.RMSprop-compatible facade (param_groups with lr / alpha / eps) over the fused clip + RMSprop kernel."""

    def __init__(self, policy, lr=None, eps=None, alpha=None):
        defaults = dict(lr=1e-2 if lr is None else lr, alpha=0.99 if alpha is None else alpha, eps=1e-8 if eps is None else eps)
        super().__init__(list(policy.parameters()), defaults)
        self._policy = policy
        self.max_grad_norm = None

    def zero_grad(self, set_to_none=False):
        eng = self._policy.engine()
        eng.bind()
        eng.flat_grad.zero_()

    @torch.no_grad()
    def step(self, closure=None, grad_norm_out=None):
        g = self.param_groups[0]
        self._policy.engine().rmsprop_step(g["lr"], g["alpha"], g["eps"], self.max_grad_norm, grad_norm_out)


class A2C_ACKTR():
    def __init__(self,
                 actor_critic,
                 value_loss_coef,
                 entropy_coef,
                 lr=None,
                 eps=None,
                 alpha=None,
                 max_grad_norm=None,
                 acktr=False,
                 process_group=None):
        if acktr:
            raise NotImplementedError("ACKTR (K-FAC) is out of scope (SURVEY.md section 2); use acktr=False")
        self.actor_critic = actor_critic
        self.acktr = acktr
        self.value_loss_coef = value_loss_coef
        self.entropy_coef = entropy_coef
        self.max_grad_norm = max_grad_norm
        self.optimizer = FusedClipRMSprop(actor_critic, lr, eps=eps, alpha=alpha)
        self.process_group = process_group

    def update(self, rollouts):
        pol = self.actor_critic
        eng = pol.engine()
        eng.bind()
        T, N = rollouts.rewards.size(0), rollouts.rewards.size(1)
        world = ppd_dist.world(self.process_group)[0]
        if getattr(rollouts, "obs_u8", False):
            # expand the uint8 frames of all T*N rows: the recurrent gather with the identity permutation is exactly that
            ident = torch.arange(N, dtype=torch.int64, device=rollouts.obs.device)
            obs = rollouts._gather("rec", ident, 0, T * N, N, None)[0]
        else:
            obs = rollouts.obs[:-1].reshape(T * N, *rollouts.obs.shape[2:])                  # a2c_acktr.py:39
        vobs = rollouts.vector_obs[:-1].reshape(T * N, -1)
        h0 = rollouts.recurrent_hidden_states[0].reshape(-1, pol.recurrent_hidden_state_size)    # :40-41
        masks = rollouts.masks[:-1].reshape(-1, 1)                                            # :42
        actions = rollouts.actions.reshape(-1, rollouts.actions.size(-1))                     # :43
        if not pol.is_recurrent:
            h0 = rollouts.recurrent_hidden_states[:-1].reshape(T * N, -1)
        ret = rollouts.returns[:-1].reshape(-1, 1)
        eng.train_minibatch((obs, vobs, h0, actions, None, ret, masks, None, None), 0.0, self.value_loss_coef, self.entropy_coef,
                            global_rows=T * N * world, loss="a2c")
        ppd_dist.all_reduce_sum(eng.flat_grad, self.process_group)
        losses = eng.flat_grad[eng.loss_off:eng.loss_off + 3].clone()
        self.optimizer.max_grad_norm = self.max_grad_norm
        self.optimizer.step()
        v = losses.tolist()          # the only device->host sync
        return v[0], v[1], v[2]
