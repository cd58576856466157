"""Oracle (test infrastructure): one ``PPO.update(rollouts)`` on torch-CPU.

Restates PKG/algo/ppo.py:34-96.  Uses torch autograd, ``clip_grad_norm_`` and
``torch.optim.Adam`` exactly as the reference does (torch is the third-party
dependency that *is* present on both machines).
"""
import torch

from . import minibatch, policy


def ppo_losses(values, logp, entropy, old_values, returns, old_logp, adv, clip, use_clipped_value_loss=True):
    """ppo.py:61-77 -> (value_loss, action_loss)."""
    ratio = torch.exp(logp - old_logp)
    s1 = ratio * adv
    s2 = torch.clamp(ratio, 1.0 - clip, 1.0 + clip) * adv
    action_loss = -torch.min(s1, s2).mean()
    if use_clipped_value_loss:
        vclip = old_values + (values - old_values).clamp(-clip, clip)
        e1 = (values - returns).pow(2)
        e2 = (vclip - returns).pow(2)
        value_loss = 0.5 * torch.max(e1, e2).mean()
    else:
        value_loss = 0.5 * (returns - values).pow(2).mean()
    return value_loss, action_loss


class UpdateState:
    """Parameters (leaf tensors keyed by state_dict name) + their Adam optimiser."""

    def __init__(self, params, lr, eps, flat_gru=False):
        self.params = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
        gru = ["base.gru.weight_ih_l0", "base.gru.weight_hh_l0", "base.gru.bias_ih_l0", "base.gru.bias_hh_l0"]
        if flat_gru and all(k in self.params for k in gru):
            # nn.GRU keeps its four tensors in ONE buffer (flatten_parameters) so that cuDNN need not compact them on every call;
            # the stock-PyTorch-CUDA comparison row of bench.py gets the same layout: leaf tensors that are views of one buffer
            flat = torch.cat([self.params[k].detach().reshape(-1) for k in gru])
            off = 0
            for k in gru:
                n = self.params[k].numel()
                self.params[k] = flat[off:off + n].view(self.params[k].shape).detach().requires_grad_(True)
                off += n
            self._flat_gru = flat
        self.optimizer = torch.optim.Adam(list(self.params.values()), lr=lr, eps=eps)


def ppo_update(state, roll, *, recurrent, clip_param, ppo_epoch, num_mini_batch, value_loss_coef,
               entropy_coef, max_grad_norm, use_clipped_value_loss=True, concat_vector=True,
               on_minibatch=None, max_minibatches=None):
    """Run all epochs x minibatches; returns (value_loss, action_loss, dist_entropy) means.

    ``roll`` holds torch CPU tensors under the reference attribute names.
    ``on_minibatch(k, info)`` (optional) receives per-minibatch losses and grads
    before the optimiser step -- used by the parity tests.  ``max_minibatches`` stops early after
    that many minibatches (bench.py times a bounded sample of the CPU path) and returns the means
    over the minibatches actually run.
    """
    p = state.params
    adv = roll["returns"][:-1] - roll["value_preds"][:-1]
    adv = (adv - adv.mean()) / (adv.std() + 1e-5)
    tot_v = tot_a = tot_e = 0.0
    k = 0
    for _ in range(ppo_epoch):
        if recurrent:
            gen = minibatch.recurrent_minibatches(roll, adv, num_mini_batch)
        else:
            gen = minibatch.feed_forward_minibatches(roll, adv, num_mini_batch)
        for obs, vobs, h0, actions, old_v, ret, masks, old_logp, adv_t in gen:
            values, logp, entropy, _ = policy.evaluate_actions(
                p, obs, vobs, h0, masks, actions, recurrent, concat_vector)
            v_loss, a_loss = ppo_losses(values, logp, entropy, old_v, ret, old_logp, adv_t,
                                        clip_param, use_clipped_value_loss)
            state.optimizer.zero_grad()
            (v_loss * value_loss_coef + a_loss - entropy * entropy_coef).backward()
            gnorm = torch.nn.utils.clip_grad_norm_(list(p.values()), max_grad_norm)
            if on_minibatch is not None:
                on_minibatch(k, dict(value_loss=v_loss.item(), action_loss=a_loss.item(),
                                     entropy=entropy.item(), grad_norm=float(gnorm),
                                     grads={n: (None if t.grad is None else t.grad.detach().clone())
                                            for n, t in p.items()}))
            state.optimizer.step()
            tot_v += v_loss.item()
            tot_a += a_loss.item()
            tot_e += entropy.item()
            k += 1
            if max_minibatches is not None and k >= max_minibatches:
                return tot_v / k, tot_a / k, tot_e / k
    n = ppo_epoch * num_mini_batch
    return tot_v / n, tot_a / n, tot_e / n
