"""Oracle (test infrastructure): the actor-critic network as pure functions.

Restates PKG/model.py:54-79 (act / get_value / evaluate_actions),
PKG/model.py:111-166 (GRU with mask reset), PKG/model.py:169-199 (CNNBase) and
PKG/distributions.py:18-27,54-68 (Categorical head) on torch-CPU.  Parameters
are a flat ``dict`` keyed by the reference ``state_dict`` names:

    base.main.{0,2,4}.{weight,bias}   conv 8x8/4, 4x4/2, 3x3/1
    base.main.7.{weight,bias}         linear 1568 -> H
    base.gru.{weight_ih_l0,weight_hh_l0,bias_ih_l0,bias_hh_l0}   (recurrent only)
    base.critic_linear.{weight,bias}
    dist.linear.{weight,bias}

``concat_vector`` selects model variant A (vector obs concatenated to the
features, PKG/model.py:195) or B (S001/ppo/model.py:195, concat commented out).
"""
import numpy as np
import torch
import torch.nn.functional as F


def init_params(num_inputs, num_actions, vector_obs_len=0, recurrent=False, hidden_size=512,
                concat_vector=True):
    """Reference initialisers, consumed in the reference's construction order
    (GRU first -- NNBase.__init__, model.py:89-95 -- then trunk, critic, dist),
    so ``torch.manual_seed(s)`` before this call reproduces ``Policy(...)``'s weights."""
    p = {}
    V = vector_obs_len if concat_vector else 0
    gru_in = hidden_size + vector_obs_len      # NNBase gets hidden+V regardless of variant (model.py:171)

    def ortho(shape, gain):
        w = torch.empty(*shape)
        torch.nn.init.orthogonal_(w, gain=gain)
        return w

    if recurrent:
        # nn.GRU construction draws uniform(-k, k) for 4 tensors first, then the
        # reference overwrites them (orthogonal weights, zero biases).
        g = torch.nn.GRU(gru_in, hidden_size)
        for name, prm in g.named_parameters():
            if "bias" in name:
                torch.nn.init.constant_(prm, 0)
            elif "weight" in name:
                torch.nn.init.orthogonal_(prm)
        for name, prm in g.named_parameters():
            p["base.gru." + name] = prm.detach().clone()
    relu_gain = torch.nn.init.calculate_gain("relu")
    shapes = [("base.main.0", (32, num_inputs, 8, 8)), ("base.main.2", (64, 32, 4, 4)),
              ("base.main.4", (32, 64, 3, 3)), ("base.main.7", (hidden_size, 32 * 7 * 7))]
    for name, shp in shapes:
        # nn.Conv2d / nn.Linear constructors consume RNG (kaiming_uniform + bias) before init_
        if len(shp) == 4:
            torch.nn.Conv2d(shp[1], shp[0], shp[2])
        else:
            torch.nn.Linear(shp[1], shp[0])
        p[name + ".weight"] = ortho(shp, relu_gain)
        p[name + ".bias"] = torch.zeros(shp[0])
    crit_in = hidden_size if recurrent else hidden_size + V
    torch.nn.Linear(crit_in, 1)
    p["base.critic_linear.weight"] = ortho((1, crit_in), 1.0)
    p["base.critic_linear.bias"] = torch.zeros(1)
    torch.nn.Linear(hidden_size, num_actions)
    p["dist.linear.weight"] = ortho((num_actions, hidden_size), 0.01)
    p["dist.linear.bias"] = torch.zeros(num_actions)
    return p


def trunk(p, visual, relu_masks=None, pre_out=None):
    """conv-relu x3, flatten (NCHW order), linear-relu  (model.py:176-180,194; no /255).

    Test instruments (never used by the reference path): ``relu_masks`` -- a list of four 0/1 tensors (NCHW for the three
    convolutions, [B,H] for the FC layer) that REPLACE the sign test of each ReLU (x * mask instead of relu(x)); the parity tests
    pass the masks the CUDA path used, to show that its gradient outliers come only from units whose pre-activation sits at
    rounding distance from zero.  ``pre_out`` -- a list that receives the four pre-activations."""
    def act(x, i):
        if pre_out is not None:
            pre_out.append(x.detach())
        return F.relu(x) if relu_masks is None else x * relu_masks[i]
    x = act(F.conv2d(visual, p["base.main.0.weight"], p["base.main.0.bias"], stride=4), 0)
    x = act(F.conv2d(x, p["base.main.2.weight"], p["base.main.2.bias"], stride=2), 1)
    x = act(F.conv2d(x, p["base.main.4.weight"], p["base.main.4.bias"], stride=1), 2)
    x = x.reshape(x.shape[0], -1)
    return act(F.linear(x, p["base.main.7.weight"], p["base.main.7.bias"]), 3)


def _gru_run(p, x_seq, h):
    """x_seq [L,E,I], h [E,H] -> (out [L,E,H], h_last [E,H]) via torch's GRU (gate order r,z,n)."""
    flat = [p["base.gru.weight_ih_l0"], p["base.gru.weight_hh_l0"],
            p["base.gru.bias_ih_l0"], p["base.gru.bias_hh_l0"]]
    out, hn = torch._VF.gru(x_seq, h.unsqueeze(0), flat, True, 1, 0.0, True, False, False)   # train=True (dropout 0): nn.GRU in .train() mode, and cuDNN needs it for backward
    return out, hn.squeeze(0)


def gru_with_resets(p, x, hxs, masks):
    """model.py:111-166.  Rollout call (rows == envs): one step with h*m.
    Training call (rows == T*E): unroll over T, multiplying the carried state by
    masks[t] whenever any env resets at t; zero-free stretches run as one call."""
    if x.shape[0] == hxs.shape[0]:
        out, h = _gru_run(p, x.unsqueeze(0), hxs * masks)
        return out.squeeze(0), h
    E = hxs.shape[0]
    T = x.shape[0] // E
    xs = x.reshape(T, E, x.shape[1])
    mk = masks.reshape(T, E)
    resets = (np.flatnonzero((mk[1:] == 0.0).any(dim=-1).cpu().numpy()) + 1).tolist()   # host sync, as model.py:129-133
    cuts = [0] + resets + [T]
    h = hxs
    outs = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        o, h = _gru_run(p, xs[a:b], h * mk[a].reshape(E, 1))
        outs.append(o)
    return torch.cat(outs, 0).reshape(T * E, -1), h


def base_forward(p, visual, vector, hxs, masks, recurrent, concat_vector=True, relu_masks=None, pre_out=None):
    """(value [B,1], features [B,F], hxs) -- CNNBase.forward, model.py:192-199."""
    x = trunk(p, visual, relu_masks, pre_out)
    if concat_vector:
        x = torch.cat((x, vector), dim=1)
    if recurrent:
        x, hxs = gru_with_resets(p, x, hxs, masks)
    value = F.linear(x, p["base.critic_linear.weight"], p["base.critic_linear.bias"])
    return value, x, hxs


def categorical(p, feats):
    logits = F.linear(feats, p["dist.linear.weight"], p["dist.linear.bias"])
    return torch.distributions.Categorical(logits=logits)


def evaluate_actions(p, visual, vector, hxs, masks, action, recurrent, concat_vector=True, relu_masks=None, pre_out=None):
    """model.py:72-79 -> (value [B,1], log_prob [B,1], mean entropy, hxs)."""
    value, feats, hxs = base_forward(p, visual, vector, hxs, masks, recurrent, concat_vector, relu_masks, pre_out)
    d = categorical(p, feats)
    logp = d.log_prob(action.squeeze(-1)).reshape(action.shape[0], -1).sum(-1).unsqueeze(-1)
    return value, logp, d.entropy().mean(), hxs


def act(p, visual, vector, hxs, masks, recurrent, deterministic=False, concat_vector=True):
    """model.py:54-66 -> (value, action [N,1] int64, log_prob [N,1], hxs)."""
    value, feats, hxs = base_forward(p, visual, vector, hxs, masks, recurrent, concat_vector)
    d = categorical(p, feats)
    # torch.multinomial(probs, 1, True) is what Categorical.sample() calls; it is used directly
    # because the reference monkey-patches Categorical.sample process-wide (distributions.py:20-21).
    action = d.probs.argmax(dim=-1, keepdim=True) if deterministic else torch.multinomial(d.probs, 1, True)
    logp = d.log_prob(action.squeeze(-1)).reshape(action.shape[0], -1).sum(-1).unsqueeze(-1)
    return value, action, logp, hxs


def get_value(p, visual, vector, hxs, masks, recurrent, concat_vector=True):
    return base_forward(p, visual, vector, hxs, masks, recurrent, concat_vector)[0]


def gru_cell_stepwise(p, x, hxs, masks):
    """Per-step statement h_t = GRU(x_t, h_{t-1} * m_t) used to cross-check
    gru_with_resets (identical for binary masks)."""
    E = hxs.shape[0]
    T = x.shape[0] // E
    xs = x.reshape(T, E, -1)
    mk = masks.reshape(T, E, 1)
    h = hxs
    outs = []
    for t in range(T):
        h = torch.gru_cell(xs[t], h * mk[t], p["base.gru.weight_ih_l0"], p["base.gru.weight_hh_l0"],
                           p["base.gru.bias_ih_l0"], p["base.gru.bias_hh_l0"])
        outs.append(h)
    return torch.stack(outs, 0).reshape(T * E, -1), h
